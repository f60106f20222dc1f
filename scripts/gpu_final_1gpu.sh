set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err; cat gpurun_out/r02aa_bench.json; tail -5 gpurun_out/r02aa_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02aa_bench_reference.json 2> gpurun_out/r02aa_bench_reference.err; cat gpurun_out/r02aa_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02aa_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02aa_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 6 -c 2 -f -o gpurun_out/r02aa_c3_T100 python scripts/profile_workload.py c3 256 6 100 > gpurun_out/r02aa_ncu_c3.log 2>&1; tail -3 gpurun_out/r02aa_ncu_c3.log
timeout 600 python scripts/other_configs.py 2>&1 | grep -v "^$" > gpurun_out/r02aa_other_configs.log; cat gpurun_out/r02aa_other_configs.log
