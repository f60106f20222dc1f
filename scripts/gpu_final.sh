set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_final.log; tail -12 gpurun_out/pytest_gpu_final.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; tail -4 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json | cut -c1-2200; tail -3 gpurun_out/bench_final.err
timeout 600 python bench.py --temperature 0 --no-cpu > gpurun_out/bench_final_T0.json 2> gpurun_out/bench_final_T0.err; cat gpurun_out/bench_final_T0.json | cut -c1-2200; tail -3 gpurun_out/bench_final_T0.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-600 gpurun_out/bench_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_final_ncu.log 2>&1; tail -1 gpurun_out/bench_final_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 4 -c 2 -f -o gpurun_out/r01j_pair_default_T0 python scripts/profile_step.py 256 4 2 0 > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 4 -c 2 -f -o gpurun_out/r01j_pair_default_T100 python scripts/profile_step.py 256 4 2 100 > gpurun_out/ncu_final2.log 2>&1; tail -2 gpurun_out/ncu_final2.log
