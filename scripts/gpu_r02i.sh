set -x
timeout 600 python scripts/other_configs.py > gpurun_out/r02i_other_configs.log 2>&1; cat gpurun_out/r02i_other_configs.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; cat gpurun_out/r02i_bench.json; tail -5 gpurun_out/r02i_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02i_ncu_bench.log 2>&1
