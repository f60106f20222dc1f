set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; lscpu | grep "Model name"
ls MEASURED_PEAKS.json 2>/dev/null && cat MEASURED_PEAKS.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
timeout 600 python scripts/quick_bench.py 256 > gpurun_out/quick_bench.log 2>&1; cat gpurun_out/quick_bench.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py 256 6 1 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_tma -s 4 -c 2 -o gpurun_out/prof_tma_r01 -f python scripts/profile_step.py 256 4 1 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_direct -s 4 -c 2 -o gpurun_out/prof_direct_r01 -f python scripts/profile_step.py 256 4 0 > gpurun_out/ncu_full0.log 2>&1; tail -3 gpurun_out/ncu_full0.log
