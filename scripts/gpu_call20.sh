set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu20.log; tail -15 gpurun_out/pytest_gpu20.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke20.log 2>&1; tail -5 gpurun_out/smoke20.log
timeout 900 python bench.py > gpurun_out/bench20.json 2> gpurun_out/bench20.err; cat gpurun_out/bench20.json; tail -3 gpurun_out/bench20.err
timeout 600 python bench.py --impl reference > gpurun_out/bench20_ref.json 2> gpurun_out/bench20_ref.err; cat gpurun_out/bench20_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01g_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench20_ncu.log 2>&1; tail -2 gpurun_out/bench20_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 4 -c 2 -f -o gpurun_out/r01g_pair_default_T100 python scripts/profile_step.py 256 4 2 100 > gpurun_out/ncu20.log 2>&1; tail -2 gpurun_out/ncu20.log
