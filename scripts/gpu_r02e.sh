set -x
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; cat gpurun_out/r02e_bench.json; tail -5 gpurun_out/r02e_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_bench_reference.err; cat gpurun_out/r02e_bench_reference.json; tail -3 gpurun_out/r02e_bench_reference.err
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -5
