set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu53.log; tail -6 gpurun_out/pytest_gpu53.log
timeout 600 python bench.py --strong --steps 50 --warmup 3 --no-cpu --temperature 0 > gpurun_out/bench_strong_1gpu_T0.json 2> gpurun_out/bench_strong_1gpu_T0.err; cut -c1-400 gpurun_out/bench_strong_1gpu_T0.json; tail -2 gpurun_out/bench_strong_1gpu_T0.err
python -c "
import json
for f in ('bench_strong_1gpu_T0.json',):
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d['roofline']['stage_ms'], d['roofline']['step_frac'])"
