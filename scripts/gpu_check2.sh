set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu41.log; tail -8 gpurun_out/pytest_gpu41.log
export JB_QB_EXTRA='[{"kernel":2}]'
timeout 900 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench41.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench41.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench41.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench41.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'], d['roofline']['frac'], d['clocks'], d['e2e']['value'])"
