set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu37.log; tail -8 gpurun_out/pytest_gpu37.log
export JB_QB_EXTRA='[{"kernel":2},{"kernel":3}]'
timeout 900 python scripts/quick_bench.py 256 0,100 > gpurun_out/quick_bench37.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench37.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench37.json 2>/dev/null; cut -c1-300 gpurun_out/bench37.json; python -c "
import json; d=json.load(open('gpurun_out/bench37.json')); print(d['roofline']['stage_ms'], d['roofline']['frac'], d['clocks'], d['e2e']['value'])"
