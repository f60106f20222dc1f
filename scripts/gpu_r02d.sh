set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -15 > gpurun_out/r02d_pytest.log; tail -15 gpurun_out/r02d_pytest.log
export JB_TRACE_DUMP=gpurun_out
timeout 900 python scripts/quick_bench.py --trace --T 100,0 '{"verbose":1}' '{"chunk_long":24,"chunk_short":2,"tail_pct":25}' '{"chunk_long":34,"chunk_short":2,"tail_pct":20}' '{"chunk_long":40,"chunk_short":2,"tail_pct":40}' '{"chunk_long":16,"chunk_short":2,"tail_pct":30}' '{"chunk_long":48,"chunk_short":2,"tail_pct":30}' '{"chunks":9}' '{"recover_u":0}' > gpurun_out/r02d_quick.log 2>&1; grep -v "^    jams" gpurun_out/r02d_quick.log
timeout 300 python scripts/quick_bench.py --dims 256x296x256 --trace --T 0 '{"verbose":1}' > gpurun_out/r02d_quick2.log 2>&1; grep -v "^    jams" gpurun_out/r02d_quick2.log
