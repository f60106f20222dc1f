set -x
export JB_TRACE_DUMP=gpurun_out
timeout 600 python scripts/quick_bench.py --trace --T 100 '{"verbose":1}' '{"chunk_long":24,"chunk_short":2,"tail_pct":25}' '{"chunk_long":34,"chunk_short":2,"tail_pct":20}' '{"chunk_long":34,"chunk_short":2,"tail_pct":30}' '{"chunks":9}' > gpurun_out/r02c_quick.log 2>&1; grep -v "^    jams" gpurun_out/r02c_quick.log
