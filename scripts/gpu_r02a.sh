# round 2, first GPU pass: smoke, parity tests, queue / chunk-plan / data-flow sweep with per-CTA traces
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 180 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; tail -5 gpurun_out/r02a_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -40 > gpurun_out/r02a_pytest.log; tail -40 gpurun_out/r02a_pytest.log
timeout 900 python scripts/quick_bench.py --trace --T 0,100 '{"verbose":1}' '{"recover_u":0}' '{"chunks":9}' '{"chunks":9,"recover_u":0}' '{"tail_pct":0}' '{"tail_pct":25}' '{"chunk_long":24}' '{"chunk_long":48,"chunk_short":8}' '{"chunk_long":16,"tail_pct":0}' '{"chunks":16}' > gpurun_out/r02a_quick.log 2>&1; grep -v "^    jams" gpurun_out/r02a_quick.log
timeout 300 python scripts/quick_bench.py --dims 256x296x256 --trace --T 0 '{"verbose":1}' '{"chunks":9}' >> gpurun_out/r02a_quick2.log 2>&1; grep -v "^    jams" gpurun_out/r02a_quick2.log
