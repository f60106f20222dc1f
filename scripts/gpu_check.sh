set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu54.log; tail -8 gpurun_out/pytest_gpu54.log
O='{"kernel": 2, "recover_u": 1, "noise_warp": 1}'
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanity_tile.py "$O" 50.0 > gpurun_out/racecheck_tmp.log 2>&1; echo "racecheck $O rc=$?"; tail -2 gpurun_out/racecheck_tmp.log
echo "== racecheck $O T=50 (per-thread arrivals on the noise ring barriers)" >> gpurun_out/r01j_sanitizer.txt; tail -2 gpurun_out/racecheck_tmp.log >> gpurun_out/r01j_sanitizer.txt
export JB_QB_EXTRA='[{"recover_u":0,"noise_warp":2}]'
timeout 300 python scripts/quick_bench.py 256 100 > gpurun_out/quick_bench54.log 2>&1; grep -v "^    jams" gpurun_out/quick_bench54.log
