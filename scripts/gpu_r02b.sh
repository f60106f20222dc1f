# round 2, second GPU pass: work-item plan sweep with traces, ncu full capture of the default kernels at T = 100 K, sanitizer
set -x
timeout 900 python scripts/quick_bench.py --trace --T 100,0 '{"verbose":1}' '{"chunk_long":24,"chunk_short":2,"tail_pct":25}' '{"chunk_long":32,"chunk_short":4,"tail_pct":25}' '{"chunk_long":16,"chunk_short":2,"tail_pct":30}' '{"chunk_long":16,"chunk_short":4,"tail_pct":20}' '{"chunk_long":103,"chunk_short":2,"tail_pct":20}' '{"chunk_long":8,"chunk_short":2,"tail_pct":20}' '{"chunk_long":40,"chunk_short":2,"tail_pct":40}' '{"grid":288}' '{"ring":5}' > gpurun_out/r02b_quick.log 2>&1; grep -v "^    jams" gpurun_out/r02b_quick.log
timeout 300 python scripts/quick_bench.py --dims 256x296x256 --trace --T 0 '{"verbose":1}' > gpurun_out/r02b_quick2.log 2>&1; grep -v "^    jams" gpurun_out/r02b_quick2.log
timeout 300 python scripts/quick_bench.py --dims 512x512x512 --trace --T 100 --steps 5 '{"verbose":1}' > gpurun_out/r02b_quick3.log 2>&1; grep -v "^    jams" gpurun_out/r02b_quick3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_pair -s 6 -c 2 -f -o gpurun_out/r02b_pair_T100 python scripts/profile_step.py 256 6 2 100 > gpurun_out/r02b_ncu.log 2>&1; tail -3 gpurun_out/r02b_ncu.log
for O in '{"kernel": 2}' '{"kernel": 2, "recover_u": 0}' '{"kernel": 2, "tile_y": 2, "tile_z": 16, "chunk_long": 4, "chunk_short": 2, "tail_pct": 50}'; do
for T in 0.0 50.0; do
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanity_tile.py "$O" $T > gpurun_out/memcheck_tmp.log 2>&1; echo "memcheck $O T=$T rc=$?"; tail -2 gpurun_out/memcheck_tmp.log
echo "== memcheck $O T=$T" >> gpurun_out/r02b_sanitizer.txt; tail -2 gpurun_out/memcheck_tmp.log >> gpurun_out/r02b_sanitizer.txt
done
done
for O in '{"kernel": 2}' '{"kernel": 2, "recover_u": 0}'; do
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanity_tile.py "$O" 50.0 > gpurun_out/racecheck_tmp.log 2>&1; echo "racecheck $O rc=$?"; tail -2 gpurun_out/racecheck_tmp.log
echo "== racecheck $O T=50" >> gpurun_out/r02b_sanitizer.txt; tail -2 gpurun_out/racecheck_tmp.log >> gpurun_out/r02b_sanitizer.txt
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -k "slab_decomposition and 2r_fold and True-2" > gpurun_out/memcheck_fold.log 2>&1; echo "memcheck fold rc=$?"; tail -3 gpurun_out/memcheck_fold.log
echo "== memcheck slab decomposition, folded halo handshake" >> gpurun_out/r02b_sanitizer.txt; tail -3 gpurun_out/memcheck_fold.log >> gpurun_out/r02b_sanitizer.txt
