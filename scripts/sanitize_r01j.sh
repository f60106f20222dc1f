# compute-sanitizer over the pair kernel's data flows added in r01j: recover_u (T = 0 and T > 0), stored u, noise warp
set -x
for O in '{"kernel": 2, "recover_u": 1}' '{"kernel": 2, "recover_u": 0}' '{"kernel": 2, "recover_u": 1, "noise_warp": 1}' '{"kernel": 2, "recover_u": 0, "noise_warp": 2}'; do
for T in 0.0 50.0; do
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanity_tile.py "$O" $T > gpurun_out/memcheck_tmp.log 2>&1; echo "memcheck $O T=$T rc=$?"; tail -2 gpurun_out/memcheck_tmp.log
echo "== memcheck $O T=$T" >> gpurun_out/r01j_sanitizer.txt; tail -2 gpurun_out/memcheck_tmp.log >> gpurun_out/r01j_sanitizer.txt
done
done
for O in '{"kernel": 2, "recover_u": 1}' '{"kernel": 2, "recover_u": 1, "noise_warp": 1}'; do
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanity_tile.py "$O" 50.0 > gpurun_out/racecheck_tmp.log 2>&1; echo "racecheck $O rc=$?"; tail -2 gpurun_out/racecheck_tmp.log
echo "== racecheck $O T=50" >> gpurun_out/r01j_sanitizer.txt; tail -2 gpurun_out/racecheck_tmp.log >> gpurun_out/r01j_sanitizer.txt
done
