# compute-sanitizer over small runs of the stage kernels of the end-of-round-2 state (memcheck: out-of-bounds / misaligned
# accesses in global and shared memory; racecheck: shared-memory hazards of the plane rings, item ring, face counters, tables)
OUT=gpurun_out/r02an_sanitizer.txt
: > $OUT
run() {  # tool, json options, T, extra python (workload override)
  echo "== $1 $2 T=$3 $4" >> $OUT
  timeout 900 compute-sanitizer --tool $1 --error-exitcode 7 python scripts/sanity_tile.py "$2" $3 $4 > gpurun_out/sanitize_tmp.log 2>&1
  echo "rc=$?" >> $OUT
  grep "opts\|ERROR SUMMARY\|RACECHECK SUMMARY\|hazard\|Invalid\|Error" gpurun_out/sanitize_tmp.log | head -8 >> $OUT
}
for T in 0.0 50.0; do
run memcheck '{"kernel": 2}' $T
run memcheck '{"kernel": 2, "recover_u": 0}' $T
run memcheck '{"kernel": 2, "tile_y": 2, "tile_z": 16, "chunk_long": 4, "chunk_short": 2, "tail_pct": 50}' $T
run memcheck '{"kernel": 4}' $T
run memcheck '{"kernel": 4, "tile_y": 4, "chunks": 3}' $T
run memcheck '{"kernel": 4}' $T c4
run memcheck '{"kernel": 2}' $T bcc
done
run racecheck '{"kernel": 2}' 50.0
run racecheck '{"kernel": 4}' 50.0
run racecheck '{"kernel": 4}' 0.0 c4
run racecheck '{"kernel": 2}' 50.0 bcc
echo "== memcheck slab decomposition in one process (folded halo handshake: pair, rows and RK4-ring kernels; general neighbour list), biquadratic, RK4 on the ring, random alloy" >> $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x --timeout=1200 -k "(slab and 2-True and (2r_fold or rows_fold or rk4_fold or pairs)) or biquadratic or (rk4_T0 and ring and not small) or impurities or rk4_thermal" > gpurun_out/sanitize_tmp.log 2>&1
echo "rc=$?" >> $OUT
grep "passed\|failed\|ERROR SUMMARY" gpurun_out/sanitize_tmp.log | tail -3 >> $OUT
echo "== racecheck RK4 on the ring (thermal, AC field)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x --timeout=800 -k "rk4_thermal and ring and not small" > gpurun_out/sanitize_tmp.log 2>&1
echo "rc=$?" >> $OUT
grep "passed\|failed\|RACECHECK SUMMARY" gpurun_out/sanitize_tmp.log | tail -3 >> $OUT
cat $OUT
