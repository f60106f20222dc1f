set -x
for O in '{}' '{"rows_mode":0}' '{"chunk_long":16,"chunk_short":4,"tail_pct":20}' '{"chunk_long":24,"chunk_short":4,"tail_pct":25}' '{"chunks":6}'; do
timeout 300 python scripts/profile_workload.py c4 128 20 0 "$O" 2>&1 | grep -v "^$" | tail -1 | cut -c1-330
done
timeout 300 python scripts/profile_workload.py c4 128 20 300 2>&1 | grep -v "^$" | tail -1 | cut -c1-330
