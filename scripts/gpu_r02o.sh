set -x
for O in '{"rows_mode":0}' '{"rows_mode":1}' '{"rows_mode":1,"chunks":9}' '{"rows_mode":1,"chunks":18}' '{"rows_mode":1,"chunks":23}' '{"rows_mode":1,"chunk_long":16,"chunk_short":4,"tail_pct":20}' '{"rows_mode":1,"motif_split":1}'; do
timeout 300 python scripts/profile_workload.py c4 128 20 0 "$O" 2>&1 | grep -v "^$" | tail -1 | cut -c1-330
done
timeout 300 python scripts/profile_workload.py c4 128 20 300 2>&1 | grep -v "^$" | tail -1
timeout 300 python scripts/profile_workload.py c4 64 20 0 2>&1 | grep -v "^$" | tail -1
