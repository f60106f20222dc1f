set -x
for O in '{"kernel":4,"tile_y":8,"rows_mode":1}' '{"kernel":4,"tile_y":8,"rows_mode":0}' '{"kernel":4,"tile_y":16,"rows_mode":1}' '{"kernel":4,"tile_y":4,"rows_mode":1}' '{"kernel":4,"tile_y":8,"rows_mode":1,"chunks":8}' '{"kernel":4,"tile_y":8,"rows_mode":2}'; do
timeout 300 python scripts/profile_workload.py c2 128 50 300 "$O" 2>&1 | grep -v "^$" | tail -2 | cut -c1-330
done
timeout 300 python scripts/profile_workload.py c3 256 20 100 '{"kernel":4,"rows_mode":1}' 2>&1 | grep -v "^$" | tail -2 | cut -c1-330
timeout 300 python scripts/profile_workload.py c3 256 20 100 '{"kernel":4,"rows_mode":1,"tile_y":8}' 2>&1 | grep -v "^$" | tail -2 | cut -c1-330
