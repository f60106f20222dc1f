set -x
for O in '{}' '{"chunks":16}' '{"chunks":32}' '{"chunks":21}' '{"tile_y":2,"chunks":8}' '{"tile_y":2,"chunks":16}' '{"tile_y":2,"tile_z":32,"chunks":8}' '{"tile_y":4,"tile_z":32,"chunks":16}'; do
timeout 300 python scripts/profile_workload.py c2 64 200 300 "$O" 2>&1 | grep -v "^$" | tail -2 | cut -c1-400
done
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -x -k "deep_template" 2>&1 | tail -2
