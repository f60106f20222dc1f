set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -5
timeout 600 python scripts/other_configs.py 2>&1 | grep -v "^$"
for O in '{"motif_split":1}' '{"motif_split":2}' '{"motif_split":2,"tile_y":2,"tile_z":128}' '{"motif_split":2,"tile_y":8,"tile_z":32}'; do
timeout 300 python scripts/profile_workload.py c2 128 60 300 "$O" 2>&1 | grep -v "^$" | tail -2
done
