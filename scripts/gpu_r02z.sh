set -x
timeout 300 python scripts/profile_workload.py c2 128 50 300 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
timeout 300 python scripts/profile_workload.py c2 64 100 300 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
timeout 300 python scripts/profile_workload.py c3 256 50 100 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
timeout 300 python scripts/profile_workload.py c3 256 50 0 2>&1 | grep -v "^$" | tail -1 | cut -c1-400
