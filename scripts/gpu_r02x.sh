set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -6
