set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/time_search.py 1 20
timeout 300 python scripts/time_search.py 64 20
timeout 300 python scripts/time_search.py 2048 6
timeout 300 python scripts/time_search.py 256 6
