python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/time_shards.py c3 1 2 8
