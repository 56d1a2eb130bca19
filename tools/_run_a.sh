python -m pytest tests -m gpu -x -q 2>&1 | tail -3
BATCH_ONLY=1 python tools/time_batch.py 2>&1 | tail -3
