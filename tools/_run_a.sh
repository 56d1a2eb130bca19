python -m pytest tests -m gpu -x -q -k "q2m" 2>&1 | tail -4
