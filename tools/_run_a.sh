python -m pytest tests -m gpu -x -q -k "q2m or c4 or group or fuzz" 2>&1 | tail -15
