mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/time_cash_kinds.py 2>&1 | tail -10
bash tools/_audit.sh 2>&1 | tail -40
