python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/time_shards.py c4 1 2 4 8
for s in 2 4 16; do echo "slices $s"; SDPB_Q2_SLICES=$s python tools/time_shards.py c4 8 4; done
