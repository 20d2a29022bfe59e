python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python scripts/gpu_time.py 15000 10 15000 2>&1 | tail -1
python scripts/gpu_time.py 10000 10 125000 2>&1 | tail -1
