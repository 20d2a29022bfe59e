python scripts/gpu_check.py agg 2>&1 | grep -c "\[agg\] Q"
python scripts/gpu_time.py 100000 16 0 2>&1 | tail -1
python scripts/gpu_time.py 100000 4 0 2>&1 | tail -1
python scripts/gpu_time.py 10000 4 0 2>&1 | tail -1
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
