python scripts/gpu_check.py agg 2>&1 | grep "T=10 ragged=None"
python scripts/gpu_time.py 15000 10 15000 2>&1 | tail -1 | cut -c1-150
python scripts/gpu_time.py 100000 10 0 2>&1 | tail -1
