python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/gpu_check.py cand 2>&1 | grep cand | cut -c1-60,150-400
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:"rescore|score_topk" -s 8 -c 2 python scripts/gpu_time.py 15000 10 15000 2>&1 | grep -E "inst_executed|time_duration"
python scripts/gpu_time.py 15000 10 15000 2>&1 | tail -1
python scripts/gpu_time.py 10000 10 125000 2>&1 | tail -1
