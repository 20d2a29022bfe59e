python scripts/gpu_check.py cand topk 2>&1 | cut -c1-40,150-420 | tail -14
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/gpu_time.py 15000 10 15000 2>&1 | tail -1
python scripts/gpu_time.py 10000 10 125000 2>&1 | tail -1
SEAM_DEBUG_SCORE_MODE=1 python scripts/gpu_time.py 15000 10 15000 2>&1 | tail -1 | cut -c1-200
