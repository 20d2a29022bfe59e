python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; tail -5 gpurun_out/r2_pytest.log
python scripts/gpu_check.py cand > gpurun_out/r2_check.log 2>&1; cat gpurun_out/r2_check.log
rm -f gpurun_out/r2_time.log
for ns in 2 3 4; do SEAM_SCORE_NSEED=$ns python scripts/gpu_time.py 15000 10 15000 >> gpurun_out/r2_time.log 2>&1; done
python scripts/gpu_time.py 10000 10 125000 >> gpurun_out/r2_time.log 2>&1
python scripts/gpu_time.py 10000 3 6250 >> gpurun_out/r2_time.log 2>&1
cat gpurun_out/r2_time.log
