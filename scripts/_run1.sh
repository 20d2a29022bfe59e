python scripts/gpu_check.py agg misc topk > gpurun_out/r3_check.log 2>&1; cat gpurun_out/r3_check.log
python scripts/_dbg_rank.py > gpurun_out/r3_dbg.log 2>&1; tail -12 gpurun_out/r3_dbg.log
rm -f gpurun_out/r3_time.log
python scripts/gpu_time.py 15000 10 15000 >> gpurun_out/r3_time.log 2>&1
python scripts/gpu_time.py 10000 3 6250 >> gpurun_out/r3_time.log 2>&1
python scripts/gpu_time.py 100000 16 0 >> gpurun_out/r3_time.log 2>&1
python scripts/gpu_time.py 100000 64 0 >> gpurun_out/r3_time.log 2>&1
cat gpurun_out/r3_time.log
