timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/gpu_check.py agg topk misc > gpurun_out/r10_memcheck.log 2>&1; echo memcheck rc=$?
grep -E "ERROR SUMMARY|Invalid|error" gpurun_out/r10_memcheck.log | head -10
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/gpu_check.py agg > gpurun_out/r10_racecheck.log 2>&1; echo racecheck rc=$?
grep -E "RACECHECK SUMMARY|hazard|Race" gpurun_out/r10_racecheck.log | head -10
