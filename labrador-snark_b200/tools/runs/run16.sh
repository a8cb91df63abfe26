timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_test16.log
LAB_TRACE=1 timeout 200 python labrador-snark_b200/tools/cache_trace.py > gpurun_out/r2_cache_trace.log 2>&1
tail -5 gpurun_out/r2_test16.log; grep -v "^\[lab" gpurun_out/r2_cache_trace.log | tail -8
