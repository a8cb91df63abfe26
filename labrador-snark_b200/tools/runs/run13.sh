LAB_BENCH_LIGHT=1 LAB_BENCH_ROWS_DIV=2048 timeout 500 python bench.py --workload cfg4 --steps 1 --warmup 0 > gpurun_out/r2_cfg4_dry.json 2> gpurun_out/r2_cfg4_dry.err
tail -5 gpurun_out/r2_cfg4_dry.err; cat gpurun_out/r2_cfg4_dry.json | cut -c1-3000
