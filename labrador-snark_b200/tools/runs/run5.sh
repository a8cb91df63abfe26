python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_test5.log
python bench.py --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r2_cfg5_b.json 2> gpurun_out/r2_cfg5_b.err
python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r2_cfg2_a.json 2> gpurun_out/r2_cfg2_a.err
tail -5 gpurun_out/r2_test5.log; tail -2 gpurun_out/r2_cfg5_b.err; tail -2 gpurun_out/r2_cfg2_a.err
