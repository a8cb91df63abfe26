G=8
LAB_BENCH_LIGHT=1 LAB_BENCH_ROWS_DIV=1 timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --workload cfg4 --steps 1 --warmup 0 > gpurun_out/r2_cfg4_8gpu.json 2> gpurun_out/r2_cfg4_8gpu.err
tail -4 gpurun_out/r2_cfg4_8gpu.err; cut -c1-600 gpurun_out/r2_cfg4_8gpu.json
