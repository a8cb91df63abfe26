#!/bin/bash
set -x
LAB_BENCH_LIGHT=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_light_plain.json 2> gpurun_out/r2b_light_plain.err; tail -c 400 gpurun_out/r2b_light_plain.json
LAB_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_light_ncu.json 2> gpurun_out/r2b_light_ncu.err
wc -l gpurun_out/r2b_launches_bench.csv
python labrador-snark_b200/tools/ncu_traffic.py gpurun_out/r2b_launches_bench.csv 4 gpurun_out/r2b_ncu_traffic.json
