#!/bin/bash
set -x
timeout 300 python labrador-snark_b200/tools/microbench.py gen > gpurun_out/r2b_plain_gen2.log 2>&1; tail -1 gpurun_out/r2b_plain_gen2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gen_planes --launch-skip 1 -c 1 -f -o gpurun_out/prof_gen_planes_r2c python labrador-snark_b200/tools/microbench.py gen > gpurun_out/r2b_ncu_gen2.log 2>&1; tail -2 gpurun_out/r2b_ncu_gen2.log
