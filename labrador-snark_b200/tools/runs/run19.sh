#!/bin/bash
# generator variants (kbench gen) + ncu --set full of k_gen_planes after the word-3 cone trimming
set -x
cd labrador-snark_b200/tools && mkdir -p bin && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../csrc -o bin/kbench kbench.cu 2> ../../gpurun_out/r2b_kbench_build.log; cd ../..
timeout 300 labrador-snark_b200/tools/bin/kbench gen > gpurun_out/r2b_kbench_gen.jsonl 2> gpurun_out/r2b_kbench_gen.err
cat gpurun_out/r2b_kbench_gen.jsonl
timeout 300 python labrador-snark_b200/tools/microbench.py gen > gpurun_out/r2b_plain_gen.log 2>&1; tail -2 gpurun_out/r2b_plain_gen.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gen_planes --launch-skip 1 -c 1 -f -o gpurun_out/prof_gen_planes_r2b python labrador-snark_b200/tools/microbench.py gen > gpurun_out/r2b_ncu_gen.log 2>&1; tail -3 gpurun_out/r2b_ncu_gen.log
