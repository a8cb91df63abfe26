#!/bin/bash
set -x
cd labrador-snark_b200/tools && mkdir -p bin && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../csrc -o bin/kbench kbench.cu 2> ../../gpurun_out/r2b_kbench_build.log; cd ../..
timeout 300 labrador-snark_b200/tools/bin/kbench gen > gpurun_out/r2b_kbench_gen3.jsonl 2> gpurun_out/r2b_kbench_gen3.err
cut -c1-230 gpurun_out/r2b_kbench_gen3.jsonl
