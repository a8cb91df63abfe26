#!/bin/bash
# unroll variants of the generator (kbench gen) and K_MV variants (cfg5 through variant builds of the library)
set -x
cd labrador-snark_b200/tools && mkdir -p bin && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../csrc -o bin/kbench kbench.cu 2> ../../gpurun_out/r2b_kbench_build.log; cd ../..
timeout 300 labrador-snark_b200/tools/bin/kbench gen > gpurun_out/r2b_kbench_gen2.jsonl 2> gpurun_out/r2b_kbench_gen2.err
cat gpurun_out/r2b_kbench_gen2.jsonl
cp labrador-snark_b200/liblabrador_b200.so /tmp/lib_default.so
for v in default mv1 mv5 mv13 default; do
  if [ $v = default ]; then cp /tmp/lib_default.so labrador-snark_b200/liblabrador_b200.so; else cp labrador-snark_b200/tools/variants/lib_$v.so labrador-snark_b200/liblabrador_b200.so; fi
  timeout 200 python bench.py --workload cfg5 --steps 4 --warmup 2 --no-cpu > gpurun_out/r2b_cfg5_$v.json 2> gpurun_out/r2b_cfg5_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2b_cfg5_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value'],1), {k:(round(v['proofs_per_s']), v['s_per_batch_each_step_this_rank']) for k,v in d['extra'].items() if 'variant' in k})
"
done
cp /tmp/lib_default.so labrador-snark_b200/liblabrador_b200.so
