#!/bin/bash
set -x
for ov in 1 0 1; do
LAB_GC_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2b_cfg3_2gpu_ov$ov.json 2> gpurun_out/r2b_cfg3_2gpu_ov$ov.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_cfg3_2gpu_ov$ov.json").read().strip().splitlines()[-1])
print("overlap $ov", d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("steps"), d["e2e"].get("ms_each"))
PY
done
