#!/bin/bash
set -x
timeout 300 python labrador-snark_b200/tools/ncu_targets.py mv > gpurun_out/r2b_plain_mv.log 2>&1; tail -2 gpurun_out/r2b_plain_mv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_crs_matvec -c 1 -f -o gpurun_out/prof_kmv_r2b python labrador-snark_b200/tools/ncu_targets.py mv > gpurun_out/r2b_ncu_mv.log 2>&1; tail -2 gpurun_out/r2b_ncu_mv.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_cfg3_d.json 2> gpurun_out/r2b_cfg3_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_cfg3_d.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["roofline"]["frac"], d["extra"]["crs_resident"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cache or resident or generate_then" 2>&1 | tail -3
