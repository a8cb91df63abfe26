#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_test3.log 2>&1; tail -3 gpurun_out/r2b_test3.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_cfg3_c.json 2> gpurun_out/r2b_cfg3_c.err
python - <<'PY'
import json
for f in ("r2b_cfg3_c",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["extra"]["sharded_prove"]["matches_oracle"], d["extra"].get("prove_default_N2_R2_ms"), d["extra"].get("batch_default_proofs_per_s_per_gpu"), d["wall_s_timed_region"], d["extra"]["crs_resident"])
    except Exception as e: print(f, "ERR", e)
PY
