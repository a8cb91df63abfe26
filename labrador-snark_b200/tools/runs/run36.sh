#!/bin/bash
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_cfg3_prio.json 2> gpurun_out/r2b_cfg3_prio.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_cfg3_prio.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["roofline"]["frac"], d["extra"]["sharded_prove"]["matches_oracle"], d["extra"].get("prove_default_N2_R2_ms"), d["extra"].get("verify_default_N2_R2_ms"), d["extra"].get("verify_default_accepts"))
PY
