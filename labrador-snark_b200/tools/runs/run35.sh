#!/bin/bash
set -x
for p in 1 0; do
LAB_STREAM_PRIO=$p LAB_BENCH_CFG1_MAX_N=4 timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2b_cfg1_prio$p.json 2> gpurun_out/r2b_cfg1_prio$p.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_cfg1_prio$p.json").read().strip().splitlines()[-1])
print("prio $p")
for r in d["extra"]["sweep"]: print("  ", r.get("N"), r.get("R"), round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"], {k:round(v,3) for k,v in r.get("with_crs_cache",{}).items() if k.endswith("_ms")})
PY
done
