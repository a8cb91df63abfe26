#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "verif or full_proof or fiat or cache or graph or cpp_header or rejected or batch" > gpurun_out/r2b_test_verify2.log 2>&1; tail -3 gpurun_out/r2b_test_verify2.log
LAB_BENCH_CFG1_MAX_N=4 timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2b_cfg1_vstage.json 2> gpurun_out/r2b_cfg1_vstage.err
python - <<'PY'
import json
for f in ("r2b_cfg1_vstage",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        for r in d["extra"]["sweep"]: print("  ", r.get("N"), r.get("R"), round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"], {k:round(v,3) for k,v in r.get("with_crs_cache",{}).items() if k.endswith("_ms")})
    except Exception as e:
        print(f, "ERR", e)
PY
