#!/bin/bash
# round 2, session 2: w3-only ChaCha20 cone + single-fold warp transform + overlapped contraction
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_test1.log 2>&1; tail -3 gpurun_out/r2b_test1.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_cfg3_a.json 2> gpurun_out/r2b_cfg3_a.err
LAB_GC_OVERLAP=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_cfg3_noov.json 2> gpurun_out/r2b_cfg3_noov.err
timeout 400 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r2b_cfg1_a.json 2> gpurun_out/r2b_cfg1_a.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 2 > gpurun_out/r2b_cfg5_a.json 2> gpurun_out/r2b_cfg5_a.err
python - <<'PY'
import json
for f in ("r2b_cfg3_a","r2b_cfg3_noov"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("e2e",{}).get("ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["extra"]["sharded_prove"]["matches_oracle"], d["extra"].get("prove_default_N2_R2_ms"), d["extra"].get("batch_default_proofs_per_s_per_gpu"), d["wall_s_timed_region"])
    except Exception as e: print(f, "ERR", e)
for f in ("r2b_cfg1_a","r2b_cfg5_a"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d["extra"].get("proof_graphs"), {k:(round(v["proofs_per_s"]), v["s_per_batch_each_step_this_rank"]) for k,v in d["extra"].items() if "variant" in k})
        if "sweep" in d["extra"]:
            for r in d["extra"]["sweep"]: print("  ", r["N"], r["R"], round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"])
    except Exception as e:
        print(f, "ERR", e)
PY
