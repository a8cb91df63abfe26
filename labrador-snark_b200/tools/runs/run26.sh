#!/bin/bash
# validation of the final build of this session: GPU suite, smoke, the default bench line, the reference arm, cfg1, cfg5, cfg2
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_test4.log 2>&1; tail -3 gpurun_out/r2b_test4.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -1 gpurun_out/r2b_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_cfg3_final.json 2> gpurun_out/r2b_cfg3_final.err
LAB_GC_OVERLAP=0 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2b_cfg3_noov2.json 2> gpurun_out/r2b_cfg3_noov2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_ref_arm.json 2> gpurun_out/r2b_ref_arm.err
timeout 400 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r2b_cfg1_final.json 2> gpurun_out/r2b_cfg1_final.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 2 > gpurun_out/r2b_cfg5_final.json 2> gpurun_out/r2b_cfg5_final.err
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r2b_cfg2_final.json 2> gpurun_out/r2b_cfg2_final.err
python - <<'PY'
import json
for f in ("r2b_cfg3_final","r2b_cfg3_noov2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["extra"]["sharded_prove"]["matches_oracle"], d["extra"].get("prove_default_N2_R2_ms"), d["extra"].get("batch_default_proofs_per_s_per_gpu"), (d.get("cpu_baseline") or {}).get("value"), d["wall_s_timed_region"])
    except Exception as e: print(f, "ERR", e)
try:
    d=json.loads(open("gpurun_out/r2b_ref_arm.json").read().strip().splitlines()[-1]); print("ref", d["value"], d["ms_per_step"])
except Exception as e: print("ref ERR", e)
for f in ("r2b_cfg1_final","r2b_cfg5_final","r2b_cfg2_final"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d["extra"].get("proof_graphs"), {k:(round(v["proofs_per_s"]), v["s_per_batch_each_step_this_rank"]) for k,v in d["extra"].items() if "variant" in k}, d.get("cpu_baseline"))
        if "sweep" in d["extra"]:
            for r in d["extra"]["sweep"]: print("  ", r.get("N"), r.get("R"), round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"], {k:round(v,3) for k,v in r.get("with_crs_cache",{}).items() if k.endswith("_ms")}, r.get("cpu_prove_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
