python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_cfg3_final.json 2> gpurun_out/r2_cfg3_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref_arm.json 2> gpurun_out/r2_ref_arm.err
timeout 400 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r2_cfg1_f.json 2> gpurun_out/r2_cfg1_f.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 2 > gpurun_out/r2_cfg5_g.json 2> gpurun_out/r2_cfg5_g.err
cat gpurun_out/r2_smoke.log | tail -2; tail -2 gpurun_out/r2_cfg3_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_cfg3_final.json").read().strip().splitlines()[-1])
print("cfg3", d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline_ntt"]["frac"], d["extra"]["sharded_prove"]["matches_oracle"], d["extra"].get("prove_default_N2_R2_ms"), d["extra"].get("batch_default_proofs_per_s_per_gpu"), d["cpu_baseline"]["value"], d["wall_s_timed_region"])
d=json.loads(open("gpurun_out/r2_ref_arm.json").read().strip().splitlines()[-1]); print("ref", d["value"], d["ms_per_step"])
for f in ("r2_cfg1_f","r2_cfg5_g"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d["extra"].get("proof_graphs"), {k:(round(v["proofs_per_s"]), v["s_per_batch_each_step_this_rank"]) for k,v in d["extra"].items() if "variant" in k}, d.get("cpu_baseline"))
        if "sweep" in d["extra"]:
            for r in d["extra"]["sweep"]: print("  ", r["N"], r["R"], round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"], {k:round(v,3) for k,v in r.get("with_crs_cache",{}).items() if k.endswith("_ms")}, r.get("cpu_prove_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
