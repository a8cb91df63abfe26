timeout 600 python -m pytest tests -m gpu -x -q -k "graph or full_proof or stage_entry or cache" 2>&1 | tail -5 > gpurun_out/r2_test10.log
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_cfg1_d.json 2> gpurun_out/r2_cfg1_d.err
for W in 96 64 48 32; do
LAB_MV_WARPS_PER_SM=$W timeout 200 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_cfg5_w$W.json 2> gpurun_out/r2_cfg5_w$W.err
done
LAB_MV_WARPS_PER_SM=64 timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_cfg1_w64.json 2> gpurun_out/r2_cfg1_w64.err
tail -3 gpurun_out/r2_test10.log
python - <<'PY'
import json
for f in ("r2_cfg1_d","r2_cfg1_w64","r2_cfg5_w96","r2_cfg5_w64","r2_cfg5_w48","r2_cfg5_w32"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d["extra"].get("proof_graphs"), {k:(round(v["proofs_per_s"]), v["s_per_batch_each_step_this_rank"]) for k,v in d["extra"].items() if "variant" in k})
        if "sweep" in d["extra"]:
            for r in d["extra"]["sweep"]: print("  ", r["N"], r["R"], round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"])
    except Exception as e:
        print(f, "ERR", e)
PY
