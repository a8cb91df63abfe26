timeout 600 python -m pytest tests -m gpu -x -q -k "graph or full_proof or prove_batch or jl_rejection or fiat" 2>&1 | tail -15 > gpurun_out/r2_test8.log
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_cfg1_c.json 2> gpurun_out/r2_cfg1_c.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_cfg5_e.json 2> gpurun_out/r2_cfg5_e.err
tail -5 gpurun_out/r2_test8.log
python - <<'PY'
import json
for f in ("r2_cfg1_c","r2_cfg5_e"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["extra"].get("proof_graphs"), {k:v for k,v in d["extra"].items() if "variant" in k})
        if "sweep" in d["extra"]:
            for r in d["extra"]["sweep"]: print(r["N"], r["R"], round(r["prove_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"])
    except Exception as e:
        print(f, "ERR", e)
PY
