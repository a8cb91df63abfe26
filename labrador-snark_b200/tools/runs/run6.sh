python -m pytest tests -m gpu -x -q -k "graph or cpp or rq_add or packed" 2>&1 | tail -15 > gpurun_out/r2_test6.log
python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_cfg1_b.json 2> gpurun_out/r2_cfg1_b.err
LAB_TRACE=1 python bench.py --workload cfg5 --steps 2 --warmup 2 --no-cpu > gpurun_out/r2_cfg5_c.json 2> gpurun_out/r2_cfg5_c.err
python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_cfg5_d.json 2> gpurun_out/r2_cfg5_d.err
tail -5 gpurun_out/r2_test6.log; grep -c "not built" gpurun_out/r2_cfg5_c.err; grep "not built" gpurun_out/r2_cfg5_c.err | head -3
python - <<'PY'
import json
for f in ("r2_cfg1_b","r2_cfg5_c","r2_cfg5_d"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["extra"].get("proof_graphs"), {k:v for k,v in d["extra"].items() if "variant" in k})
    except Exception as e:
        print(f, "ERR", e)
PY
