G=${1:-2}
TR="timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $G --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_cfg3_${G}gpu.json 2> gpurun_out/r2_cfg3_${G}gpu.err
LAB_BENCH_PROVE_N=16 $TR bench.py --gpus $G --workload prove --steps 2 --warmup 2 > gpurun_out/r2_prove16_${G}gpu.json 2> gpurun_out/r2_prove16_${G}gpu.err
$TR bench.py --gpus $G --workload cfg5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_cfg5_${G}gpu.json 2> gpurun_out/r2_cfg5_${G}gpu.err
tail -3 gpurun_out/r2_cfg3_${G}gpu.err; tail -3 gpurun_out/r2_prove16_${G}gpu.err; tail -3 gpurun_out/r2_cfg5_${G}gpu.err
python - <<PY
import json
for f in ("r2_cfg3_${G}gpu","r2_prove16_${G}gpu","r2_cfg5_${G}gpu"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],2), d["value"], (d.get("e2e") or {}).get("ms_per_step"), d["extra"].get("sharded_prove"), d["extra"].get("transcript_sha256"), d.get("checks"))
    except Exception as e:
        print(f, "ERR", e)
PY
