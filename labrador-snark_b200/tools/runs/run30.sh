#!/bin/bash
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload cfg5 --steps 4 --warmup 2 > gpurun_out/r2b_cfg5_8gpu.json 2> gpurun_out/r2b_cfg5_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload prove --steps 3 --warmup 2 > gpurun_out/r2b_prove_8gpu.json 2> gpurun_out/r2b_prove_8gpu.err
python - <<'PY'
import json
for f in ("r2b_cfg5_8gpu","r2b_prove_8gpu"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"],1), d["ms_per_step"], {k:(round(v["proofs_per_s"]), v["s_per_batch_each_step_this_rank"]) for k,v in d["extra"].items() if "variant" in k}, (d.get("cpu_baseline") or {}).get("bit_exact_sample"), {k:v for k,v in d["extra"].items() if k in ("transcript_sha256",)})
    except Exception as e: print(f, "ERR", e)
PY
