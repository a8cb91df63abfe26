#!/bin/bash
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2b_cfg3_2gpu.json 2> gpurun_out/r2b_cfg3_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5 --steps 4 --warmup 2 > gpurun_out/r2b_cfg5_2gpu.json 2> gpurun_out/r2b_cfg5_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload prove --steps 3 --warmup 2 > gpurun_out/r2b_prove_2gpu.json 2> gpurun_out/r2b_prove_2gpu.err
python - <<'PY'
import json
for f in ("r2b_cfg3_2gpu","r2b_cfg5_2gpu","r2b_prove_2gpu"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"],1), d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d["extra"].get("sharded_prove") or {}).get("matches_oracle"), d.get("checks"), {k:v for k,v in d["extra"].items() if k in ("transcript_sha256","digest","all_ranks_equal")})
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r2b_cfg3_2gpu.err
