#!/bin/bash
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2b_cfg3_8gpu.json 2> gpurun_out/r2b_cfg3_8gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_cfg3_8gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"],1), d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["extra"]["sharded_prove"]["matches_oracle"], d["extra"]["sharded_prove"]["all_ranks_equal"], d["checks"], d["clocks"])
PY
tail -2 gpurun_out/r2b_cfg3_8gpu.err
