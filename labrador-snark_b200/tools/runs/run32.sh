#!/bin/bash
# last validation of the round: the committed tree as the driver will run it
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_test_last.log 2>&1; tail -3 gpurun_out/r2b_test_last.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke_last.log 2>&1; tail -1 gpurun_out/r2b_smoke_last.log
timeout 600 python bench.py > gpurun_out/r2b_default_last.json 2> gpurun_out/r2b_default_last.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_default_last.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["frac_at_596_ops"], d["roofline"]["traffic"], d["extra"]["sharded_prove"]["matches_oracle"], d["cpu_baseline"]["value"], d["steps"], d["warmup"], d["wall_s_timed_region"])
PY
