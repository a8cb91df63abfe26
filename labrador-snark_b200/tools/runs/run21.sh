#!/bin/bash
# unrolled double rounds + split shuffles everywhere: K_A / expand A-B through kbench, then tests and the benches
set -x
cd labrador-snark_b200/tools && mkdir -p bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../csrc -o bin/kbench kbench.cu 2> ../../gpurun_out/r2b_kbench_build.log &
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../csrc -DLAB_KA_VAR=0 -o bin/kbench_ka0 kbench.cu 2> ../../gpurun_out/r2b_kbench_build0.log &
wait; cd ../..
timeout 300 labrador-snark_b200/tools/bin/kbench_ka0 > gpurun_out/r2b_kbench_ka0.jsonl 2> gpurun_out/r2b_kbench_ka0.err
timeout 300 labrador-snark_b200/tools/bin/kbench > gpurun_out/r2b_kbench_ka13.jsonl 2> gpurun_out/r2b_kbench_ka13.err
grep -h "k_commit_inner\|0x00000000\|0x00010001" gpurun_out/r2b_kbench_ka0.jsonl gpurun_out/r2b_kbench_ka13.jsonl | cut -c1-220
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_test2.log 2>&1; tail -3 gpurun_out/r2b_test2.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_cfg3_b.json 2> gpurun_out/r2b_cfg3_b.err
timeout 400 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r2b_cfg1_b.json 2> gpurun_out/r2b_cfg1_b.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 2 > gpurun_out/r2b_cfg5_b.json 2> gpurun_out/r2b_cfg5_b.err
python - <<'PY'
import json
for f in ("r2b_cfg3_b",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["extra"]["sharded_prove"]["matches_oracle"], d["extra"].get("prove_default_N2_R2_ms"), d["extra"].get("batch_default_proofs_per_s_per_gpu"), d["wall_s_timed_region"])
    except Exception as e: print(f, "ERR", e)
for f in ("r2b_cfg1_b","r2b_cfg5_b"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d["extra"].get("proof_graphs"), {k:(round(v["proofs_per_s"]), v["s_per_batch_each_step_this_rank"]) for k,v in d["extra"].items() if "variant" in k})
        if "sweep" in d["extra"]:
            for r in d["extra"]["sweep"]: print("  ", r["N"], r["R"], round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"])
    except Exception as e:
        print(f, "ERR", e)
PY
