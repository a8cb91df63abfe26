timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_test15.log
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r2_cfg1_e.json 2> gpurun_out/r2_cfg1_e.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 2 > gpurun_out/r2_cfg5_f.json 2> gpurun_out/r2_cfg5_f.err
python labrador-snark_b200/tools/ncu_targets.py mul phi > gpurun_out/r2_ncu_targets_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_polymul_regs|k_piT_omega2" -c 4 -o gpurun_out/r2_mulphi python labrador-snark_b200/tools/ncu_targets.py mul phi > gpurun_out/r2_ncu_mulphi.log 2>&1
tail -5 gpurun_out/r2_test15.log; tail -2 gpurun_out/r2_ncu_mulphi.log
python - <<'PY'
import json
for f in ("r2_cfg1_e","r2_cfg5_f"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],3), d["extra"].get("proof_graphs"), {k:(round(v["proofs_per_s"]), v["s_per_batch_each_step_this_rank"]) for k,v in d["extra"].items() if "variant" in k})
        if "sweep" in d["extra"]:
            for r in d["extra"]["sweep"]: print("  ", r["N"], r["R"], round(r["prove_ms"],3), round(r["prove_c_call_ms"],3), round(r["verify_ms"],3), r["launches_per_proof"], r.get("with_crs_cache",{}).get("prove_again_same_crs_ms"), r.get("with_crs_cache",{}).get("verify_after_prove_ms"), r.get("cpu_prove_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
