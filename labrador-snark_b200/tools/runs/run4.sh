python -m pytest tests -m gpu -x -q -k "fiat or graph or packed or cfg3 or full_proof" 2>&1 | tail -15 > gpurun_out/r2_test3.log
python labrador-snark_b200/tools/jlbench.py > gpurun_out/r2_jlbench4.jsonl 2>&1
LAB_NO_GRAPH=1 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_cfg5_nograph.json 2> gpurun_out/r2_cfg5_nograph.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_cfg3_a.json 2> gpurun_out/r2_cfg3_a.err
tail -5 gpurun_out/r2_test3.log; cat gpurun_out/r2_jlbench4.jsonl; tail -3 gpurun_out/r2_cfg3_a.err; tail -3 gpurun_out/r2_cfg5_nograph.err
