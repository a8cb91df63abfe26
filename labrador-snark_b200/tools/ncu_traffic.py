#!/usr/bin/env python3
"""Launch list of bench.py (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) -> per-kernel totals and
the DRAM traffic of one inner commitment, as profiles/ncu_traffic_<tag>.json (bench.py quotes it as roofline.traffic).
Usage: ncu_traffic.py launches.csv commitments_in_the_run out.json"""
import collections
import csv
import json
import sys

path, n_commit, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iid = hdr.index("ID")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
per = collections.defaultdict(lambda: collections.defaultdict(float))
launches = collections.defaultdict(set)
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("void ", "").replace("lab::", "")
    name = name.split("<")[0]
    v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1)
    per[name][r[im]] += v
    launches[name].add(r[iid])
res = {"source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, LAB_BENCH_LIGHT=1 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e",
       "kernels": {}}
tot_ms = sum(p["gpu__time_duration.sum"] for p in per.values())
for k, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    res["kernels"][k] = {"launches": len(launches[k]), "gpu_time_ms_under_ncu": p["gpu__time_duration.sum"], "share_of_listed_time": p["gpu__time_duration.sum"] / tot_ms,
                         "dram_bytes_read": p["dram__bytes_read.sum"], "dram_bytes_write": p["dram__bytes_write.sum"]}
inner = [k for k in per if k.startswith("k_gen_planes") or k.startswith("k_umma")]
res["inner_commitment_cfg3"] = {
    "calls": n_commit,
    "launches_per_call": {k: len(launches[k]) / n_commit for k in inner},
    "dram_bytes_read_per_call": sum(per[k]["dram__bytes_read.sum"] for k in inner) / n_commit,
    "dram_bytes_write_per_call": sum(per[k]["dram__bytes_write.sum"] for k in inner) / n_commit,
    "gpu_time_ms_per_call_under_ncu": sum(per[k]["gpu__time_duration.sum"] for k in inner) / n_commit,
    "gen_time_share": sum(per[k]["gpu__time_duration.sum"] for k in inner if k.startswith("k_gen_planes")) / sum(per[k]["gpu__time_duration.sum"] for k in inner),
    "shape": "cfg3: N=4096 R=64 rows=262144 (one GPU), generate-then-contract in 4 GB row chunks",
    "algorithmic_bytes": 64 * 4096 * 128 + 64 * 262144 * 256,
    "note": "A is deliberately spilled through HBM as int8 limb planes between k_gen_planes and k_umma_commit"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res["inner_commitment_cfg3"], indent=1))
for k, v in list(res["kernels"].items())[:10]:
    print(k, v["launches"], round(v["gpu_time_ms_under_ncu"], 2), round(v["share_of_listed_time"], 4))
