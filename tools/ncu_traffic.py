"""Turn an ncu CSV (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum over one
forward) into profiles/r1_traffic.json: per kernel family, launches, total time and DRAM bytes per launch.
Usage: python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r1_traffic.json"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
fam = collections.defaultdict(lambda: {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
seen = collections.defaultdict(set)
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
for r in csv.DictReader(lines):
    name = r["Kernel Name"]
    m = re.search(r"(gemm2_kernel|attention2_kernel|attention_kernel|gn_stats_kernel|gn_apply_kernel|gn_fused_kernel|layernorm\w*|scaleu_\w+)", name)
    key = m.group(1) if m else "other"
    if key.startswith("layernorm"):
        key = "layernorm_kernel"
    v = float(r["Metric Value"].replace(",", "")) * unit.get(r["Metric Unit"], 1.0)
    if r["Metric Name"] == "gpu__time_duration.sum":
        fam[key]["time_us"] += v
        fam[key]["launches"] += 1
    elif r["Metric Name"].startswith("dram__bytes"):
        fam[key]["dram_bytes"] += v
out = {}
for k, d in fam.items():
    out[k] = {"launches_per_forward": d["launches"], "time_us_per_forward_cold": round(d["time_us"], 1),
              "dram_bytes_per_launch": d["dram_bytes"] / max(d["launches"], 1),
              "dram_bytes_per_forward": d["dram_bytes"]}
out["_note"] = ("one eager cond+uncond forward at forward batch 8 under ncu (cold caches, serialised): "
                "dram__bytes_read.sum + dram__bytes_write.sum per kernel family")
json.dump(out, open(dst, "w"), indent=1)
for k, d in sorted(out.items()):
    if k != "_note":
        print(k, d)
