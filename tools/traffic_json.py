"""profiles/traffic.json from the ncu csv written by tools/gpu_traffic.sh (developer tool).
usage: traffic_json.py gpurun_out/traffic_<tag>.csv profiles/<tag>_traffic.csv"""
import collections
import csv
import json
import os
import re
import shutil
import sys

src, kept = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = collections.defaultdict(lambda: collections.defaultdict(float))
ids = collections.defaultdict(set)
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
for r in rows:
    k = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("be::", ""))
    agg[k][r[12]] += float(r[14].replace(",", "")) * MULT.get(r[13], 1)
    ids[k].add(r[0])
out = {}
for k, v in agg.items():
    n = len(ids[k])
    tr = v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
    out[k] = tr / n
    print(f"{k:18s} launches {n:3d}  dram {tr / 1e9:8.2f} GB  per launch {tr / n / 1e6:10.1f} MB  "
          f"{tr / v['gpu__time_duration.sum'] / 1e6:6.0f} GB/s while running")
shutil.copy(src, kept)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump({"source": f"{os.path.relpath(kept, root)}: ncu dram__bytes_read.sum + dram__bytes_write.sum over ONE bench step "
                     "(cfg2, 6 cells = 144 problems), averaged per launch of each kernel",
           "bytes_per_launch": out}, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
