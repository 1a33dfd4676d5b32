"""Prints the one-line summary of a tools/prof_dba.py result (developer tool)."""
import json, sys
for f in sys.argv[1:]:
    d = json.load(open(f))
    print(d["cfg"], "T", d["T"], "ms", round(d["ms"], 1), "it", d["n_iter_mean"],
          {k: (round(v["ms"], 1), round(v.get("gcells_per_s", 0))) for k, v in d["kernels"].items()})
