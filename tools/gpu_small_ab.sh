#!/bin/bash
# A/B of several builds of the library on the cfg4 step (developer tool).  usage: gpu_small_ab.sh <tag> [lib-suffix ...]
# "base" = the in-tree libbe_b200.so, any other word w = bayesian_ensembling_b200/libbe_b200_<w>.so; two rounds.
tag=${1:-x}; shift
libs=${@:-base varB}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_shapes.py -m gpu -x -q -k "small_t or cfg4 or one_nan" > $out/${tag}_tests.log 2>&1; echo "tests rc=$? $(tail -1 $out/${tag}_tests.log)"
for round in 1 2; do
for v in $libs; do
  if [ $v = base ]; then unset BE_B200_LIB; else export BE_B200_LIB=$PWD/bayesian_ensembling_b200/libbe_b200_$v.so; fi
  timeout 300 python bench.py --workload cfg4 --cells-per-step 256 --no-side-configs --l2-iters 0 --dba-iters 0 --hbm-points 0 --factored-steps 0 --no-reference-api --no-cpu-baseline --no-svgp > $out/${tag}_bench_cfg4_$v.json 2> $out/${tag}_bench_cfg4_$v.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_bench_cfg4_$v.json"))
print("$v", round(d["value"], 1), "cells/s", {k: round(s["ms_per_step"], 3) for k, s in d["stages"].items() if k.startswith("k_small")})
PY
done
done
