#!/bin/bash
# iteration loop for the per-point weight kernels (CRPS / KSD / similarity): parity tests + the hbm_stages block
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "crps or ksd or similarity or nan or perfect_model or weights" > $out/${tag}_tests.log 2>&1; echo "tests rc=$? $(tail -1 $out/${tag}_tests.log)"
timeout 600 python bench.py --steps 2 --warmup 3 --no-side-configs --l2-iters 0 --dba-iters 0 --factored-steps 0 --no-reference-api --no-cpu-baseline --no-member-sharded --no-svgp > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -2 $out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
for k, v in d["hbm_stages"]["kernels"].items():
    print(k, round(v["ms"], 3), "ms", round(v["frac_of_hbm_peak"], 3), v.get("evaluations_per_sec"))
PY
