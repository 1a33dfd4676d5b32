#!/bin/bash
# iteration loop for the SVGP stage: parity tests, the svgp_stage bench block, the launch list of three steps
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_svgp.py -m gpu -x -q > $out/${tag}_svgp_tests.log 2>&1; echo "tests rc=$? $(tail -1 $out/${tag}_svgp_tests.log)"
timeout 600 python bench.py --steps 1 --warmup 3 --no-side-configs --l2-iters 0 --dba-iters 0 --factored-steps 0 --hbm-points 0 --no-reference-api --no-member-sharded > $out/${tag}_svgp_bench.json 2> $out/${tag}_svgp_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$out/${tag}_svgp_bench.json')); print({k: v for k, v in d['svgp_stage'].items() if k != 'note'})"
python tools/prof_svgp.py 3 > $out/${tag}_svgp_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_svgp_launches.csv python tools/prof_svgp.py 3 > $out/${tag}_svgp_ncu.log 2>&1; echo "launch list rc=$?"
