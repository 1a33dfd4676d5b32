#!/bin/bash
# iteration loop for the L2 training loop (be_vgp_fit): parity tests, ms per iteration at cfg2 (bench) and at T = 251
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q -k "vgp or l2 or sqrtm or w2 or fullcov" > $out/${tag}_vgp_tests.log 2>&1; echo "tests rc=$? $(tail -1 $out/${tag}_vgp_tests.log)"
python tools/prof_vgp_small.py 16 5; python tools/prof_vgp_small.py 64 5
timeout 600 python bench.py --steps 1 --warmup 3 --no-side-configs --dba-iters 0 --factored-steps 0 --hbm-points 0 --no-reference-api --no-member-sharded --no-svgp --no-cpu-baseline > $out/${tag}_vgp_bench.json 2> $out/${tag}_vgp_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$out/${tag}_vgp_bench.json')); print({k: v for k, v in d['l2_training_loop'].items() if k not in ('note', 'cpu')})"
