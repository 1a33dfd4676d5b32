#!/bin/bash
# quick iteration loop for the small-T kernels: parity tests, cfg4 bench, optional ncu (developer tool)
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_shapes.py -m gpu -x -q -k "small_t or cfg4 or one_nan" > $out/${tag}_tests.log 2>&1; echo "tests rc=$? $(tail -1 $out/${tag}_tests.log)"
timeout 300 python bench.py --workload cfg4 --cells-per-step 256 --no-side-configs --l2-iters 0 --dba-iters 0 --hbm-points 0 --factored-steps 0 --no-reference-api --no-cpu-baseline > $out/${tag}_bench_cfg4.json 2> $out/${tag}_bench_cfg4.err; echo "bench rc=$?"; tail -2 $out/${tag}_bench_cfg4.err
python tools/show_bench.py $out/${tag}_bench_cfg4.json 2>/dev/null | head -14
if [ "${2:-}" = ncu ]; then
  CMD="python bench.py --workload cfg4 --cells-per-step 64 --steps 1 --warmup 3 --no-cpu-baseline --l2-iters 0 --dba-iters 0 --factored-steps 0 --hbm-points 0 --no-side-configs --no-reference-api"
  $CMD > $out/plain_small_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_small -s 6 -c 2 -f -o $out/small_$tag $CMD > $out/ncu_small_$tag.log 2>&1
  echo "ncu rc=$?"
fi
