#!/bin/bash
# A/B of the diagonal / product warp split of the small-T kernels (developer tool)
mkdir -p gpurun_out
for w in 2 3 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DBE_SMALL_DIAG_WARPS=$w \
    -o /tmp/libbe_split$w.so bayesian_ensembling_b200/csrc/be_api.cu || exit 1
  BE_B200_LIB=/tmp/libbe_split$w.so python bench.py --workload cfg4 --cells-per-step 256 --no-side-configs --l2-iters 0 --dba-iters 0 --hbm-points 0 \
     --factored-steps 0 --no-reference-api --no-cpu-baseline > gpurun_out/split${w}_bench_cfg4.json 2>/dev/null
  echo "diag warps $w: $(python tools/show_bench.py gpurun_out/split${w}_bench_cfg4.json | head -1)"
done
