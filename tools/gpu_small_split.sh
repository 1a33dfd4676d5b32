#!/bin/bash
# A/B of build-time variants of the small-T kernels (developer tool): usage gpu_small_split.sh "<-D flags A>" "<-D flags B>" ...
mkdir -p gpurun_out
i=0
for flags in "$@"; do
  i=$((i+1))
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC $flags \
    -o /tmp/libbe_variant$i.so bayesian_ensembling_b200/csrc/be_api.cu || exit 1
  BE_B200_LIB=/tmp/libbe_variant$i.so timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k small_t > gpurun_out/variant${i}_tests.log 2>&1
  echo "variant $i [$flags] tests: $(tail -1 gpurun_out/variant${i}_tests.log)"
  BE_B200_LIB=/tmp/libbe_variant$i.so python bench.py --workload cfg4 --cells-per-step 256 --no-side-configs --l2-iters 0 --dba-iters 0 --hbm-points 0 \
     --factored-steps 0 --no-reference-api --no-cpu-baseline > gpurun_out/variant${i}_bench_cfg4.json 2>/dev/null
  echo "variant $i [$flags]: $(python tools/show_bench.py gpurun_out/variant${i}_bench_cfg4.json 2>/dev/null | head -1 | cut -c1-60)"
done
