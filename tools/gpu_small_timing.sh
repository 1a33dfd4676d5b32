#!/bin/bash
# builds a -DBE_SMALL_TIMING copy of the library and prints the per-phase cycle table (developer tool)
tag=${1:-x}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DBE_SMALL_TIMING \
  -o /tmp/libbe_b200_timing.so bayesian_ensembling_b200/csrc/be_api.cu || exit 1
BE_B200_LIB=/tmp/libbe_b200_timing.so python tools/prof_small_timing.py 256 > gpurun_out/${tag}_small_timing.txt 2>&1
cat gpurun_out/${tag}_small_timing.txt
