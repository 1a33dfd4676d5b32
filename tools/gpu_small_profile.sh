#!/bin/bash
# ncu --set full of the two small-T member kernels at the cfg4 shape (developer tool)
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
CMD="python bench.py --workload cfg4 --cells-per-step 64 --steps 1 --warmup 3 --no-cpu-baseline --l2-iters 0 --dba-iters 0 --factored-steps 0 --hbm-points 0 --no-side-configs --no-reference-api"
$CMD > $out/plain_small_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_small -s 6 -c 2 -f -o $out/small_$tag $CMD > $out/ncu_small_$tag.log 2>&1
echo "small rc=$?"
(cd tools && nvcc -O3 -arch=sm_100a -o ubench_lat ubench_lat.cu 2>/dev/null); ./tools/ubench_lat > $out/ubench_lat_$tag.txt 2>&1; cat $out/ubench_lat_$tag.txt
