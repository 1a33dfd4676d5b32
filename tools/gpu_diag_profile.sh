#!/bin/bash
tag=${1:-r01}
out=gpurun_out
CMD="python bench.py --workload cfg4 --cells-per-step 16 --steps 1 --warmup 3 --no-cpu-baseline --l2-iters 0 --dba-iters 0 --factored-steps 0"
$CMD > $out/plain_diag_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_diag_block -s 12 -c 2 -f -o $out/diag_$tag $CMD > $out/ncu_diag_$tag.log 2>&1
echo "diag rc=$?"
