#!/bin/bash
# ncu --set full captures of the weights kernel (at the hbm_stages size) and of the DTW DP / backtrack kernels
tag=${1:-r01m}
out=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --l2-iters 0 --dba-iters 0 --factored-steps 0"
$CMD > $out/plain_w_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_loglik_weights_mvn -s 7 -c 1 -f -o $out/weights_$tag $CMD > $out/ncu_w_$tag.log 2>&1
echo "weights rc=$?"
CMD2="python tools/prof_dba.py cfg2 6 3 0"
$CMD2 > $out/plain_dp_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_dtw_dp|k_dtw_backtrack' -s 2 -c 2 -f -o $out/dtw_$tag $CMD2 > $out/ncu_dp_$tag.log 2>&1
echo "dtw rc=$?"
