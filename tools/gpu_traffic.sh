#!/bin/bash
# DRAM traffic of every launch of the tile kernels over one bench run (feeds profiles/traffic.json),
# plus a --set full capture of k_matern32.   usage: tools/gpu_traffic.sh <tag>
tag=${1:-r01}
out=gpurun_out
# one step = 190 launches of these kernels (48 + 48 + 69 + 23 + 1 + 1); 3 warm-up steps are skipped
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --l2-iters 0 --hbm-points 0 --dba-iters 0 --factored-steps 0 --no-side-configs --no-reference-api"
$CMD > $out/plain_traffic_$tag.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k 'regex:k_chol_update|k_trtri_accum|k_lauum_cov|k_panel_scale|k_diag_block|k_matern32' -s 570 -c 190 \
    --csv --log-file $out/traffic_$tag.csv $CMD > $out/ncu_traffic_$tag.log 2>&1
echo "traffic rc=$?"
$CMD > $out/plain_matern_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_matern32 -s 3 -c 1 -f -o $out/matern_$tag $CMD > $out/ncu_matern_$tag.log 2>&1
echo "matern rc=$?"
