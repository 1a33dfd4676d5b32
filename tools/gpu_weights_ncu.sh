#!/bin/bash
# ncu --set full of the CRPS / KSD / similarity weight kernels at the hbm_stages size
tag=${1:-x}
out=gpurun_out
CMD="python tools/prof_weights_next.py"
$CMD > $out/plain_wn_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_crps_weights|k_ksd_weights|k_similarity_pointwise' -s 3 -c 3 -f -o $out/wnext_$tag $CMD > $out/ncu_wn_$tag.log 2>&1
echo "ncu rc=$?"
