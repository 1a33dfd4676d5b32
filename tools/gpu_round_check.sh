#!/bin/bash
# One gpurun call: GPU parity tests, the bench lines, the ncu launch list and full captures.
# usage: tools/gpu_round_check.sh <tag>
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"
# the profiled command is the step alone (no L2 line, no stand-alone stage measurements), so that the
# launch list's kernel shares can be compared with the bench line's "stages"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --l2-iters 0 --hbm-points 0 --dba-iters 0 --factored-steps 0"
$CMD > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_$tag.csv $CMD > $out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
$CMD > $out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_chol_update -s 68 -c 3 -f -o $out/chol_update_$tag $CMD > $out/ncu_full_$tag.log 2>&1
echo "full chol rc=$?"
$CMD > $out/plain3_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_lauum_cov|k_matern32|k_trtri_accum' -s 22 -c 3 -f -o $out/others_$tag $CMD > $out/ncu_full2_$tag.log 2>&1
echo "full others rc=$?"
