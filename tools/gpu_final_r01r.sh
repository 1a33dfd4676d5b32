#!/bin/bash
# Final check of the round: GPU parity tests, the bench line, the reference arm, the launch list, and a full
# ncu capture of the table-exp weights kernel at the hbm_stages size (tools/prof_weights.py's operands).
tag=r01r
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"
python tools/prof_weights.py > $out/plain_w_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_loglik_weights_mvn -s 2 -c 1 -f -o $out/weights_$tag python tools/prof_weights.py > $out/ncu_w_$tag.log 2>&1
echo "weights ncu rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --l2-iters 0 --hbm-points 0 --dba-iters 0 --factored-steps 0"
$CMD > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_$tag.csv $CMD > $out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
