#!/bin/bash
# The weights tests over all shapes, then the bench at the cfg4 / cfg1 / cfg3 shapes with the final library.
tag=r01r
out=gpurun_out; mkdir -p $out
timeout 200 python -m pytest tests -m gpu -x -q -k "weights" > $out/pytest_weights_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_weights_$tag.log
for c in "cfg4 128" "cfg1 64" "cfg3 6"; do
  set -- $c
  timeout 250 python bench.py --workload $1 --cells-per-step $2 --l2-iters 0 --dba-iters 0 --factored-steps 0 > $out/bench_${1}_$tag.json 2> $out/bench_${1}_$tag.err; echo "$1 rc=$?"
done
