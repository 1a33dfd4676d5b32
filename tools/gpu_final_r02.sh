#!/bin/bash
# Final check of round 2 in one gpurun call: GPU parity tests, smoke, the bench line (both arms), the ncu launch list of
# the step and a --set full capture of the dominant kernel.   usage: tools/gpu_final_r02.sh <tag>
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $out/${tag}_pytest_gpu.log)"
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $out/${tag}_smoke.log)"
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"
# the profiled command is the step alone, so that the launch list's kernel SHARES can be compared with `stages`
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --l2-iters 0 --hbm-points 0 --dba-iters 0 --factored-steps 0 --no-side-configs --no-reference-api"
$CMD > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_chol_update|k_lauum_cov' -s 68 -c 3 -f -o $out/${tag}_tile_kernels $CMD > $out/${tag}_ncu_full.log 2>&1
echo "full rc=$?"
