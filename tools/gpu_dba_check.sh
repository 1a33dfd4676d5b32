#!/bin/bash
# DBA parity tests + timing at BASELINE shapes (one gpurun call). usage: tools/gpu_dba_check.sh [tag] [skip_tests]
tag=${1:-dba}
mkdir -p gpurun_out
if [ -z "$2" ]; then
timeout 900 python -m pytest tests/test_gpu_dba.py -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
fi
for c in "cfg2 6" "cfg3 6" "cfg4 128" "cfg1 64"; do
set -- $c
timeout 300 python tools/prof_dba.py $1 $2 50 1 > gpurun_out/prof_${tag}_$1.json 2> gpurun_out/prof_${tag}_$1.err; echo "$1 rc=$?"; cat gpurun_out/prof_${tag}_$1.json; tail -2 gpurun_out/prof_${tag}_$1.err
done
