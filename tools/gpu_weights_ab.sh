#!/bin/bash
# One gpurun call: the table-exp weights kernel against the library-exp one (accuracy, special cases, time),
# then the GPU tests that touch the weights .
out=gpurun_out; mkdir -p $out
python tools/prof_weights_ab.py > $out/wab_new.json 2> $out/wab_new.err; echo "new rc=$?"
BE_WEIGHTS_LIBEXP=1 python tools/prof_weights_ab.py > $out/wab_old.json 2> $out/wab_old.err; echo "old rc=$?"
cat $out/wab_new.json $out/wab_old.json; tail -3 $out/wab_new.err
timeout 250 python -m pytest tests -m gpu -x -q -k "weights or grid or pipeline or perfect or sharded or cell" > $out/pytest_wab.log 2>&1; echo "pytest rc=$?"
tail -4 $out/pytest_wab.log
