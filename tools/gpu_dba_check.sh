mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dba.py -x -q > gpurun_out/pytest_dba.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_dba.log
timeout 300 python tools/prof_dba.py cfg2 6 50 1 > gpurun_out/prof_dba_cfg2.json 2> gpurun_out/prof_dba_cfg2.err; echo rc=$?; cat gpurun_out/prof_dba_cfg2.json; tail -3 gpurun_out/prof_dba_cfg2.err
timeout 300 python tools/prof_dba.py cfg4 128 50 1 > gpurun_out/prof_dba_cfg4.json 2> gpurun_out/prof_dba_cfg4.err; echo rc=$?; cat gpurun_out/prof_dba_cfg4.json; tail -3 gpurun_out/prof_dba_cfg4.err
