#!/bin/bash
# compute-sanitizer over the hot path (race detection / sanitizers).  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
# Logs -> gpurun_out/sanitize_<tool>.log (copy the summaries to profiles/).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "tour fails WITHOUT the sanitizer"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  [ "$tool" = initcheck ] && extra="--track-unused-memory no"
  timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $tool $extra --print-limit 20 \
      python tools/sanitize_cases.py --quick > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
