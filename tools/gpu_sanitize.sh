#!/bin/bash
# Sanitizer tour of the hot path (race detection / sanitizers).  Run on a GPU box:
#   gpurun --timeout 900 -- 'bash tools/gpu_sanitize.sh'
# compute-sanitizer is CLOSED on this pool (profiles/r02a_compute_sanitizer_closed.log: "runs under it have left
# GPUs needing a reset"), so the default is what the pool's message asks for instead: the tour of every kernel
# family at small sizes, checked against the CPU oracle, (1) plain, (2) with the workspace poisoned with NaN
# before every call (BE_B200_POISON_WORKSPACE: a read of unwritten workspace turns the result NaN -- the
# initcheck substitute), (3) repeated, results compared bit for bit (a shared-memory race shows up as run-to-run
# differences).  BE_TRY_SANITIZER=1 adds the compute-sanitizer passes on a pool where it is open.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$? $(tail -1 gpurun_out/sanitize_plain.log)"
BE_B200_POISON_WORKSPACE=1 python tools/sanitize_cases.py > gpurun_out/sanitize_poisoned_workspace.log 2>&1
echo "poisoned-workspace rc=$? $(tail -1 gpurun_out/sanitize_poisoned_workspace.log)"
python tools/sanitize_cases.py --repeat 5 > gpurun_out/sanitize_repeat.log 2>&1; echo "repeat rc=$? $(tail -1 gpurun_out/sanitize_repeat.log)"
if [ "${BE_TRY_SANITIZER:-0}" = 1 ]; then
  for tool in memcheck racecheck synccheck initcheck; do
    extra=""
    [ "$tool" = racecheck ] && extra="--racecheck-report all"
    [ "$tool" = initcheck ] && extra="--track-unused-memory no"
    timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $tool $extra --print-limit 20 \
        python tools/sanitize_cases.py --quick > gpurun_out/sanitize_$tool.log 2>&1
    echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  done
fi
