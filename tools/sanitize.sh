#!/bin/bash
# compute-sanitizer passes over the library's kernels (run on a GPU box through gpurun):
#   memcheck  on smoke() and on the op-level parity tests (every kernel family at small shapes),
#   racecheck on smoke() (shared-memory hazards of one full SDGPR step).
# Each leg is bounded by its own timeout; logs go to gpurun_out/.
out=${1:-gpurun_out}
mkdir -p "$out"
run() {   # name, seconds, tool args..., -- command
    name=$1; secs=$2; shift 2
    echo "== $name" | tee -a "$out/sanitize_summary.txt"
    timeout "$secs" compute-sanitizer --error-exitcode 7 --print-limit 20 --log-file "$out/sanitize_$name.log" "$@" > "$out/sanitize_$name.out" 2>&1
    rc=$?
    echo "rc=$rc" | tee -a "$out/sanitize_summary.txt"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" "$out/sanitize_$name.log" | sort | uniq -c | head -20 | tee -a "$out/sanitize_summary.txt"
    tail -3 "$out/sanitize_$name.out" | tee -a "$out/sanitize_summary.txt"
}
run memcheck_smoke 240 --tool memcheck python -c "import __graft_entry__ as g; g.smoke()"
run memcheck_ops 420 --tool memcheck python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider
run racecheck_smoke 240 --tool racecheck python -c "import __graft_entry__ as g; g.smoke()"
