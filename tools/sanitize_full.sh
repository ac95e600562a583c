#!/bin/bash
# Second, wider compute-sanitizer pass (after tools/sanitize.sh): memcheck over the WHOLE -m gpu suite, racecheck
# over the op-level tests (every kernel family), synccheck + initcheck over smoke().
out=${1:-gpurun_out}
mkdir -p "$out"
run() {
    name=$1; secs=$2; shift 2
    echo "== $name" | tee -a "$out/sanitize_full_summary.txt"
    timeout "$secs" compute-sanitizer --error-exitcode 7 --print-limit 20 --log-file "$out/sanitize_$name.log" "$@" > "$out/sanitize_$name.out" 2>&1
    rc=$?
    echo "rc=$rc" | tee -a "$out/sanitize_full_summary.txt"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|Uninitialized|Barrier error" "$out/sanitize_$name.log" | sort | uniq -c | head -20 | tee -a "$out/sanitize_full_summary.txt"
    tail -3 "$out/sanitize_$name.out" | tee -a "$out/sanitize_full_summary.txt"
}
run memcheck_all 600 --tool memcheck python -m pytest tests -q -m gpu -p no:cacheprovider
run racecheck_ops 600 --tool racecheck python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider
run synccheck_smoke 240 --tool synccheck python -c "import __graft_entry__ as g; g.smoke()"
run initcheck_smoke 240 --tool initcheck python -c "import __graft_entry__ as g; g.smoke()"
