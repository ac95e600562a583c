#!/bin/bash
# One GPU round trip: parity tests, per-kernel timings, bench lines, one ncu capture of the pair kernels.
# usage: gpurun -- bash tools/gpu_check.sh TAG
TAG=${1:-dev}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kbench.py quick > gpurun_out/kbench_$TAG.log 2>&1
python bench.py --no-cpu > gpurun_out/bench_${TAG}_fp64.json 2> gpurun_out/bench_$TAG.err
python bench.py --no-cpu --prec fp32 > gpurun_out/bench_${TAG}_fp32.json 2>> gpurun_out/bench_$TAG.err
python bench.py --no-cpu --workload ns_sgpr > gpurun_out/bench_${TAG}_ns.json 2>> gpurun_out/bench_$TAG.err
python - <<PY
import json
for f in ["fp64", "fp32", "ns"]:
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["kernel_ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_$TAG.err
grep '"mm_\|det_' gpurun_out/kbench_$TAG.log | cut -c1-200
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mm_pairs_kernel|det_fwd|det_syrk_" -s 2 -c 4 -f \
    -o gpurun_out/prof_mm_pairs_$TAG python tools/ncu_target.py fp64 65536 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
fi
