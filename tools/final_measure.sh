#!/bin/bash
# End-of-round measurement set (1 GPU): default bench + reference arm, per-workload lines, ncu launch
# list of one bench run, ncu --set full of every hot kernel (tools/ncu_target.py) and of the dominant
# kernel at the bench's own launch size (roofline.traffic).
TAG=${1:-final}
python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
for wl in ns_sgpr cfg2_sgplvm cfg4_sgpssm cfg5_sgpr cfg1_sgpr; do
  python bench.py --no-cpu --workload $wl > gpurun_out/bench_${TAG}_$wl.json 2> gpurun_out/bench_${TAG}_$wl.err
done
python bench.py --no-cpu --prec fp32 > gpurun_out/bench_${TAG}_fp32.json 2> gpurun_out/bench_${TAG}_fp32.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("_${TAG}_")[1][:-5], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_ncu_${TAG}.log 2>&1
