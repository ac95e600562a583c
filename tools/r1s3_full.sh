#!/bin/bash
# Full GPU check of the session: every GPU test, default bench (+ reference arm), the other workloads.
TAG=${1:-s3}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err
for wl in ns_sgpr cfg2_sgplvm cfg4_sgpssm cfg5_sgpr cfg1_sgpr; do
  python bench.py --no-cpu --workload $wl > $O/bench_${TAG}_$wl.json 2> $O/bench_${TAG}_$wl.err
done
python bench.py --no-cpu --prec fp32 > $O/bench_${TAG}_fp32.json 2> $O/bench_${TAG}_fp32.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("_${TAG}_")[1][:-5], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("gpu_launches"))
    except Exception as e:
        print(f, "failed", e)
PY
