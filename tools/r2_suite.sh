#!/bin/bash
# usage: gpurun -- bash tools/r2_suite.sh TAG "workload1 workload2 ..."   (tests, then one bench line per workload)
TAG=$1; shift
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/${TAG}_gputest.log
tail -6 $O/${TAG}_gputest.log
for w in $1; do
  extra="--no-cpu"
  if [ "$w" = cfg3_sdgpr ]; then extra=""; fi
  python bench.py --workload $w $extra > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open('$O/${TAG}_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches'],
          'roof', round(d['roofline']['frac'], 4), 'whole', round(d['roofline']['whole_step_frac'], 4), d.get('parity'))
    print('   ', d['kernel_ms_per_step'])
except Exception as e:
    print('$w', 'FAILED', e)
    print(open('$O/${TAG}_bench_$w.err').read()[-1500:])
PY
done
