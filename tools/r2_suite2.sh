#!/bin/bash
# usage: gpurun -- bash tools/r2_suite2.sh TAG "workloads for bench" "workloads for ncu launch lists"
TAG=$1
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/${TAG}_gputest.log
tail -4 $O/${TAG}_gputest.log
for w in $2; do
  extra="--no-cpu"
  if [ "$w" = cfg3_sdgpr ]; then extra=""; fi
  python bench.py --workload $w $extra > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open('$O/${TAG}_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches'],
          'roof', round(d['roofline']['frac'], 4), 'whole', round(d['roofline']['whole_step_frac'], 4))
    print('   ', d['kernel_ms_per_step'])
    if 'parity' in d: print('    parity', json.dumps(d['parity'])[:900])
    if 'secondary' in d: print('    secondary', json.dumps(d['secondary'])[:600])
except Exception as e:
    print('$w', 'FAILED', e)
    print(open('$O/${TAG}_bench_$w.err').read()[-1500:])
PY
done
bash tools/r2_launches.sh $TAG "$3"
