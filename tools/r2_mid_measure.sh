#!/bin/bash
# mid-round check: GPU suite, default bench without the CPU leg, fp32 lines
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/${TAG}_gputest.log; tail -2 $O/${TAG}_gputest.log
timeout 600 python bench.py --no-cpu > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
timeout 600 python bench.py --prec fp32 --no-cpu --no-secondary > $O/${TAG}_bench_fp32.json 2> $O/${TAG}_bench_fp32.err
for w in ns_sgpr cfg1_sgpr cfg4_sgpssm; do
  timeout 600 python bench.py --workload $w --no-cpu --no-secondary > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
timeout 600 python bench.py --workload ns_sgpr --prec fp32 --no-cpu --no-secondary > $O/${TAG}_bench_ns_sgpr_fp32.json 2> $O/${TAG}_bench_ns_sgpr_fp32.err
timeout 600 python bench.py --workload cfg4_sgpssm --prec fp32 --no-cpu --no-secondary > $O/${TAG}_bench_cfg4_sgpssm_fp32.json 2> $O/${TAG}_bench_cfg4_sgpssm_fp32.err
python - <<PY
import json, glob
for f in sorted(glob.glob('$O/${TAG}_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f.split('${TAG}_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 3), 'roof', round(r.get('frac') or 0, 4),
              'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('worst_grad_rel'), d.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
