#!/bin/bash
# Round-2 measurement set on one B200: GPU test suite, the default bench line (cfg3 fp64 + cpu_baseline + parity +
# secondary NS block), the reference arm, fp32-psi mode, and the other BASELINE configs at full size.
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${TAG}_gputest.log; tail -2 $O/${TAG}_gputest.log
python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
python bench.py --impl reference > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
python bench.py --prec fp32 --no-secondary > $O/${TAG}_bench_fp32.json 2> $O/${TAG}_bench_fp32.err
python bench.py --prec fp32 --workload ns_sgpr --no-cpu --no-secondary --steps 20 --warmup 10 > $O/${TAG}_bench_ns_sgpr_fp32.json 2> $O/${TAG}_bench_ns_sgpr_fp32.err
python bench.py --prec fp32 --workload cfg4_sgpssm --no-cpu --no-secondary > $O/${TAG}_bench_cfg4_sgpssm_fp32.json 2> $O/${TAG}_bench_cfg4_sgpssm_fp32.err
for w in ns_sgpr cfg1_sgpr cfg2_sgplvm cfg4_sgpssm cfg5_sgpr; do
  python bench.py --workload $w --no-cpu --steps 10 --warmup 10 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('$O/${TAG}_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f.split('${TAG}_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1),
              'roof', round(r.get('frac') or 0, 4), 'whole', round(r.get('whole_step_frac') or 0, 4),
              'pairs', {k: round(v['frac'], 4) for k, v in (r.get('pair_kernels') or {}).items()},
              'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('worst_grad_rel'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
