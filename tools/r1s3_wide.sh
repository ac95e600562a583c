#!/bin/bash
# wide-layer (cfg2) check: ops + model parity on the GPU, kernel timing at the cfg2 shape, bench line
TAG=${1:-w1}
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -x -q 2>&1 | tail -4
python - <<'PY'
import sys, json, torch, numpy as np
sys.path.insert(0, '.')
from geepee_b200 import ops
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).to(dev)
n, M, Q, Do = 100000, 128, 5, 50
mx, z = rnd(n, Q), rnd(M, Q)
vx = (0.1 + torch.rand(n, Q, dtype=torch.float64)).to(dev)
ls, sf = torch.full((Q,), 0.3, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
A, B = rnd(Do, M), (0.01 * rnd(Do, M, M)).contiguous()
dm, dv = rnd(n, Do), rnd(n, Do)
mo, vo, va, p1 = ops.mm_fwd(ops.F64, mx, vx, z, ls, sf, A, B)
ops.profile_enable(True)
for it in range(3):
    ops.mm_bwd(ops.F64, mx, vx, z, ls, sf, A, B, dm, dv, mo, va, p1)
torch.cuda.synchronize()
print('profile', {k: v for k, v in ops.profile_collect().items() if v[1]})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(3):
    ops.mm_bwd(ops.F64, mx, vx, z, ls, sf, A, B, dm, dv, mo, va, p1)
e1.record(); torch.cuda.synchronize()
print('mm_bwd cfg2 shape ms', e0.elapsed_time(e1) / 3)
PY
python bench.py --no-cpu --workload cfg2_sgplvm > $O/bench_${TAG}_cfg2.json 2> $O/bench_${TAG}_cfg2.err
python -c "
import json
d = json.loads(open('$O/bench_${TAG}_cfg2.json').read().strip().splitlines()[-1])
print('cfg2', d['ms_per_step'], d['value'], d['energy'], d['roofline']['frac'], d['kernel_ms_per_step'])
"
