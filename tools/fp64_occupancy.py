#!/usr/bin/env python
"""How many warps does the FP64 pipe need?  fma_peak (8 independent chains per thread) at
1..8 resident 256-thread blocks per SM (2..16 warps per scheduler)."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops
dev = torch.device('cuda:0')
for prec, name, iters in [(ops.F64, 'fp64', 20000), (ops.F32, 'fp32', 40000)]:
    for b in (1, 2, 3, 4, 8):
        fl = ops.fma_peak(prec, iters, dev, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fl = ops.fma_peak(prec, iters, dev, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('%s blocks/SM=%d (warps/scheduler=%d): %.2f TFLOP/s' % (name, b, 2 * b, fl / best / 1e9))
