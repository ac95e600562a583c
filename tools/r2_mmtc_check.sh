#!/bin/bash
# tcgen05 fp32 pair forward: parity tests and fp32 bench lines
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q -k "fp32" 2>&1 | tail -15 | tee $O/${TAG}_mmtc_tests.log
timeout 600 python bench.py --prec fp32 --no-secondary --steps 10 --warmup 5 2>$O/${TAG}_bench_fp32.err | tee $O/${TAG}_bench_fp32.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms_per_step'], d['parity']['ok'], d['parity']['worst_grad_rel'], d['parity']['energy_rel'])"
