#!/bin/bash
# tcgen05 fp32 forward: parity tests, ncu capture, fp32 bench lines
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q -k "fp32" 2>&1 | tail -5 | tee $O/${TAG}_umma_tests.log
bash tools/r2_ncu_umma.sh $TAG
timeout 600 python bench.py --prec fp32 --no-secondary --steps 10 --warmup 5 2>$O/${TAG}_bench_fp32.err | tee $O/${TAG}_bench_fp32.json | cut -c1-600
timeout 600 python bench.py --prec fp32 --workload ns_sgpr --no-secondary --steps 20 --warmup 10 2>$O/${TAG}_bench_ns_fp32.err | tee $O/${TAG}_bench_ns_fp32.json | cut -c1-600
