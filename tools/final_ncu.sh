#!/bin/bash
# ncu --set full of every hot kernel (tools/ncu_target.py, M=256, n=65536) and of the dominant kernel at
# the bench's own launch size (roofline.traffic).  Reports are kept small (gpurun returns <= 64 MiB).
TAG=${1:-final}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"det_fwd|det_bwd|det_syrk_mma|mm_pairs|mm_psi1|mm_rows|mm_cols|spd_inverse" -s 9 -c 11 -f -o gpurun_out/prof_all_${TAG} python tools/ncu_target.py fp64 65536 > gpurun_out/ncu_all_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_all_${TAG}.log
timeout 900 ncu --set full --clock-control none -k regex:mm_pairs_kernel -s 14 -c 2 -f -o gpurun_out/prof_bench_pairs_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench_pairs_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_bench_pairs_${TAG}.log
ls -la gpurun_out
