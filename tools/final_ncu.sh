#!/bin/bash
# ncu --set full of every hot kernel (tools/ncu_target.py, M=256, n=65536) and of the dominant kernel at
# the bench's own launch size (roofline.traffic).  The reports are digested ON the GPU box
# (tools/ncu_digest.py) and removed: gpurun returns at most 64 MiB.
TAG=${1:-final}
mkdir -p /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"det_fwd|det_bwd|det_syrk_mma|mm_pairs|mm_psi1|mm_rows|mm_cols|spd_inverse" -s 9 -c 11 -f -o /tmp/ncu/prof_all python tools/ncu_target.py fp64 65536 > gpurun_out/ncu_all_${TAG}.log 2>&1
python tools/ncu_digest.py /tmp/ncu/prof_all.ncu-rep > gpurun_out/ncu_digest_all_${TAG}.txt 2>> gpurun_out/ncu_all_${TAG}.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mm_pairs_kernel -s 14 -c 2 -f -o /tmp/ncu/prof_bench_pairs python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench_pairs_${TAG}.log 2>&1
python tools/ncu_digest.py /tmp/ncu/prof_bench_pairs.ncu-rep > gpurun_out/ncu_digest_bench_pairs_${TAG}.txt 2>> gpurun_out/ncu_bench_pairs_${TAG}.log
grep -c "== kernel" gpurun_out/ncu_digest_all_${TAG}.txt gpurun_out/ncu_digest_bench_pairs_${TAG}.txt
ls -la gpurun_out
