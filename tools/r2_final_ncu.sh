#!/bin/bash
# Round 2 evidence: (a) ncu --set full of every hot kernel at the ncu_target shapes (M=256, n=65536),
# (b) of the four pair-kernel launches of one bench step at the bench's own size (roofline.traffic),
# (c) the launch list of a short default bench run (share of each kernel in a step).
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"det_fwd|det_bwd|det_syrk_mma|mm_pairs|mm_psi1|mm_rows|mm_cols|spd_inverse|tail_gemm|tail_khyper" -s 12 -c 14 -f -o /tmp/ncu/prof_all python tools/ncu_target.py fp64 65536 > $O/${TAG}_ncu_all.log 2>&1
python tools/ncu_digest.py /tmp/ncu/prof_all.ncu-rep > $O/${TAG}_ncu_digest_all.txt 2>> $O/${TAG}_ncu_all.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mm_pairs -s 12 -c 4 -f -o /tmp/ncu/prof_bench_pairs python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > $O/${TAG}_ncu_bench_pairs.log 2>&1
python tools/ncu_digest.py /tmp/ncu/prof_bench_pairs.ncu-rep > $O/${TAG}_ncu_digest_bench_pairs.txt 2>> $O/${TAG}_ncu_bench_pairs.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches_cfg3_sdgpr.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > $O/${TAG}_ncu_cfg3.log 2>&1
python tools/launch_share.py $O/${TAG}_launches_cfg3_sdgpr.csv > $O/${TAG}_launch_share_cfg3.txt
grep -c "== kernel" $O/${TAG}_ncu_digest_all.txt $O/${TAG}_ncu_digest_bench_pairs.txt
head -30 $O/${TAG}_launch_share_cfg3.txt
grep -A3 "== kernel" $O/${TAG}_ncu_digest_bench_pairs.txt | grep -E "kernel|duration" 
