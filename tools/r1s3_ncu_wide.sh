#!/bin/bash
TAG=${1:-w}
O=gpurun_out; mkdir -p $O /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mm_bwd_wide_mma|mm_fwd_wide" -s 2 -c 2 -f -o /tmp/ncu/wide python tools/ncu_wide_target.py 32768 > $O/ncu_wide_${TAG}.log 2>&1
python tools/ncu_digest.py /tmp/ncu/wide.ncu-rep > $O/ncu_digest_wide_${TAG}.txt 2>> $O/ncu_wide_${TAG}.log
ncu -i /tmp/ncu/wide.ncu-rep --page source --csv > /tmp/ncu/wide_src.csv 2>> $O/ncu_wide_${TAG}.log
for k in 0 1 2 3; do python tools/ncu_src.py /tmp/ncu/wide_src.csv $k 45 > $O/ncu_src_wide_${TAG}_k$k.txt 2>&1; done
cat $O/ncu_digest_wide_${TAG}.txt | cut -c1-250
