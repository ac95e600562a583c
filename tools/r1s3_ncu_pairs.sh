#!/bin/bash
# ncu --set full of the two pair kernels (tools/ncu_target.py, n=65536) + per-SASS-line hot spots.
TAG=${1:-s3}
O=gpurun_out; mkdir -p $O /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_pairs_kernel -s 2 -c 2 -f -o /tmp/ncu/pairs python tools/ncu_target.py fp64 65536 > $O/ncu_pairs_${TAG}.log 2>&1
python tools/ncu_digest.py /tmp/ncu/pairs.ncu-rep > $O/ncu_digest_pairs_${TAG}.txt 2>> $O/ncu_pairs_${TAG}.log
ncu -i /tmp/ncu/pairs.ncu-rep --page source --csv > /tmp/ncu/pairs_src.csv 2>> $O/ncu_pairs_${TAG}.log
python tools/ncu_src.py /tmp/ncu/pairs_src.csv 0 70 > $O/ncu_src_pairs_${TAG}_k0.txt 2>&1
python tools/ncu_src.py /tmp/ncu/pairs_src.csv 1 70 > $O/ncu_src_pairs_${TAG}_k1.txt 2>&1
python tools/ncu_src.py /tmp/ncu/pairs_src.csv 2 70 > $O/ncu_src_pairs_${TAG}_k2.txt 2>&1
python tools/ncu_src.py /tmp/ncu/pairs_src.csv 3 70 > $O/ncu_src_pairs_${TAG}_k3.txt 2>&1
head -50 $O/ncu_digest_pairs_${TAG}.txt
