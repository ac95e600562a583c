#!/bin/bash
# ncu --set full of the pair kernels (tools/ncu_target.py, n=65536) + per-SASS-line hot spots.
# usage: r2_ncu_pairs.sh TAG [kernel regex]
TAG=${1:-r2}
RX=${2:-mm_pairs}
O=gpurun_out; mkdir -p $O /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s 2 -c 2 -f -o /tmp/ncu/pairs python tools/ncu_target.py fp64 65536 > $O/${TAG}_ncu_pairs.log 2>&1
python tools/ncu_digest.py /tmp/ncu/pairs.ncu-rep > $O/${TAG}_ncu_digest_pairs.txt 2>> $O/${TAG}_ncu_pairs.log
ncu -i /tmp/ncu/pairs.ncu-rep --page source --csv > /tmp/ncu/pairs_src.csv 2>> $O/${TAG}_ncu_pairs.log
python tools/ncu_src.py /tmp/ncu/pairs_src.csv 0 60 > $O/${TAG}_ncu_src_pairs_k0.txt 2>&1
python tools/ncu_src.py /tmp/ncu/pairs_src.csv 1 60 > $O/${TAG}_ncu_src_pairs_k1.txt 2>&1
cp /tmp/ncu/pairs.ncu-rep $O/${TAG}_pairs.ncu-rep
head -60 $O/${TAG}_ncu_digest_pairs.txt
