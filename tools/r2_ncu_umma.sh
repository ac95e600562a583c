#!/bin/bash
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:det_fwd_umma -s 1 -c 1 -f -o /tmp/ncu/umma python tools/ncu_target.py fp32 65536 > $O/${TAG}_ncu_umma.log 2>&1
python tools/ncu_digest.py /tmp/ncu/umma.ncu-rep > $O/${TAG}_ncu_digest_umma.txt 2>> $O/${TAG}_ncu_umma.log
ncu -i /tmp/ncu/umma.ncu-rep --page source --csv > /tmp/ncu/umma_src.csv 2>> $O/${TAG}_ncu_umma.log
python tools/ncu_src.py /tmp/ncu/umma_src.csv 0 45 > $O/${TAG}_ncu_src_umma.txt 2>&1
cat $O/${TAG}_ncu_digest_umma.txt | head -24; head -50 $O/${TAG}_ncu_src_umma.txt | cut -c1-160
