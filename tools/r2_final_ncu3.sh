#!/bin/bash
# after the slot reducer: launch list of the default step + full capture of the tensor-exponent forward kernel
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O /tmp/ncu
bash tools/r2_launches.sh $TAG "cfg3_sdgpr" | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_pairsx -s 2 -c 2 -f -o /tmp/ncu/px python tools/ncu_target.py fp64 65536 > $O/${TAG}_ncu_px.log 2>&1
python tools/ncu_digest.py /tmp/ncu/px.ncu-rep > $O/${TAG}_ncu_digest_pairsx.txt 2>> $O/${TAG}_ncu_px.log
head -40 $O/${TAG}_ncu_digest_pairsx.txt
