#!/bin/bash
# end-of-round profiler evidence: launch lists of the default (fp64) and fp32 steps, full capture of the tcgen05 rank update
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O /tmp/ncu
bash tools/r2_launches.sh $TAG "cfg3_sdgpr" | head -30
bash tools/r2_launches_fp32.sh $TAG | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:det_syrk_umma -s 1 -c 1 -f -o /tmp/ncu/syrk python tools/ncu_target.py fp32 65536 > $O/${TAG}_ncu_syrk.log 2>&1
python tools/ncu_digest.py /tmp/ncu/syrk.ncu-rep > $O/${TAG}_ncu_digest_syrk_umma.txt 2>> $O/${TAG}_ncu_syrk.log
ncu -i /tmp/ncu/syrk.ncu-rep --page source --csv > /tmp/ncu/syrk_src.csv 2>> $O/${TAG}_ncu_syrk.log
python tools/ncu_src.py /tmp/ncu/syrk_src.csv 0 30 >> $O/${TAG}_ncu_digest_syrk_umma.txt 2>&1
head -22 $O/${TAG}_ncu_digest_syrk_umma.txt
