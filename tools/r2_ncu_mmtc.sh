#!/bin/bash
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_pairs_tc -s 1 -c 1 -f -o /tmp/ncu/mmtc python tools/mmtc_time.py 65536 2 2 > $O/${TAG}_ncu_mmtc.log 2>&1
python tools/ncu_digest.py /tmp/ncu/mmtc.ncu-rep > $O/${TAG}_ncu_digest_mmtc.txt 2>> $O/${TAG}_ncu_mmtc.log
ncu -i /tmp/ncu/mmtc.ncu-rep --page source --csv > /tmp/ncu/mmtc_src.csv 2>> $O/${TAG}_ncu_mmtc.log
python tools/ncu_src.py /tmp/ncu/mmtc_src.csv 0 45 > $O/${TAG}_ncu_src_mmtc.txt 2>&1
ncu -i /tmp/ncu/mmtc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
for r in rows[2:3]:
    for h, v in zip(hdr, r):
        if any(k in h for k in ('pipe_xu', 'inst_executed_pipe', 'pipe_fma', 'pipe_alu', 'issue_active', 'tmem', 'pipe_tensor', 'mio', 'lsu')) and 'pct' in h:
            print(h, v)
" > $O/${TAG}_ncu_pipes_mmtc.txt 2>&1
cat $O/${TAG}_ncu_digest_mmtc.txt | head -24; head -40 $O/${TAG}_ncu_src_mmtc.txt | cut -c1-160
