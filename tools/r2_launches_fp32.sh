#!/bin/bash
TAG=$1
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/${TAG}_launches_cfg3_fp32.csv \
    python bench.py --prec fp32 --steps 1 --warmup 3 --no-cpu --no-secondary > $O/${TAG}_ncu_cfg3_fp32.log 2>&1
python tools/launch_share.py $O/${TAG}_launches_cfg3_fp32.csv | head -40
