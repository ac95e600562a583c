#!/bin/bash
# ncu launch lists (gpu__time_duration per launch) of short bench runs: usage r2_launches.sh TAG "wl1 wl2"
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for w in $1; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 3 --no-cpu --no-secondary > $O/${TAG}_ncu_$w.log 2>&1
  python tools/launch_share.py $O/${TAG}_launches_$w.csv | head -40
done
