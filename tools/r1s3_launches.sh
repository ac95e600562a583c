#!/bin/bash
TAG=${1:-v27}
O=gpurun_out; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/bench_ncu_$TAG.log 2>&1
python tools/launch_share.py $O/launches_$TAG.csv | tee $O/launch_share_$TAG.txt
ls -la $O/launches_$TAG.csv
