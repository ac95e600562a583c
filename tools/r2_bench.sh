#!/bin/bash
# usage: gpurun -- bash tools/r2_bench.sh TAG [bench.py args]   -> gpurun_out/TAG_bench.json
TAG=$1; shift
O=gpurun_out; mkdir -p $O
python bench.py "$@" > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 6000 $O/${TAG}_bench.json
tail -5 $O/${TAG}_bench.err
