#!/bin/bash
# A/B of library variants on the pair kernels (tools/kbench.py mm) + the default bench line.
# usage: gpurun -- bash tools/r1s3_ab.sh TAG variant1 variant2 ...   ("default" = the product build)
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib=$PWD/geepee_b200/csrc/libgpb_$v.so; fi
  GPB_LIB_PATH=$lib python tools/kbench.py mm 2>&1 | grep '"mm_\|fma_peak' | cut -c1-260 > $O/kbench_mm_${TAG}_$v.log
  echo "== $v"; python -c "
import sys, json
for l in open('$O/kbench_mm_${TAG}_$v.log'):
    d = json.loads(l)
    if d['kind'] != 'fma_peak': print(d['kind'], d['M'], d['Q'], d['Do'], d['ms_min'])
"
  GPB_LIB_PATH=$lib python bench.py --no-cpu > $O/bench_${TAG}_cfg3_$v.json 2> $O/bench_${TAG}_cfg3_$v.err
  python -c "
import json
d = json.loads(open('$O/bench_${TAG}_cfg3_$v.json').read().strip().splitlines()[-1])
print('cfg3', d['ms_per_step'], d['energy'], d['roofline']['frac'], d['kernel_ms_per_step'])
"
done
