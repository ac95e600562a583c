#!/bin/bash
# Round 2: GPU tests + A/B of library variants on the pair kernels (tools/kbench.py mm) + bench line.
# usage: gpurun -- bash tools/r2_ab.sh TAG [test|notest] variant1 variant2 ...
#   "default" = the product build, other names = tools/variants/libv_<name>.so
TAG=$1; shift
DOTEST=$1; shift
O=gpurun_out; mkdir -p $O
if [ "$DOTEST" = test ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/${TAG}_gputest.log
  tail -5 $O/${TAG}_gputest.log
fi
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib=$PWD/tools/variants/libv_$v.so; fi
  GPB_LIB_PATH=$lib python tools/kbench.py mm 2>&1 | grep '"mm_\|fma_peak' | cut -c1-260 > $O/${TAG}_kbench_mm_$v.log
  echo "== $v"; python -c "
import sys, json
for l in open('$O/${TAG}_kbench_mm_$v.log'):
    d = json.loads(l)
    if d['kind'] != 'fma_peak': print(d['kind'], d['M'], d['Q'], d['Do'], d['ms_min'])
"
done
