#!/bin/bash
# A/B of library variants at the cfg2 shape.  usage: bash tools/r1s3_wide_ab.sh TAG default w8 ...
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib=$PWD/geepee_b200/csrc/libgpb_$v.so; fi
  echo "== $v"
  GPB_LIB_PATH=$lib python -m pytest tests/test_gpu_ops.py -x -q -k "mm" 2>&1 | tail -1
  GPB_LIB_PATH=$lib python bench.py --no-cpu --workload cfg2_sgplvm > $O/bench_${TAG}_cfg2_$v.json 2> $O/bench_${TAG}_cfg2_$v.err
  python -c "
import json
d = json.loads(open('$O/bench_${TAG}_cfg2_$v.json').read().strip().splitlines()[-1])
print('cfg2', d['ms_per_step'], d['energy'], d['roofline']['frac'], d['kernel_ms_per_step'])
"
done
