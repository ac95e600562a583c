#!/bin/bash
# Session-3 GPU round trip: graph-replayed tails (tests + bench A/B), exp-polynomial variants of the
# pair kernels (kbench A/B), conversion-throughput probe.   usage: gpurun -- bash tools/r1s3_check.sh TAG
TAG=${1:-s3}
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_graphs.py -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_graphs.py 2>&1 | tail -5
./tools/probe/dp_probe 2>&1 | tail -8 > $O/dp_probe_cvt_$TAG.txt; cat $O/dp_probe_cvt_$TAG.txt
for v in "" deg3 deg2; do
  if [ -z "$v" ]; then lib=""; else lib=$PWD/geepee_b200/csrc/libgpb_$v.so; fi
  GPB_LIB_PATH=$lib python tools/kbench.py mm 2>&1 | grep '"mm_\|fma_peak' | cut -c1-260 > $O/kbench_mm_${TAG}_${v:-deg4}.log
  echo "== ${v:-deg4}"; cat $O/kbench_mm_${TAG}_${v:-deg4}.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['kind'] != 'fma_peak': print(d['kind'], d['M'], d['Q'], d['Do'], d['ms_min'])
"
done
for wl in ns_sgpr cfg1_sgpr; do
  for g in 1 0; do
    GPB_TAIL_GRAPHS=$g python bench.py --no-cpu --workload $wl > $O/bench_${TAG}_${wl}_g$g.json 2> $O/bench_${TAG}_${wl}_g$g.err
  done
done
GPB_TAIL_GRAPHS=1 python bench.py --no-cpu > $O/bench_${TAG}_cfg3_g1.json 2> $O/bench_${TAG}_cfg3_g1.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("_${TAG}_")[1][:-5], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("ms_per_step"), d.get("gpu_launches"), d.get("energy"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 $O/bench_${TAG}_*.err | tail -20
