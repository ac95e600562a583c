#!/bin/bash
# multi-GPU bench line(s) exactly as the driver launches them.  usage: gpurun --gpus N -- bash tools/r2_scale.sh TAG N [workload]
TAG=$1; N=$2; WL=${3:-cfg3_sdgpr}
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $WL > $O/${TAG}_scale_${WL}_$N.json 2> $O/${TAG}_scale_${WL}_$N.err
python - <<PY
import json
try:
    d = json.loads(open('$O/${TAG}_scale_${WL}_$N.json').read().strip().splitlines()[-1])
    print('$WL', 'gpus', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['ms_per_step'], 3), d['kernel_ms_per_step'])
except Exception as e:
    print('FAILED', e); print(open('$O/${TAG}_scale_${WL}_$N.err').read()[-2500:])
PY
grep -i "warn\|error\|capture" $O/${TAG}_scale_${WL}_$N.err | head -10
