#!/bin/bash
# multi-GPU bench line(s) exactly as the driver launches them.  usage: gpurun --gpus N -- bash tools/r1s3_scale.sh TAG N [workload]
TAG=$1; N=$2; WL=${3:-cfg3_sdgpr}
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --workload $WL > $O/scale_${TAG}_${WL}_$N.json 2> $O/scale_${TAG}_${WL}_$N.err
tail -c 1500 $O/scale_${TAG}_${WL}_$N.json; grep -i "warn\|error\|capture" $O/scale_${TAG}_${WL}_$N.err | head -10
