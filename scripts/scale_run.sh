#!/bin/bash
# bench at N GPUs (torchrun for N>1); usage: scale_run.sh N
N=$1
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3
fi
