#!/bin/bash
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 scripts/scale_probe.py 2>&1 | grep MODE; }
MODE=plain run
MODE=pipe COMM_SMS=16 NCCL_MAX_CTAS=16 run
MODE=pipe COMM_SMS=32 NCCL_MAX_CTAS=32 run
MODE=pipe COMM_SMS=0 run
MODE=pipe COMM_SMS=16 DW_CHUNKS=2 NCCL_MAX_CTAS=16 run
