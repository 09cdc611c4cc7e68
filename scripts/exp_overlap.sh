#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export REPEAT=1
for n in "5000 784" "6144 784" "2048 256"; do
  for m in 0 1; do DB_GP_OVERLAP=$m timeout 300 python scripts/exp_overlap.py $n 2>&1 | tail -1; done
done
DB_GP_OVERLAP=1 timeout 300 python scripts/timeline_gp.py 2>/dev/null > gpurun_out/timeline_m1.txt
N_IT=60 timeout 300 python scripts/gpdbn_once.py 2>&1 | tail -1
timeout 300 python scripts/time_gp.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_gp.py tests/test_gpu_gpdbn.py -m gpu -x -q 2>&1 | tail -5
