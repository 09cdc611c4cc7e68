#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export REPEAT=1
for ks in 0 1; do
  echo "ksplit $ks"
  DB_GP_KSPLIT=$ks DB_GP_OVERLAP=0 timeout 120 python scripts/exp_overlap.py 2>&1 | tail -1
  DB_GP_KSPLIT=$ks DB_GP_OVERLAP=1 timeout 120 python scripts/exp_overlap.py 2>&1 | tail -1
  DB_GP_KSPLIT=$ks timeout 120 python scripts/exp_overlap.py 8192 784 2>&1 | tail -1
  DB_GP_KSPLIT=$ks timeout 120 python scripts/exp_overlap.py 6144 784 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_gp.py tests/test_gpu_gpdbn.py -m gpu -x -q 2>&1 | tail -4
