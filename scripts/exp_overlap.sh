#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export REPEAT=1
for n in "1544 64" "2568 128" "3072 64" "4104 256"; do
  DB_POTRF_TAIL_ROWS=0 DB_GP_OVERLAP=0 timeout 120 python scripts/exp_overlap.py $n 2>&1 | tail -1
  DB_GP_OVERLAP=1 timeout 120 python scripts/exp_overlap.py $n 2>&1 | tail -1
  DB_GP_OVERLAP=0 timeout 120 python scripts/exp_overlap.py $n 2>&1 | grep -v "^replay" | tail -1
done
timeout 900 python -m pytest tests/test_gpu_gp.py tests/test_gpu_gpdbn.py tests/test_gpu_gp_overlap.py -m gpu -x -q 2>&1 | tail -3
