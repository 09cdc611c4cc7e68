#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export REPEAT=1
timeout 120 python scripts/time_gp.py 2>&1 | tail -3
DB_GP_OVERLAP=0 timeout 120 python scripts/exp_overlap.py 2>&1 | tail -1
DB_GP_OVERLAP=1 timeout 120 python scripts/exp_overlap.py 2>&1 | tail -1
timeout 120 python scripts/exp_overlap.py 8192 784 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_gp.py tests/test_gpu_gpdbn.py -m gpu -x -q 2>&1 | tail -4
