#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export REPEAT=1
DB_GP_OVERLAP=1 timeout 300 python scripts/exp_overlap.py 2>&1 | tail -1
DB_GP_OVERLAP=1 timeout 300 python scripts/exp_overlap.py 2048 256 2>&1 | tail -1
DB_GP_OVERLAP=1 timeout 300 python scripts/timeline_gp.py 2>/dev/null | grep "mem_kernel\|fused_kernel\|span"
timeout 900 python -m pytest tests/test_gpu_gp.py tests/test_gpu_gpdbn.py -m gpu -x -q 2>&1 | tail -5
