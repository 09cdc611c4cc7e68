#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in 1.5 1.75 2.0 2.5 3.0; do
  echo "ratio $r"; DB_GP_UPD_RATIO=$r timeout 300 python scripts/exp_overlap.py 2>&1 | tail -1
done
for r in 1.5 2.0 2.5; do
  echo "ratio $r N=2048/6144"; DB_GP_UPD_RATIO=$r timeout 300 python scripts/exp_overlap.py 2048 256 2>&1 | tail -1; DB_GP_UPD_RATIO=$r timeout 300 python scripts/exp_overlap.py 6144 784 2>&1 | tail -1
done
