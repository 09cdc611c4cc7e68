#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export REPEAT=1
for n in "5000 784" "8192 784" "6144 784" "2048 256" "1536 64"; do
  for m in 0 1; do DB_GP_OVERLAP=$m timeout 300 python scripts/exp_overlap.py $n 2>&1 | tail -2; done
done
DB_GP_OVERLAP=1 timeout 300 python scripts/timeline_gp.py 2>/dev/null > gpurun_out/timeline_m1.txt
DB_GP_OVERLAP=1 N=8192 timeout 300 python scripts/timeline_gp.py 2>/dev/null > gpurun_out/timeline_m1_8192.txt
