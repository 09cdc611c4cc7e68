#!/bin/bash
# Final round-2 evidence for the GP chain after the two-lane triangular inverse / fused contraction / fp16 GPDBN backward
# (1 GPU, under gpurun). Every profiled command first runs plain and must exit 0; .ncu-rep files are exported to csv and removed.
set -x
O=gpurun_out
mkdir -p $O
export_rep() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2> /dev/null; rm -f $O/$1.ncu-rep; }
python scripts/gp_once.py > $O/r02f_plain_gp_once.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_gplvm_n5000.csv python scripts/gp_once.py > $O/r02f_ncu_gp.log 2>&1
ncu --set full --clock-control none -k regex:potrf_step -s 40 -c 2 -o $O/r02_prof_potrf_step python scripts/gp_once.py > $O/r02f_ncu_potrf.log 2>&1
export_rep r02_prof_potrf_step
ncu --set full --clock-control none -k regex:"gp_contract_fused|mirror_lower_split|kinv_diag_kernel|transpose_dots" -s 0 -c 4 -o $O/r02_prof_gp_tail python scripts/gp_once.py > $O/r02f_ncu_tail.log 2>&1
export_rep r02_prof_gp_tail
DB_GP_OVERLAP=0 python scripts/timeline_gp.py 2>/dev/null > $O/r02_timeline_gplvm_n5000_one_stream.txt
DB_GP_OVERLAP=1 python scripts/timeline_gp.py 2>/dev/null > $O/r02_timeline_gplvm_n5000_lanes.txt
python scripts/timeline_gpdbn.py 2>/dev/null > $O/r02_timeline_gpdbn_config4_eager.txt
python scripts/time_gp.py > $O/r02_time_gp.txt 2>&1
python scripts/bench_small.py > $O/r02_bench_small.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02f_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02f_ncu_bench.log 2>&1
du -sh $O
