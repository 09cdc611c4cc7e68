#!/bin/bash
# Round-2 ncu evidence (1 GPU, under gpurun). Every profiled command first runs plain and must exit 0.
# The .ncu-rep files are exported to csv pages on the box and removed (gpurun_out is capped at 64 MiB).
set -x
O=gpurun_out
export_rep() {   # $1 = report stem: raw page always, source page when $2 = src
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2> /dev/null
  if [ "$2" = "src" ]; then ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_source.csv 2> /dev/null; fi
  rm -f $O/$1.ncu-rep
}
if [ "$1" != "gp" ]; then
python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02_ncu_bench.log 2>&1
python scripts/cd_once.py > $O/r02_plain_cd_once.log 2>&1 || exit 1
# launches 20..22 of tc_gemm_kernel: last down, last up, the fused dW launch
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 19 -c 3 -o $O/r02_prof_tc_gemm python scripts/cd_once.py > $O/r02_ncu_tc.log 2>&1
export_rep r02_prof_tc_gemm src
ncu --set full --clock-control none -k regex:"apply_update|tc_bias_grads|cvt_transpose|rescale_after" -s 1 -c 8 -o $O/r02_prof_cd_tail python scripts/cd_once.py > $O/r02_ncu_tail.log 2>&1
export_rep r02_prof_cd_tail
fi
python scripts/gp_once.py > $O/r02_plain_gp_once.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_gplvm_n5000.csv python scripts/gp_once.py > $O/r02_ncu_gp.log 2>&1
ncu --set full --clock-control none -k regex:"gp_contract_mem|dot_sum" -s 0 -c 4 -o $O/r02_prof_gp_contract python scripts/gp_once.py > $O/r02_ncu_gp2.log 2>&1
export_rep r02_prof_gp_contract
# the three big 3 x fp16 products of one evaluation: K^-1 = U U^T (launch 26), K^-1 Y (27), A A^T (28); and one K = 512 trailing update (0)
ncu --set full --clock-control none -k regex:f16x3_gemm -s 26 -c 3 -o $O/r02_prof_f16x3_big python scripts/gp_once.py > $O/r02_ncu_gp3.log 2>&1
export_rep r02_prof_f16x3_big
ncu --set full --clock-control none -k regex:f16x3_gemm -s 0 -c 1 -o $O/r02_prof_f16x3_trailing python scripts/gp_once.py > $O/r02_ncu_gp4.log 2>&1
export_rep r02_prof_f16x3_trailing
du -sh $O
