#!/bin/bash
# round 2, GPU call 26: one `ncu --set full` capture of every kernel above 2 % of the step (final build), for profiles/r02_ncu_full_summary.txt
LOG=gpurun_out/r2_call26.log
mkdir -p gpurun_out; : > $LOG
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"enc_tail|enc_attention_f16|tc_gemm_kernel|dec_cross_attention_mem|frame_scores|sgemm_nt|gather_window_rows_f16|span_mean_pool|window_ranklist|split3_f16" \
  -s 60 -c 60 -o gpurun_out/r2_prof26_step -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu26.log 2>&1
echo "ncu rc=$?" >> $LOG
tail -5 gpurun_out/r2_ncu26.log >> $LOG
tail -20 $LOG
