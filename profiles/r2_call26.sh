#!/bin/bash
# round 2, GPU call 26: one `ncu --set full` capture of the large kernels of one step (final build); the summary is made on the box
# (the report itself can exceed what travels back)
LOG=gpurun_out/r2_call26.log
mkdir -p gpurun_out; : > $LOG
timeout 600 ncu --set full --clock-control none \
  -k regex:"enc_tail|enc_attention_f16|dec_cross_attention_mem|frame_scores|sgemm_nt|gather_window_rows_f16|span_mean_pool|window_ranklist" \
  -s 12 -c 12 -o /tmp/r2_prof26_step -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu26.log 2>&1
echo "ncu rc=$?" >> $LOG
python profiles/ncu_summary.py /tmp/r2_prof26_step.ncu-rep > gpurun_out/r2_ncu26_summary.txt 2>> $LOG
ls -la /tmp/r2_prof26_step.ncu-rep >> $LOG
sz=$(stat -c %s /tmp/r2_prof26_step.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -lt 40000000 ]; then cp /tmp/r2_prof26_step.ncu-rep gpurun_out/; fi
cat gpurun_out/r2_ncu26_summary.txt | cut -c1-260 >> $LOG
tail -30 $LOG
