#!/bin/bash
# round 2, GPU call 34: full record of the final build (window slicing fused into the encoder tail, residual panels first): GPU suite, smoke, bench (both arms),
# launch list with DRAM bytes and pipe counters, ncu --set full summary of the large kernels, Ego4D and stress configurations
LOG=gpurun_out/r2_call34.log
mkdir -p gpurun_out; : > $LOG
rm -f gpurun_out/hatches.log
source profiles/gpu_guard.sh
timeout 200 python __graft_entry__.py --smoke >> $LOG 2>&1 || echo "SMOKE FAILED rc=$?" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py -s > gpurun_out/r2_pytest34.log 2>&1
echo "pytest rc=$?" >> $LOG
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/r2_pytest34.log | grep -v "R@K\|hatch" | head -40 >> $LOG
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench34.json 2> gpurun_out/r2_bench34.err
echo "bench rc=$?" >> $LOG
cat gpurun_out/r2_bench34.json >> $LOG
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench34_ref.json 2> gpurun_out/r2_bench34_ref.err
echo "bench ref rc=$?" >> $LOG
cut -c1-300 gpurun_out/r2_bench34_ref.json >> $LOG
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none --csv --log-file gpurun_out/r2_launches34.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_list34.log 2>&1
echo "ncu list rc=$?" >> $LOG
timeout 600 ncu --set full --clock-control none \
  -k regex:"enc_tail|enc_attention_f16|dec_cross_attention_mem|frame_scores|sgemm_nt|span_mean_pool|window_ranklist" \
  -s 11 -c 11 -o /tmp/r2_prof34_step -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu34.log 2>&1
echo "ncu full rc=$?" >> $LOG
python profiles/ncu_summary.py /tmp/r2_prof34_step.ncu-rep > gpurun_out/r2_ncu34_summary.txt 2>> $LOG
timeout 300 python bench.py --config ego4d --movies 320 --queries-per-movie 5 --videos-per-step 64 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench34_ego4d.json 2> gpurun_out/r2_bench34_ego4d.err
echo "bench ego4d rc=$?" >> $LOG
cut -c1-330 gpurun_out/r2_bench34_ego4d.json >> $LOG
timeout 300 python bench.py --workload stress --stress-queries 10000 --no-cpu-baseline > gpurun_out/r2_bench34_stress.json 2> gpurun_out/r2_bench34_stress.err
echo "bench stress rc=$?" >> $LOG
cut -c1-330 gpurun_out/r2_bench34_stress.json >> $LOG
tail -90 $LOG | cut -c1-1600
