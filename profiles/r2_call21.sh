#!/bin/bash
# round 2, GPU call 21: full record of the final build: GPU suite, smoke, bench (both arms), launch list with DRAM bytes, ncu of the top kernels
LOG=gpurun_out/r2_call21.log
mkdir -p gpurun_out; : > $LOG
rm -f gpurun_out/hatches.log
source profiles/gpu_guard.sh
timeout 200 python __graft_entry__.py --smoke >> $LOG 2>&1 || echo "SMOKE FAILED rc=$?" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py -s > gpurun_out/r2_pytest21.log 2>&1
echo "pytest rc=$?" >> $LOG
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/r2_pytest21.log | head -70 >> $LOG
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err
echo "bench rc=$?" >> $LOG
cat gpurun_out/r2_bench21.json >> $LOG
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench21_ref.json 2> gpurun_out/r2_bench21_ref.err
echo "bench ref rc=$?" >> $LOG
cat gpurun_out/r2_bench21_ref.json >> $LOG
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none --csv --log-file gpurun_out/r2_launches21.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_list21.log 2>&1
echo "ncu list rc=$?" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"enc_tail|enc_attention_f16|tc_gemm_kernel<256, true, 0>" -s 6 -c 6 -o gpurun_out/r2_prof21_top -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu21.log 2>&1
echo "ncu full rc=$?" >> $LOG
timeout 300 python bench.py --config ego4d --movies 320 --queries-per-movie 5 --videos-per-step 64 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench21_ego4d.json 2> gpurun_out/r2_bench21_ego4d.err
echo "bench ego4d rc=$?" >> $LOG
cat gpurun_out/r2_bench21_ego4d.json >> $LOG
timeout 300 python bench.py --workload stress --stress-queries 10000 --no-cpu-baseline > gpurun_out/r2_bench21_stress.json 2> gpurun_out/r2_bench21_stress.err
echo "bench stress rc=$?" >> $LOG
cat gpurun_out/r2_bench21_stress.json >> $LOG
tail -3 gpurun_out/r2_bench21_stress.err >> $LOG
CONE_ATTN_TC=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench21_attn_tc.json 2> gpurun_out/r2_bench21_attn_tc.err
echo "bench CONE_ATTN_TC=1 rc=$?" >> $LOG
cat gpurun_out/r2_bench21_attn_tc.json >> $LOG
tail -100 $LOG | cut -c1-1500
