#!/bin/bash
# round 2, GPU call 1: bring-up of the fused encoder tail (each check under its own timeout), then the GPU suite + bench
mkdir -p gpurun_out
LOG=gpurun_out/r2_call1.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv >> $LOG 2>&1
ok1=1; ok2=1
for M in 300 5000 38000; do
  timeout 150 python profiles/enc_tail_check.py $M 1 >> $LOG 2>&1 || { echo "FAILED cg=1 M=$M rc=$?" >> $LOG; ok1=0; }
done
for M in 300 5000 38000; do
  timeout 150 python profiles/enc_tail_check.py $M 2 >> $LOG 2>&1 || { echo "FAILED cg=2 M=$M rc=$?" >> $LOG; ok2=0; }
done
if [ $ok2 = 0 ]; then export CONE_ENC_TAIL_CG=1; echo "using CONE_ENC_TAIL_CG=1" >> $LOG; fi
if [ $ok1 = 0 ] && [ $ok2 = 0 ]; then export CONE_FUSED_TAIL=0; echo "using CONE_FUSED_TAIL=0" >> $LOG; fi
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py -s > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> $LOG
tail -40 gpurun_out/r2_pytest1.log >> $LOG
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?" >> $LOG
cat gpurun_out/r2_bench1.json >> $LOG
tail -5 gpurun_out/r2_bench1.err >> $LOG
tail -60 $LOG
