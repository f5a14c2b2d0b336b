#!/bin/bash
# round 2, GPU call 32: gather4 tail vs gathered copy (bit identity), edge cases and the single-call pre-filter on the final build
LOG=gpurun_out/r2_call32.log
mkdir -p gpurun_out; : > $LOG
timeout 600 python -m pytest tests/test_gpu_tail_gather.py tests/test_gpu_edge.py tests/test_gpu_enc_tail.py -m gpu -q > gpurun_out/r2_pytest32.log 2>&1
echo "pytest rc=$?" >> $LOG; tail -6 gpurun_out/r2_pytest32.log >> $LOG
tail -12 $LOG
