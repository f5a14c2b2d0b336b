#!/bin/bash
# round 2, GPU call 25: the single-call pre-filter against the oracle and the staged entry points
LOG=gpurun_out/r2_call25.log
mkdir -p gpurun_out; : > $LOG
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "prefilter or video_context or ranklist" > gpurun_out/r2_pytest25.log 2>&1
echo "pytest rc=$?" >> $LOG; tail -30 gpurun_out/r2_pytest25.log >> $LOG
tail -40 $LOG
