#!/bin/bash
# round 2, GPU call 36: the fall-back switches still work on the final build (unfused tail; gathered copy; CTA-pair width 1)
LOG=gpurun_out/r2_call36.log
mkdir -p gpurun_out; : > $LOG
for v in "CONE_FUSED_TAIL=0" "CONE_TAIL_GATHER=0" "CONE_ENC_TAIL_CG=1"; do
  env $v timeout 300 python -m pytest tests/test_gpu_tc.py -k "end_to_end or dense_vs_oracle" -x -q > gpurun_out/r2_pytest36.log 2>&1
  echo "$v pytest rc=$?" >> $LOG; tail -2 gpurun_out/r2_pytest36.log >> $LOG
done
tail -12 $LOG
