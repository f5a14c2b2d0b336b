#!/bin/bash
# round 2, GPU call 24: the whole GPU suite on the final tree (incl. the edge cases and the opt-in attention), smoke, default bench
LOG=gpurun_out/r2_call24.log
mkdir -p gpurun_out; : > $LOG
rm -f gpurun_out/hatches.log
source profiles/gpu_guard.sh
timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -x -q -s > gpurun_out/r2_pytest24_edge.log 2>&1
echo "pytest edge rc=$?" >> $LOG; tail -25 gpurun_out/r2_pytest24_edge.log >> $LOG
timeout 200 python __graft_entry__.py --smoke >> $LOG 2>&1 || echo "SMOKE FAILED rc=$?" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2_pytest24.log 2>&1
echo "pytest rc=$?" >> $LOG
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest24.log | head -20 >> $LOG
timeout 600 python bench.py > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err
echo "bench rc=$?" >> $LOG
cut -c1-900 gpurun_out/r2_bench24.json >> $LOG
cat gpurun_out/hatches.log >> $LOG 2>/dev/null
tail -80 $LOG | cut -c1-1000
