#!/bin/bash
# round 2, GPU call 28: what the driver runs at round end, on the final tree: GPU suite, smoke, both bench arms with default flags
LOG=gpurun_out/r2_call28.log
mkdir -p gpurun_out; : > $LOG
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest28.log 2>&1
echo "pytest rc=$?" >> $LOG; tail -3 gpurun_out/r2_pytest28.log >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1 || echo "SMOKE FAILED" >> $LOG
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench28_ref.json 2> gpurun_out/r2_bench28_ref.err
echo "bench ref rc=$? lines=$(wc -l < gpurun_out/r2_bench28_ref.json)" >> $LOG; cut -c1-300 gpurun_out/r2_bench28_ref.json >> $LOG
timeout 600 python bench.py > gpurun_out/r2_bench28.json 2> gpurun_out/r2_bench28.err
echo "bench rc=$? lines=$(wc -l < gpurun_out/r2_bench28.json)" >> $LOG; cut -c1-400 gpurun_out/r2_bench28.json >> $LOG
tail -30 $LOG
