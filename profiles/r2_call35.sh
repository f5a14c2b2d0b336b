#!/bin/bash
# round 2, GPU call 35: ncu --set full with source counters of the two launches of the fused tail in one step (layer 0: gather4, layer 1: dense)
LOG=gpurun_out/r2_call35.log
mkdir -p gpurun_out; : > $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tail -s 2 -c 2 -o gpurun_out/r2_prof35_tail -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu35.log 2>&1
echo "ncu rc=$?" >> $LOG; ls -la gpurun_out/r2_prof35_tail.ncu-rep >> $LOG
tail -5 $LOG
