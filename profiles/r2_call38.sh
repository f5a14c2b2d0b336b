#!/bin/bash
# round 2, GPU call 38 (2 GPUs): both bench arms launched exactly as the driver launches them for N > 1
LOG=gpurun_out/r2_call38.log
mkdir -p gpurun_out; : > $LOG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench38_ref.json 2> gpurun_out/r2_bench38_ref.err
echo "reference arm rc=$? stdout lines=$(wc -l < gpurun_out/r2_bench38_ref.json)" >> $LOG; cut -c1-260 gpurun_out/r2_bench38_ref.json >> $LOG
timeout 500 $TR bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench38.json 2> gpurun_out/r2_bench38.err
echo "our arm rc=$? stdout lines=$(wc -l < gpurun_out/r2_bench38.json)" >> $LOG; cut -c1-260 gpurun_out/r2_bench38.json >> $LOG
tail -8 $LOG
