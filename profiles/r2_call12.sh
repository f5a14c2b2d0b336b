#!/bin/bash
# round 2, GPU call 12 (N GPUs, pass N as $1): multi-GPU tests, weak-scaling bench, strong scaling (fixed set, LPT), stress (query-sharded)
N=${1:-2}
LOG=gpurun_out/r2_call12_n$N.log
mkdir -p gpurun_out; : > $LOG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ $N = 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r2_pytest12_n$N.log 2>&1
  echo "pytest multi rc=$?" >> $LOG
  grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest12_n$N.log | head >> $LOG
fi
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-parity-pass > gpurun_out/r2_bench12_weak_n$N.json 2> gpurun_out/r2_bench12_weak_n$N.err
echo "bench weak rc=$?" >> $LOG
cat gpurun_out/r2_bench12_weak_n$N.json >> $LOG
timeout 900 $TR bench.py --gpus $N --scaling strong --set-movies 112 > gpurun_out/r2_bench12_strong_n$N.json 2> gpurun_out/r2_bench12_strong_n$N.err
echo "bench strong rc=$?" >> $LOG
cat gpurun_out/r2_bench12_strong_n$N.json >> $LOG
tail -3 gpurun_out/r2_bench12_strong_n$N.err >> $LOG
timeout 600 $TR bench.py --gpus $N --workload stress --stress-queries 10000 > gpurun_out/r2_bench12_stress_n$N.json 2> gpurun_out/r2_bench12_stress_n$N.err
echo "bench stress rc=$?" >> $LOG
cat gpurun_out/r2_bench12_stress_n$N.json >> $LOG
tail -3 gpurun_out/r2_bench12_stress_n$N.err >> $LOG
tail -40 $LOG | cut -c1-1200
