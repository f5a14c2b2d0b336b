set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 400 gpurun_out/bench_2gpu.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_a.json 2> gpurun_out/bench_1gpu_a.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_b.json 2> gpurun_out/bench_1gpu_b.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-sample-queries 32 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err
