set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_xmem.json 2> gpurun_out/bench_xmem.err; tail -3 gpurun_out/bench_xmem.err
