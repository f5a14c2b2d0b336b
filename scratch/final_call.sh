set -x
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; tail -4 gpurun_out/pytest_tc.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
