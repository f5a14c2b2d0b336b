set -x
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_m1.json 2> gpurun_out/bench_m1.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_m2.json 2> gpurun_out/bench_m2.err
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scale.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; tail -3 gpurun_out/pytest_tc.log
