set -x
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_weights.py tests/test_localizer.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; tail -6 gpurun_out/pytest_tc.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_h2.json 2> gpurun_out/bench_h2.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_h2b.json 2> gpurun_out/bench_h2b.err
