set -x
timeout 300 python profiles/tc_error_stats.py > gpurun_out/tc_err_new.txt 2>&1; tail -4 gpurun_out/tc_err_new.txt
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scale.py tests/test_localizer.py tests/test_gpu_weights.py -m gpu -q -x > gpurun_out/pytest_tc.log 2>&1; tail -4 gpurun_out/pytest_tc.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -2 gpurun_out/bench_r1.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
