set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
