set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; tail -8 gpurun_out/pytest_multi.log
