set -x
timeout 300 python scratch/tc_err.py > gpurun_out/tc_err_new.txt 2>&1; tail -4 gpurun_out/tc_err_new.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_xmem.json 2> gpurun_out/bench_xmem.err; tail -3 gpurun_out/bench_xmem.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_xmem2.json 2> gpurun_out/bench_xmem2.err
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
