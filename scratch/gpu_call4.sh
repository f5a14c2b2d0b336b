set -x
CONE_TC_SPIN=1 timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest_spin.log 2>&1; tail -3 gpurun_out/pytest_spin.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
CONE_TC_SPIN=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_spin.json 2> gpurun_out/bench_spin.err
CONE_TC_SPIN=1 CONE_TC_EPI_DB=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_spindb.json 2> gpurun_out/bench_spindb.err
