set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err
