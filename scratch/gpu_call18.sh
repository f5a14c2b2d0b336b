set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_p1.json 2> gpurun_out/bench_p1.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_p2.json 2> gpurun_out/bench_p2.err
timeout 300 python bench.py --no-cpu-baseline --config ego4d --movies 64 --queries-per-movie 16 > gpurun_out/bench_ego4d.json 2> gpurun_out/bench_ego4d.err; tail -2 gpurun_out/bench_ego4d.err
