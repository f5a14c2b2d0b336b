set -x
timeout 600 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
timeout 600 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
