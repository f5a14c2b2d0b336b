set -x
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_f1.json 2> gpurun_out/bench_f1.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_f2.json 2> gpurun_out/bench_f2.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_f3.json 2> gpurun_out/bench_f3.err
