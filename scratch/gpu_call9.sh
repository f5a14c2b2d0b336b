set -x
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_xmem.json 2> gpurun_out/bench_xmem.err; tail -3 gpurun_out/bench_xmem.err
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dec_cross_attention_mem' -s 2 -c 1 -o gpurun_out/prof_xmem2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_xmem.log 2>&1
