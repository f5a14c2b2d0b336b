set -x
timeout 300 python scratch/tc_err.py > gpurun_out/tc_err_new.txt 2>&1; tail -4 gpurun_out/tc_err_new.txt
CONE_XATTN_KV=1 timeout 300 python scratch/tc_err.py > gpurun_out/tc_err_old.txt 2>&1; tail -4 gpurun_out/tc_err_old.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_xmem.json 2> gpurun_out/bench_xmem.err; tail -3 gpurun_out/bench_xmem.err
CONE_XATTN_KV=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_xkv.json 2> gpurun_out/bench_xkv.err
