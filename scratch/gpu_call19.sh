set -x
CONE_PROJ_TC=1 timeout 300 python scratch/tc_err.py > gpurun_out/tc_err_proj.txt 2>&1; tail -4 gpurun_out/tc_err_proj.txt
timeout 300 python scratch/tc_err.py > gpurun_out/tc_err_base.txt 2>&1; tail -4 gpurun_out/tc_err_base.txt
CONE_PROJ_TC=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_proj.json 2> gpurun_out/bench_proj.err; tail -2 gpurun_out/bench_proj.err
