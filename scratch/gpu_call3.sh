set -x
CONE_TC_EPI_DB=1 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/pytest_db.log 2>&1; tail -5 gpurun_out/pytest_db.log
python bench.py --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
CONE_TC_EPI_DB=1 python bench.py --no-cpu-baseline > gpurun_out/bench_db.json 2> gpurun_out/bench_db.err
python bench.py --no-cpu-baseline > gpurun_out/bench_base2.json 2> gpurun_out/bench_base2.err
CONE_TC_EPI_DB=1 python bench.py --no-cpu-baseline > gpurun_out/bench_db2.json 2> gpurun_out/bench_db2.err
for f in base db base2 db2; do python profiles/show_bench.py gpurun_out/bench_$f.json | head -8; done
