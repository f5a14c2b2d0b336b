set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dec_cross_attention_mem' -s 2 -c 1 -o gpurun_out/prof_xmem python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_xmem.log 2>&1
ls -la gpurun_out/prof_xmem.ncu-rep
