#!/bin/bash
# round 2, GPU call 4: enc_tail with early TMEM hand-off + single-pass LayerNorm, decoder FFN fp16 hidden; full suite; bench; reference arm
mkdir -p gpurun_out
LOG=gpurun_out/r2_call4.log
: > $LOG
rm -f gpurun_out/hatches.log
for cg in 1 2; do
  timeout 150 python profiles/enc_tail_check.py 38000 $cg >> $LOG 2>&1 || echo "FAILED cg=$cg rc=$?" >> $LOG
done
timeout 200 python __graft_entry__.py --smoke >> $LOG 2>&1 || echo "SMOKE FAILED rc=$?" >> $LOG
timeout 1700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py -s > gpurun_out/r2_pytest4.log 2>&1
echo "pytest rc=$?" >> $LOG
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/r2_pytest4.log | head -70 >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
echo "bench rc=$?" >> $LOG
cat gpurun_out/r2_bench4.json >> $LOG
tail -3 gpurun_out/r2_bench4.err >> $LOG
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench4_ref.json 2> gpurun_out/r2_bench4_ref.err
echo "bench ref rc=$?" >> $LOG
cat gpurun_out/r2_bench4_ref.json >> $LOG
tail -3 gpurun_out/r2_bench4_ref.err >> $LOG
CONE_ENC_TAIL_CG=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tail -s 2 -c 1 -o gpurun_out/r2_prof4_enc_tail -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu4.log 2>&1
echo "ncu rc=$?" >> $LOG
tail -90 $LOG
