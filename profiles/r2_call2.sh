#!/bin/bash
# round 2, GPU call 2: full GPU suite with the new precision design, bench A/B (fused tail off / cg 1 / cg 2), launch list, ncu of enc_tail
mkdir -p gpurun_out
LOG=gpurun_out/r2_call2.log
: > $LOG
rm -f gpurun_out/hatches.log
timeout 1700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py -s > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?" >> $LOG
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/r2_pytest2.log | head -60 >> $LOG
for v in "CONE_FUSED_TAIL=0" "CONE_ENC_TAIL_CG=1" "CONE_ENC_TAIL_CG=2"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench2_$v.json 2> gpurun_out/r2_bench2_$v.err
  echo "bench $v rc=$?" >> $LOG
  python - <<PY >> $LOG 2>&1
import json
try:
    d = json.loads(open("gpurun_out/r2_bench2_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 2), "q/s", round(d["value"]), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items()})
except Exception as e:
    print("$v parse failed", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r2_launches2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_bench2.log 2>&1
echo "ncu list rc=$?" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tail -s 2 -c 1 -o gpurun_out/r2_prof_enc_tail -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_full2.log 2>&1
echo "ncu full rc=$?" >> $LOG
tail -80 $LOG
