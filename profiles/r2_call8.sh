#!/bin/bash
# round 2, GPU call 8: two MMA-issuing warps in enc_tail
mkdir -p gpurun_out
LOG=gpurun_out/r2_call8.log
: > $LOG
for cg in 1 2; do
  timeout 150 python profiles/enc_tail_check.py 38000 $cg >> $LOG 2>&1 || echo "FAILED cg=$cg rc=$?" >> $LOG
done
timeout 1200 python -m pytest tests/test_gpu_enc_tail.py tests/test_gpu_tc.py tests/test_gpu_weights.py -m gpu -q -s > gpurun_out/r2_pytest8.log 2>&1
echo "pytest rc=$?" >> $LOG
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest8.log | head -20 >> $LOG
for v in "CONE_ENC_TAIL_CG=1" "CONE_ENC_TAIL_CG=2"; do
  env $v timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench8_$v.json 2> gpurun_out/r2_bench8_$v.err
  echo "bench $v rc=$?" >> $LOG
  python - <<PY >> $LOG 2>&1
import json
try:
    d = json.loads(open("gpurun_out/r2_bench8_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 2), "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items() if isinstance(v, dict)})
except Exception as e:
    print("$v parse failed", e)
PY
  tail -3 gpurun_out/r2_bench8_$v.err >> $LOG
done
CONE_ENC_TAIL_CG=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tail -s 2 -c 1 -o gpurun_out/r2_prof8_enc_tail -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu8.log 2>&1
echo "ncu rc=$?" >> $LOG
tail -40 $LOG
