#!/bin/bash
# round 2, GPU call 10: two epilogue groups alternating hidden chunks
LOG=gpurun_out/r2_call10.log
mkdir -p gpurun_out; : > $LOG
source profiles/gpu_guard.sh
timeout 600 python -m pytest tests/test_gpu_enc_tail.py tests/test_gpu_tc.py -m gpu -q -s -x > gpurun_out/r2_pytest10.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $LOG
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest10.log | head -20 >> $LOG
if [ $rc != 0 ]; then tail -30 $LOG; exit 1; fi
for v in "CONE_ENC_TAIL_CG=2"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench10_$v.json 2> gpurun_out/r2_bench10_$v.err
  echo "bench $v rc=$?" >> $LOG
  python - <<PY >> $LOG 2>&1
import json
try:
    d = json.loads(open("gpurun_out/r2_bench10_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 2), "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items() if isinstance(v, dict)})
except Exception as e:
    print("$v parse failed", e)
PY
  tail -2 gpurun_out/r2_bench10_$v.err >> $LOG
done
CONE_ENC_TAIL_CG=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:enc_tail -s 2 -c 1 -o gpurun_out/r2_prof10_enc_tail -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu10.log 2>&1
echo "ncu rc=$?" >> $LOG
tail -40 $LOG
