#!/bin/bash
# round 2, GPU call 33: residual panels first in GEMM0 (producer and MMA order)
LOG=gpurun_out/r2_call33.log
mkdir -p gpurun_out; : > $LOG
source profiles/gpu_guard.sh
timeout 240 python -m pytest tests/test_gpu_tc.py -k "end_to_end" -x -q -s > gpurun_out/r2_pytest33a.log 2>&1
rc=$?; echo "pytest e2e rc=$rc" >> $LOG; grep -E "^\[tc-vs-oracle\].*max|passed|failed|FAILED|Error|error" gpurun_out/r2_pytest33a.log | head -12 >> $LOG
if [ $rc != 0 ]; then
  tail -30 gpurun_out/r2_pytest33a.log >> $LOG
  CONE_TAIL_GATHER=0 timeout 240 python -m pytest tests/test_gpu_tc.py -k "end_to_end" -x -q > gpurun_out/r2_pytest33b.log 2>&1
  echo "same tests with CONE_TAIL_GATHER=0 rc=$?" >> $LOG
  tail -70 $LOG; exit 1
fi
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2_pytest33c.log 2>&1
rc=$?; echo "pytest all rc=$rc" >> $LOG
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest33c.log | head -12 >> $LOG
for v in "CONE_TAIL_GATHER=0" "CONE_TAIL_GATHER=1"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench33_$v.json 2> gpurun_out/r2_bench33_$v.err
  echo "bench $v rc=$?" >> $LOG
  python - <<PY >> $LOG 2>&1
import json
try:
    d = json.loads(open("gpurun_out/r2_bench33_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 2), "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items() if isinstance(v, dict)})
except Exception as e:
    print("$v parse failed", e)
PY
  tail -2 gpurun_out/r2_bench33_$v.err >> $LOG
done
tail -40 $LOG
