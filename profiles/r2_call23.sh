#!/bin/bash
# round 2, GPU call 23: tcgen05 encoder attention v10 (position rows through a per-thread cp.async ring)
LOG=gpurun_out/r2_call23.log
mkdir -p gpurun_out; : > $LOG
source profiles/gpu_guard.sh
export CONE_ATTN_TC=1
timeout 240 python -m pytest tests/test_gpu_tc.py -k "dense_vs_oracle or other_window or end_to_end" -x -q -s > gpurun_out/r2_pytest23a.log 2>&1
rc=$?; echo "pytest dense rc=$rc" >> $LOG; grep -E "^\[tc-vs-oracle\].*max|passed|failed|FAILED|Error" gpurun_out/r2_pytest23a.log | head -12 >> $LOG
if [ $rc != 0 ]; then tail -30 gpurun_out/r2_pytest23a.log >> $LOG; tail -60 $LOG; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_localizer.py tests/test_gpu_parity.py -m gpu -x -q -s > gpurun_out/r2_pytest23c.log 2>&1
rc=$?; echo "pytest tc rc=$rc" >> $LOG
grep -E "^\[tc-vs-oracle\].*max|passed|failed|FAILED|Error" gpurun_out/r2_pytest23c.log | head -12 >> $LOG
if [ $rc != 0 ]; then tail -60 $LOG; exit 1; fi
for v in "CONE_ATTN_TC=0" "CONE_ATTN_TC=1"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench23_$v.json 2> gpurun_out/r2_bench23_$v.err
  echo "bench $v rc=$?" >> $LOG
  python - <<PY >> $LOG 2>&1
import json
try:
    d = json.loads(open("gpurun_out/r2_bench23_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 2), "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items() if isinstance(v, dict)})
except Exception as e:
    print("$v parse failed", e)
PY
  tail -2 gpurun_out/r2_bench23_$v.err >> $LOG
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:enc_attn_tc -s 0 -c 2 -o gpurun_out/r2_prof23_attn -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu23.log 2>&1
echo "ncu rc=$?" >> $LOG
tail -60 $LOG
