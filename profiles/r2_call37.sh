#!/bin/bash
# round 2, GPU call 37: mma.sync encoder attention at 5 CTAs per SM (72 registers, 60 bytes of spills) against 4 (96 registers)
LOG=gpurun_out/r2_call37.log
mkdir -p gpurun_out; : > $LOG
for i in 1 2; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench37.json 2> gpurun_out/r2_bench37.err
  python - <<PY >> $LOG 2>&1
import json
d = json.loads(open("gpurun_out/r2_bench37.json").read().strip().splitlines()[-1])
print("run $i ms/step", round(d["ms_per_step"], 2), "q/s", round(d["value"]), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items() if isinstance(v, dict) and k in ("enc_attention", "enc_tail", "enc_tail_gather")})
PY
done
timeout 200 python -m pytest tests/test_gpu_tc.py -k "end_to_end" -x -q > gpurun_out/r2_pytest37.log 2>&1; echo "pytest rc=$?" >> $LOG
tail -5 $LOG
