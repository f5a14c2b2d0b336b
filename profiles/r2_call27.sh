#!/bin/bash
# round 2, GPU call 27: timing experiment — the tcgen05 attention with the in-kernel position add switched off (results wrong): what would it run at
# if the position term arrived inside q and k?
LOG=gpurun_out/r2_call27.log
mkdir -p gpurun_out; : > $LOG
source profiles/gpu_guard.sh
for v in "CONE_ATTN_TC=1 CONE_ATTN_TC_NOPOS=0" "CONE_ATTN_TC=1 CONE_ATTN_TC_NOPOS=1"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench27.json 2> gpurun_out/r2_bench27.err
  echo "bench $v rc=$?" >> $LOG
  python - <<PY >> $LOG 2>&1
import json
try:
    d = json.loads(open("gpurun_out/r2_bench27.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 2), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items() if isinstance(v, dict)})
except Exception as e:
    print("$v parse failed", e)
PY
done
tail -12 $LOG
