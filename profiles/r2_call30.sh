#!/bin/bash
# round 2, GPU call 30: gathered encoder tail with / without the L2 prefetch of the next tile's rows; DRAM bytes of the step
LOG=gpurun_out/r2_call30.log
mkdir -p gpurun_out; : > $LOG
source profiles/gpu_guard.sh
timeout 240 env CONE_TAIL_GATHER_PF=1 python -m pytest tests/test_gpu_tc.py -k "end_to_end" -x -q > gpurun_out/r2_pytest30a.log 2>&1
rc=$?; echo "pytest e2e (prefetch) rc=$rc" >> $LOG; tail -3 gpurun_out/r2_pytest30a.log >> $LOG
if [ $rc != 0 ]; then tail -40 $LOG; exit 1; fi
for v in "CONE_TAIL_GATHER=0" "CONE_TAIL_GATHER_PF=0" "CONE_TAIL_GATHER_PF=1"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-pass > gpurun_out/r2_bench30_$v.json 2> gpurun_out/r2_bench30_$v.err
  echo "bench $v rc=$?" >> $LOG
  python - <<PY >> $LOG 2>&1
import json
try:
    d = json.loads(open("gpurun_out/r2_bench30_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 2), "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: (v["ms"], v["launches"]) for k, v in d.get("stages", {}).items() if isinstance(v, dict) and k in ("enc_tail", "rowops", "enc_attention")})
except Exception as e:
    print("$v parse failed", e)
PY
done
for v in "CONE_TAIL_GATHER=0" "CONE_TAIL_GATHER_PF=0"; do
  env $v timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches30_$v.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu30.log 2>&1
  echo "ncu list $v rc=$?" >> $LOG
  python profiles/summarize_step.py gpurun_out/r2_launches30_$v.csv gpurun_out/r2_step30_$v.md gpurun_out/r2_traffic30_$v.json --step 1 > /dev/null 2>> $LOG
  python - <<PY >> $LOG 2>&1
import json
d = json.load(open("gpurun_out/r2_traffic30_$v.json"))
print("$v step DRAM GB", round(d["step_dram_bytes"] / 1e9, 2), {k: (round(v["dram_bytes_per_launch"] / 1e9, 2), round(v["time_us_per_launch"])) for k, v in d["kernels"].items()})
PY
done
tail -30 $LOG
