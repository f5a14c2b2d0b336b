#!/bin/bash
# Shared prologue of the round-2 GPU call scripts: bring-up check of the fused encoder tail under its own timeout; a
# protocol bug in a warp-specialised kernel shows up as a HANG, so everything after it is skipped when the check fails
# (the remote call cannot be cancelled and GPU-minutes are budgeted).
mkdir -p gpurun_out
guard_ok=1
for cg in 1 2; do
  timeout 120 python profiles/enc_tail_check.py 38000 $cg >> "$LOG" 2>&1 || { echo "FAILED enc_tail check cg=$cg rc=$?" >> "$LOG"; guard_ok=0; }
done
if [ $guard_ok = 0 ]; then echo "guard failed: skipping the rest of the call" >> "$LOG"; tail -20 "$LOG"; exit 1; fi
