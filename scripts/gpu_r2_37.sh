#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/power_triage_hint_ab.txt
for rep in 1 2; do for dbg in 16 0; do
  echo "== TLB200_TC_DEBUG=$dbg (16 = no hint) rep $rep" >> gpurun_out/power_triage_hint_ab.txt
  TLB200_TC_DEBUG=$dbg timeout 200 python scripts/power_triage.py 2>&1 | grep -E "HFoff" >> gpurun_out/power_triage_hint_ab.txt
done; done
cat gpurun_out/power_triage_hint_ab.txt
