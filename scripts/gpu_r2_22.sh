#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/power_triage6.txt
for dbg in 6 2 4 0; do
TLB200_TC_DEBUG=$dbg timeout 200 python scripts/power_triage.py 2>&1 | grep -E "HFoff" >> gpurun_out/power_triage6.txt
done
cat gpurun_out/power_triage6.txt
