#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/power_triage.txt
for hf in 0 1; do for dbg in 0 2 4 6; do
  TLB200_DISABLE_HF=$hf TLB200_TC_DEBUG=$dbg timeout 120 python scripts/power_triage.py >> gpurun_out/power_triage.txt 2>&1
done; done
RANK_R=32 TLB200_DISABLE_HF=1 timeout 120 python scripts/power_triage.py >> gpurun_out/power_triage.txt 2>&1
RANK_R=32 TLB200_DISABLE_HF=0 TLB200_HF_MIN_RANK=1 timeout 120 python scripts/power_triage.py >> gpurun_out/power_triage.txt 2>&1
cat gpurun_out/power_triage.txt
