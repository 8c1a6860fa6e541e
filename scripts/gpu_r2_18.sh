#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/power_triage3.txt
for cfg in "3 8" "6 8" "6 12" "6 16" "4 16" "6 24"; do
  set -- $cfg
  echo "== NB=$1 FLUSH=$2" >> gpurun_out/power_triage3.txt
  TLB200_TC_NB=$1 TLB200_TC_FLUSH=$2 timeout 120 python scripts/power_triage.py 2>&1 | grep HFoff >> gpurun_out/power_triage3.txt
  TLB200_TC_NB=$1 TLB200_TC_FLUSH=$2 timeout 300 python -m pytest tests -m gpu -x -q -k "fp16_engine_mttkrp" 2>&1 | tail -2 >> gpurun_out/power_triage3.txt
done
cat gpurun_out/power_triage3.txt
