#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fp16_engine or fused or mttkrp" > gpurun_out/tests19.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests19.txt
: > gpurun_out/power_triage4.txt
for cfg in "6 16" "4 16" "3 16"; do
  set -- $cfg
  echo "== NB_F16=$1 FLUSH_F16=$2" >> gpurun_out/power_triage4.txt
  TLB200_TC_NB_F16=$1 TLB200_TC_FLUSH_F16=$2 timeout 120 python scripts/power_triage.py 2>&1 | grep HFoff >> gpurun_out/power_triage4.txt
  TLB200_TC_NB_F16=$1 TLB200_TC_FLUSH_F16=$2 timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c2 > gpurun_out/bench19_c5_$1.json 2> gpurun_out/bench19_c5_$1.err; echo "bench c5 rc=$?"
  python - $1 >> gpurun_out/power_triage4.txt <<'P'
import json, sys
d=json.loads(open(f'gpurun_out/bench19_c5_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('C5', round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), [round(v) for v in d['roofline']['per_mode_gbs']], d['clocks']['sm_mhz'], 'sustained', round(d.get('sustained',{}).get('value',0),2))
P
done
cat gpurun_out/power_triage4.txt
