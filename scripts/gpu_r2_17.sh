#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/power_triage2.txt
TLB200_TC_NB=2 TLB200_DISABLE_HF=0 TLB200_TC_DEBUG=0 timeout 120 python scripts/power_triage.py >> gpurun_out/power_triage2.txt 2>&1
TLB200_TC_NB=2 TLB200_DISABLE_HF=0 TLB200_TC_DEBUG=6 timeout 120 python scripts/power_triage.py >> gpurun_out/power_triage2.txt 2>&1
TLB200_TC_NB=2 TLB200_DISABLE_HF=1 TLB200_TC_DEBUG=0 timeout 120 python scripts/power_triage.py >> gpurun_out/power_triage2.txt 2>&1
TLB200_DISABLE_HF=1 TLB200_TC_DEBUG=0 timeout 120 python scripts/power_triage.py >> gpurun_out/power_triage2.txt 2>&1
grep HFoff gpurun_out/power_triage2.txt
TLB200_TC_NB=2 timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c2 > gpurun_out/bench17_c5.json 2> gpurun_out/bench17_c5.err; echo "bench c5 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench17_c5.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],3), d['roofline']['frac'], d['roofline']['per_mode_gbs'], d['clocks']['sm_mhz'], d.get('sustained',{}).get('value'))
P
