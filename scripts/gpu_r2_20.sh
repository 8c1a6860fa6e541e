#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/power_triage5.txt
timeout 200 python scripts/power_triage.py 2>&1 | grep -E "HFoff|ceiling" >> gpurun_out/power_triage5.txt
RANK_R=32 TLB200_HF_MIN_RANK=1 timeout 200 python scripts/power_triage.py 2>&1 | grep -E "HFoff|ceiling" >> gpurun_out/power_triage5.txt
cat gpurun_out/power_triage5.txt
for mr in 33 1; do
TLB200_HF_MIN_RANK=$mr timeout 900 python bench.py --workload c2 --steps 20 --warmup 5 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 > gpurun_out/bench20_c2_$mr.json 2> gpurun_out/bench20_c2_$mr.err; echo "bench c2 rc=$?"
python - $mr <<'P'
import json, sys
d=json.loads(open(f'gpurun_out/bench20_c2_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('C2 minrank', sys.argv[1], round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), [round(v) for v in d['roofline']['per_mode_gbs']], d['roofline']['kernel'][:24], d['clocks']['sm_mhz'], 'sustained', round(d.get('sustained',{}).get('value',0),2))
P
done
