#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mttkrp or fp16_engine or mode_dot" > gpurun_out/tests36.txt 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests36.txt
timeout 200 python scripts/power_triage.py 2>&1 | grep -E "HFoff|ceiling" > gpurun_out/power_triage_hint.txt; cat gpurun_out/power_triage_hint.txt
timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c4 --no-n4 > gpurun_out/bench36.json 2> gpurun_out/bench36.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench36.json').read().strip().splitlines()[-1])
print('C5', round(d['value'],2), round(d['roofline']['frac'],3), [round(v) for v in d['roofline']['per_mode_gbs']], d['clocks']['sm_mhz'], 'sustained', round(d['sustained']['value'],2), 'c2', round(d['c2']['value'],1), round(d['c2']['sustained']['value'],1))
P
