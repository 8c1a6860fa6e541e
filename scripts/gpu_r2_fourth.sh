#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/tests4.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests4.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c3b.csv python scripts/prof_c3.py > gpurun_out/prof_c3b.log 2>&1; echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c3b.csv > gpurun_out/launches_c3b_summary.txt 2>&1
timeout 900 python bench.py --no-e2e --no-cpu --no-fp64 > gpurun_out/bench4_c5_n1.json 2> gpurun_out/bench4_c5_n1.err; echo "bench rc=$?"
grep -v "^$" gpurun_out/tests4.txt | tail -n 30; head -n 14 gpurun_out/launches_c3b_summary.txt; tail -n 5 gpurun_out/bench4_c5_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench4_c5_n1.json').read().strip().splitlines()[-1])
print('c5', d['value'], d['launches_per_sweep'], d['roofline']['achieved'], d['clocks'])
print('c2', d['c2']['value'], d['c2']['launches_per_sweep'], d['c2']['roofline']['achieved'], d['c2'].get('reference_driver_on_b200'))
c3=d['c3']; print('c3', c3['value'], c3['ms_per_step'], c3['launches_per_sweep'], c3.get('parity_vs_reference_driver'), c3.get('reference_driver_on_b200'))
P
