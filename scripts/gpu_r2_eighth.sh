#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/tests8.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests8.txt
python scripts/c3_truth.py > gpurun_out/c3_truth_b.txt 2>&1
timeout 900 python bench.py --no-e2e --no-cpu --no-fp64 --no-c2 > gpurun_out/bench8_c5_n1.json 2> gpurun_out/bench8_c5_n1.err; echo "bench rc=$?"
grep -v "^$" gpurun_out/tests8.txt | tail -n 8; cat gpurun_out/c3_truth_b.txt; tail -n 5 gpurun_out/bench8_c5_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench8_c5_n1.json').read().strip().splitlines()[-1])
print('c5', d['value'], d['launches_per_sweep'], d['roofline']['achieved'], d['clocks']['sm_mhz'])
c3=d['c3']; print('c3', c3['value'], c3['ms_per_step'], c3['launches_per_sweep'], c3.get('parity_vs_exact_hooi_fp64_svd'), c3.get('reference_driver_on_b200'))
P
