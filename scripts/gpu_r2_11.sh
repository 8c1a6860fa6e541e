#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/tests11.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests11.txt
python scripts/ps_trace.py > gpurun_out/ps_trace2.txt 2>&1
python scripts/c3_truth.py > gpurun_out/c3_truth_c.txt 2>&1
timeout 900 python bench.py --workload c2 --no-e2e --no-cpu --no-fp64 --no-refdriver > gpurun_out/bench11_c2.json 2> gpurun_out/bench11.err; echo "bench rc=$?"
grep -v "^$" gpurun_out/tests11.txt | tail -n 8; cat gpurun_out/ps_trace2.txt; cat gpurun_out/c3_truth_c.txt | cut -c1-200; tail -n 3 gpurun_out/bench11.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench11_c2.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['sustained']['value'], d['launches_per_sweep'])
c3=d['c3']; print('c3', c3['value'], c3['ms_per_step'], c3['launches_per_sweep'], c3['parity_vs_exact_hooi_fp64_svd']['max_rel_dev'])
P
