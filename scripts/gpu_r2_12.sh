#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "tucker or orthonormalize or hals or smoke" > gpurun_out/tests12.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests12.txt
python scripts/ps_trace.py > gpurun_out/ps_trace3.txt 2>&1
echo skip
timeout 900 python bench.py --workload small --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c2 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench rc=$?"
grep -v "^$" gpurun_out/tests12.txt | tail -n 6; cat gpurun_out/ps_trace3.txt; tail -n 3 gpurun_out/bench12.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench12.json').read().strip().splitlines()[-1])
c3=d['c3']; print('c3', c3['value'], c3['ms_per_step'], c3['launches_per_sweep'], c3['parity_vs_exact_hooi_fp64_svd']['max_rel_dev'])
P
