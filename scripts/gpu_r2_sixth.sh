#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/tests6.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests6.txt
timeout 900 python bench.py --no-e2e --no-cpu --no-fp64 > gpurun_out/bench6_c5_n1.json 2> gpurun_out/bench6_c5_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c5slab --no-e2e --no-cpu --no-c2 --no-c3 --no-fp64 --no-refdriver > gpurun_out/bench6_c5slab.json 2> gpurun_out/bench6_c5slab.err
grep -v "^$" gpurun_out/tests6.txt | tail -n 12; tail -n 5 gpurun_out/bench6_c5_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench6_c5_n1.json').read().strip().splitlines()[-1])
print('c5', d['value'], d['launches_per_sweep'], d['roofline']['achieved'], d['clocks']['sm_mhz'])
print('c2', d['c2']['value'], d['c2']['sustained']['value'], d['c2']['launches_per_sweep'], d['c2']['roofline']['achieved'], d['c2'].get('reference_driver_on_b200'))
c3=d['c3']; print('c3', c3['value'], c3['ms_per_step'], c3['launches_per_sweep'], c3.get('parity_vs_reference_driver'), c3.get('reference_driver_on_b200'))
d=json.loads(open('gpurun_out/bench6_c5slab.json').read().strip().splitlines()[-1])
print('c5slab', d['value'], d['ms_per_step'])
P
