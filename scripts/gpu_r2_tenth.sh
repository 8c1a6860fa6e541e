#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/tests10.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests10.txt
timeout 900 python bench.py --no-e2e --no-cpu --no-fp64 --no-c3 > gpurun_out/bench10_c5_n1.json 2> gpurun_out/bench10_c5_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c5slab --no-e2e --no-cpu --no-c2 --no-c3 --no-fp64 --no-refdriver > gpurun_out/bench10_c5slab.json 2> gpurun_out/bench10_c5slab.err
TLB200_FUSED_UPDATE=0 timeout 600 python bench.py --workload c5slab --no-e2e --no-cpu --no-c2 --no-c3 --no-fp64 --no-refdriver > gpurun_out/bench10_c5slab_unfused.json 2> /dev/null
grep -v "^$" gpurun_out/tests10.txt | tail -n 12; tail -n 5 gpurun_out/bench10_c5_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench10_c5_n1.json').read().strip().splitlines()[-1])
print('c5', d['value'], d['launches_per_sweep'], d['roofline']['achieved'], d['clocks']['sm_mhz'], d['final_rel_error'], d['parity'])
print('c2', d['c2']['value'], d['c2']['sustained']['value'], d['c2']['launches_per_sweep'], d['c2']['final_rel_error'])
for f in ('bench10_c5slab','bench10_c5slab_unfused'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['launches_per_sweep'])
P
