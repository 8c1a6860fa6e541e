#!/bin/bash
# final bench of both arms with the driver's commands
mkdir -p gpurun_out
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "ours exit $?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; echo "reference exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_reference_n1.json').read().strip().splitlines()[-1])
print('C5', round(d['value'],2), 'e2e', round(d['e2e']['value'],3), 'ref', round(r['value'],4), 'ratio e2e', round(d['e2e']['value']/r['value'],1), 'ratio resident', round(d['value']/r['value']))
print('roofline', round(d['roofline']['frac'],3), d['clocks'])
for k in ('c2','c3','c4','fp64','n4'):
    v=d[k]; print(k, v.get('value', v.get('mttkrp_frac_of_fp64_peak', v.get('masked_sweep'))))
P
