#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2slab --steps 200 --warmup 20 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c4 --no-n4 --no-c2 --no-sustained > gpurun_out/bench38.json 2> gpurun_out/bench38.err; echo "rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench38.json').read().strip().splitlines()[-1])
print('c2slab', round(d['value'],1), round(d['ms_per_step']*1e3,1), 'us', d['launches_per_sweep'], [round(v) for v in d['roofline']['per_mode_gbs']], d['roofline']['ms_per_launch'])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches38.csv python bench.py --workload c2slab --steps 3 --warmup 3 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c4 --no-n4 --no-c2 --no-sustained > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/launches38.csv | head -16
