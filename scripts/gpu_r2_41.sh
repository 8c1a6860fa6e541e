#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "hals or smoke" 2>&1 | tail -2
timeout 600 python bench.py --workload small --steps 5 --warmup 3 --no-e2e --no-cpu --no-refdriver --no-c3 --no-c2 --no-c4 --no-sustained --no-fp64 > gpurun_out/bench41.json 2> gpurun_out/bench41.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench41.json').read().strip().splitlines()[-1])
print(d['n4']['hals_update'])
P
