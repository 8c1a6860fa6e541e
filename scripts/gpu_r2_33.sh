#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests33.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests33.txt
timeout 600 python bench.py --workload small --steps 10 --warmup 3 --no-e2e --no-cpu --no-refdriver --no-c3 --no-c2 --no-sustained --no-fp64 > gpurun_out/bench33.json 2> gpurun_out/bench33.err; echo "bench rc=$?"; tail -3 gpurun_out/bench33.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench33.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('c4')))
print(json.dumps(d.get('n4')))
P
