#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cp_to_tensor or impute or masked or smoke" > gpurun_out/tests31.txt 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests31.txt
timeout 600 python bench.py --workload small --steps 5 --warmup 3 --no-e2e --no-cpu --no-refdriver --no-c3 --no-c2 --no-sustained --no-fp64 > gpurun_out/bench31.json 2> gpurun_out/bench31.err; echo "bench rc=$?"; tail -3 gpurun_out/bench31.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench31.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('n4')))
P
