#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fp16_engine or range_hint or backend_registers" > gpurun_out/tests30.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests30.txt
timeout 600 python bench.py --workload small --steps 5 --warmup 3 --no-e2e --no-cpu --no-refdriver --no-c3 --no-c2 --no-sustained --no-fp64 > gpurun_out/bench30.json 2> gpurun_out/bench30.err; echo "bench rc=$?"; tail -3 gpurun_out/bench30.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench30.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('n4'), indent=1))
P
