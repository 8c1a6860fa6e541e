#!/bin/bash
# visit 13: error folded into the last solve (9 launches per sweep)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or parafac or als or smoke or sharded or launch" > gpurun_out/tests13.txt 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/tests13.txt
timeout 600 python bench.py --workload c5slab --steps 20 --warmup 5 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench13.json').read().strip().splitlines()[-1])
print('c5slab', d['value'], d['ms_per_step'], d.get('gpu_launches'), d['roofline']['frac'], d.get('c2',{}).get('value'), d.get('c2',{}).get('gpu_launches'))
P
