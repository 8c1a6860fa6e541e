#!/bin/bash
mkdir -p gpurun_out
echo skip tests
for v in 0 1; do
TLB200_FP64_NO_ASYNC=$v timeout 600 python bench.py --workload small --steps 20 --warmup 5 --no-e2e --no-cpu --no-refdriver --no-c3 --no-c2 --no-sustained > gpurun_out/bench26_$v.json 2> gpurun_out/bench26_$v.err; echo "bench rc=$?"
python - "$v" <<'P'
import json, sys
d=json.loads(open(f'gpurun_out/bench26_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('no_async=', sys.argv[1], json.dumps(d.get('fp64'))[:1500])
P
done
