#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --workload small > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_small.json").read().strip().splitlines()[-1])
print({k: ("UNAVAILABLE " + v["unavailable"] if isinstance(v, dict) and "unavailable" in v else type(v).__name__) for k, v in d.items()})
P
