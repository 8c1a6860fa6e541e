#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/tests9.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests9.txt
timeout 600 python - > gpurun_out/fp64_ab.txt 2>&1 <<'P'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
class E: pass
env = E(); env.device = torch.device("cuda", 0)
print("dmma", json.dumps(bench.fp64_leg(env)))
P
TLB200_FP64_SIMT=1 timeout 600 python - >> gpurun_out/fp64_ab.txt 2>&1 <<'P'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
class E: pass
env = E(); env.device = torch.device("cuda", 0)
print("simt", json.dumps(bench.fp64_leg(env)))
P
timeout 300 python bench.py --workload c1 --no-e2e --no-c2 --no-c3 --no-fp64 --no-refdriver --steps 50 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
grep -v "^$" gpurun_out/tests9.txt | tail -n 8; cat gpurun_out/fp64_ab.txt | cut -c1-900; cut -c1-600 gpurun_out/bench_c1.json; tail -n 3 gpurun_out/bench_c1.err
