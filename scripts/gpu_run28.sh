#!/bin/bash
# 8-GPU verification: sharded-vs-single parity, the default bench command, and the C5 workload
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== dist check (8 GPUs) =="; timeout 200 $TR --master-port 29511 scripts/dist_check.py 2>&1 | grep -E "rank|Error|error" | head -10
echo "== bench 8 GPUs C2 (driver command) =="; timeout 300 $TR --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 5 2> gpurun_out/bench8.err | tee gpurun_out/bench8_c2.json | cut -c1-900; echo "exit ${PIPESTATUS[0]}"; grep -iE "error|Traceback" gpurun_out/bench8.err | head -5
echo "== bench 8 GPUs C5 =="; timeout 300 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 --workload c5 --no-e2e 2> gpurun_out/bench8c5.err | tee gpurun_out/bench8_c5.json | cut -c1-900; echo "exit ${PIPESTATUS[0]}"; grep -iE "error|Traceback" gpurun_out/bench8c5.err | head -5
