#!/bin/bash
# 2-GPU check of the sharded driver: parity with the single-GPU trajectory, then the bench command
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== dist check (2 GPUs) =="; timeout 200 $TR --master-port 29511 scripts/dist_check.py 2>&1 | grep -E "rank|Error|error" | head -4
echo "== bench 2 GPUs C2 =="; timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e 2> gpurun_out/bench2.err | tee gpurun_out/bench2_c2.json | cut -c1-200; echo "exit ${PIPESTATUS[0]}"; grep -iE "error|Traceback" gpurun_out/bench2.err | head -5
echo "== bench 2 GPUs C5 =="; timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload c5 --no-e2e 2> gpurun_out/bench2c5.err | tee gpurun_out/bench2_c5.json | cut -c1-200; echo "exit ${PIPESTATUS[0]}"; grep -iE "error|Traceback" gpurun_out/bench2c5.err | head -5
