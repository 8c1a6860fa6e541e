#!/bin/bash
mkdir -p gpurun_out
echo "== dist check (2 GPUs) =="; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -E "rank|Error|error" | head
echo "== bench 2 GPUs C2 graph =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e 2> gpurun_out/bench2.err | cut -c1-330; grep -iE "error|Traceback" gpurun_out/bench2.err | head -5
echo "== bench 2 GPUs C2 eager =="; TLB200_DIST_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e 2> gpurun_out/bench2b.err | cut -c1-330
