#!/bin/bash
mkdir -p gpurun_out
echo "== bench 2 GPUs C2 graph =="; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e 2> gpurun_out/bench2.err | cut -c1-330; echo "exit ${PIPESTATUS[0]}"; grep -iE "error|Traceback" gpurun_out/bench2.err | head -3
