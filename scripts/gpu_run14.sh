#!/bin/bash
mkdir -p gpurun_out
echo "== dist check (2 GPUs) =="; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -E "rank|Error|error" | head
echo "== bench 2 GPUs C2 =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 2> gpurun_out/bench2.err | tee gpurun_out/bench_c2_n2.txt | cut -c1-1800; tail -3 gpurun_out/bench2.err
echo "== bench 2 GPUs C5 =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --workload c5 --no-e2e 2> gpurun_out/bench2c5.err | tee gpurun_out/bench_c5_n2.txt | cut -c1-1800; tail -3 gpurun_out/bench2c5.err
