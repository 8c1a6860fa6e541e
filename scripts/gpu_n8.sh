#!/bin/bash
# 8-GPU bench lines (the driver's command for C2, and the C5 workload)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 100 $TR --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 2> gpurun_out/bench8.err > gpurun_out/bench8_c2.json; echo "c2 exit $?"; cut -c1-170 gpurun_out/bench8_c2.json
timeout 100 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --workload c5 --no-e2e 2> gpurun_out/bench8c5.err > gpurun_out/bench8_c5.json; echo "c5 exit $?"; cut -c1-170 gpurun_out/bench8_c5.json
