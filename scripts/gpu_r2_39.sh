#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cp_update_kernel -s 12 -c 3 -o gpurun_out/prof_cpu python bench.py --workload c2slab --steps 3 --warmup 3 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c4 --no-n4 --no-c2 --no-sustained > gpurun_out/ncu_cpu.log 2>&1; echo "exit $?"
python scripts/ncu_extract.py gpurun_out/prof_cpu.ncu-rep | grep -E "Kernel Name|duration|grid_size|block_size|registers"
