#!/bin/bash
# First GPU pass: probes, parity tests, a short bench, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== probe ==" ; timeout 300 ./probes/tc_probe > gpurun_out/probe.txt 2>&1; echo "probe exit $?"; cat gpurun_out/probe.txt
echo "== smoke =="; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.txt
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -30 gpurun_out/tests.txt
echo "== bench =="; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
