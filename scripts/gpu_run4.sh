#!/bin/bash
mkdir -p gpurun_out
echo "== tc_check small =="; timeout 300 python scripts/tc_check.py 256 32 uniform 2>&1 | tail -4
echo "== tc_check 512 R64 randn =="; timeout 300 python scripts/tc_check.py 512 64 randn 2>&1 | tail -4
echo "== tc_check odd (kmajor_1) =="; timeout 300 python scripts/tc_check.py 0 32 uniform 200x300x404 2>&1 | tail -4
echo "== tc_check C2 =="; timeout 600 python scripts/tc_check.py 1024 32 uniform 2>&1 | tail -4
echo "== ttm_check =="; timeout 600 python scripts/ttm_check.py 512 64 2>&1 | tail -20
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -15 gpurun_out/tests.txt
echo "== bench =="; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
