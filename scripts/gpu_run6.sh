#!/bin/bash
mkdir -p gpurun_out
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -8 gpurun_out/tests.txt
for dbg in 0 1 2 3; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1; done
for dbg in 0 1 2 3; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1; done
echo "== tc_check C2 =="; timeout 600 python scripts/tc_check.py 1024 32 uniform 2>&1 | tail -4
echo "== tc_check 512 R64 randn =="; timeout 300 python scripts/tc_check.py 512 64 randn 2>&1 | tail -4
echo "== flush sweep =="; for fl in 4 16 32; do TLB200_TC_FLUSH=$fl timeout 300 python scripts/tc_check.py 512 32 uniform 2>&1 | tail -1; done
echo "== ttm_check =="; timeout 600 python scripts/ttm_check.py 512 64 2>&1 | grep -E "auto|chain"
echo "== bench =="; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
