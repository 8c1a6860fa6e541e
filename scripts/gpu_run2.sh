#!/bin/bash
mkdir -p gpurun_out
echo "== tma_bw =="; timeout 300 ./probes/tma_bw > gpurun_out/tma_bw.txt 2>&1; echo "exit $?"; cat gpurun_out/tma_bw.txt
echo "== tc_check small =="; timeout 300 python scripts/tc_check.py 256 32 uniform > gpurun_out/tc_check.txt 2>&1; echo "exit $?"; tail -8 gpurun_out/tc_check.txt
for fl in 4 8 16 64 100000; do
  echo "== tc_check 512 flush $fl =="; TLB200_TC_FLUSH=$fl timeout 300 python scripts/tc_check.py 512 32 uniform 2>&1 | tail -4
done
echo "== tc_check randn/R64 =="; timeout 300 python scripts/tc_check.py 512 64 randn 2>&1 | tail -4
echo "== tc_check C2 =="; timeout 600 python scripts/tc_check.py 1024 32 uniform 2>&1 | tail -4
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -15 gpurun_out/tests.txt
echo "== bench =="; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
