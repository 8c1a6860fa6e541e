#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 3 7 15; do
  echo "== C2 debug mask $dbg =="; TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1
done
echo "== R64 512 =="; for dbg in 0 1 2 7; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1; done
echo "== ncu full on TC kernel =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 3 -o gpurun_out/prof_tc_r1b python scripts/prof_mttkrp.py 1024 32 2 > gpurun_out/ncu_tc.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_tc.log
