#!/bin/bash
mkdir -p gpurun_out
echo "== tma_bw =="; timeout 300 ./probes/tma_bw > gpurun_out/tma_bw2.txt 2>&1; echo "exit $?"; grep -E "tma3|tmaC" gpurun_out/tma_bw2.txt
echo "== ncu full on TC kernel =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mttkrp_tc -s 3 -c 3 -o gpurun_out/prof_tc_r1 python scripts/prof_mttkrp.py 1024 32 2 > gpurun_out/ncu_tc.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_tc.log
echo "== launch list =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-refdriver > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_bench.log
ls -la gpurun_out
