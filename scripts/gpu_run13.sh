#!/bin/bash
mkdir -p gpurun_out
echo "== bench =="; timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_c2.txt 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-2500 gpurun_out/bench_c2.txt; tail -3 gpurun_out/bench.err
echo "== bench reference arm =="; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1; cut -c1-900 gpurun_out/bench_ref.txt
echo "== launch list =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-refdriver > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
echo "== ncu full =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 2 -o gpurun_out/prof_tc_r1_final python scripts/prof_mttkrp.py 1024 32 2 > gpurun_out/ncu_tc.log 2>&1; echo "exit $?"
echo "== tc_check =="; timeout 600 python scripts/tc_check.py 1024 32 uniform 2>&1 | tail -3
timeout 300 python scripts/tc_check.py 512 64 randn 2>&1 | tail -3
