#!/bin/bash
# refresh of the judged evidence: tests, smoke, bench (both arms), C5 at N=1, launch list, full ncu capture of the MTTKRP kernel
mkdir -p gpurun_out
echo "== gpu tests =="; timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
echo "== smoke =="; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench =="; timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_c2.txt 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-3500 gpurun_out/bench_c2.txt; tail -3 gpurun_out/bench.err
echo "== bench reference arm =="; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1; cut -c1-600 gpurun_out/bench_ref.txt
echo "== bench c5 N=1 =="; timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-e2e --no-cpu --no-refdriver > gpurun_out/bench_c5_n1.txt 2> gpurun_out/bench_c5.err; cut -c1-2500 gpurun_out/bench_c5_n1.txt; tail -2 gpurun_out/bench_c5.err
echo "== launch list =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-refdriver > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
echo "== ncu full =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 2 -o gpurun_out/prof_tc_r1_final python scripts/prof_mttkrp.py 1024 32 2 > gpurun_out/ncu_tc.log 2>&1; echo "exit $?"
