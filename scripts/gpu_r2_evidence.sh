#!/bin/bash
# round-2 judged evidence on the final code: GPU tests, smoke, bench (both arms), launch list, full ncu captures
mkdir -p gpurun_out
echo "== gpu tests =="; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests.txt 2>&1; tail -n 3 gpurun_out/r2_tests.txt
echo "== smoke =="; timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.txt 2>&1; tail -n 2 gpurun_out/r2_smoke.txt
echo "== bench (driver command) =="; date +%s > gpurun_out/r2_bench_t0; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "exit $?"; tail -n 3 gpurun_out/r2_bench_n1.err
date +%s > gpurun_out/r2_bench_t1; echo "bench wall: $(( $(cat gpurun_out/r2_bench_t1) - $(cat gpurun_out/r2_bench_t0) )) s"
echo "== bench reference arm =="; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; echo "exit $?"; cut -c1-400 gpurun_out/r2_bench_reference_n1.json
echo "== launch list =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-refdriver --no-c3 --no-c4 --no-n4 --no-fp64 --no-sustained > gpurun_out/r2_ncu_bench.log 2>&1; echo "exit $?"
python scripts/launch_summary.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_summary.txt 2>&1; head -n 12 gpurun_out/r2_launches_summary.txt
echo "== ncu full: MTTKRP C5 rank 64 =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 2 -o gpurun_out/r2_prof_tc_c5 python scripts/prof_mttkrp.py 2048 64 2 > gpurun_out/r2_ncu_tc.log 2>&1; echo "exit $?"
python scripts/ncu_extract.py gpurun_out/r2_prof_tc_c5.ncu-rep > gpurun_out/r2_tc_stream_c5_ncu_full.txt 2>&1; head -n 20 gpurun_out/r2_tc_stream_c5_ncu_full.txt
echo "== ncu full: MTTKRP C2 rank 32 =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 1 -o gpurun_out/r2_prof_tc_c2 python scripts/prof_mttkrp.py 1024 32 2 > gpurun_out/r2_ncu_tc2.log 2>&1; echo "exit $?"
python scripts/ncu_extract.py gpurun_out/r2_prof_tc_c2.ncu-rep > gpurun_out/r2_tc_stream_c2_ncu_full.txt 2>&1
echo "== ncu full: MTTKRP C5 rank 64, 3xTF32 engine (no range hint) =="
TLB200_DISABLE_HF=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 1 -o gpurun_out/r2_prof_tc_c5_tf32 python scripts/prof_mttkrp.py 2048 64 2 > gpurun_out/r2_ncu_tc3.log 2>&1; echo "exit $?"
python scripts/ncu_extract.py gpurun_out/r2_prof_tc_c5_tf32.ncu-rep > gpurun_out/r2_tc_stream_c5_tf32_ncu_full.txt 2>&1
echo "== power triage (sustained, 2 s per variant) =="
timeout 300 python scripts/power_triage.py 2>&1 | grep -E "HFoff|ceiling" > gpurun_out/r2_power_triage_final.txt; cat gpurun_out/r2_power_triage_final.txt
