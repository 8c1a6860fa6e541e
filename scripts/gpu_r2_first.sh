#!/bin/bash
# round-2 first GPU visit: host facts, GPU tests, the new bench (both arms), sanitizer logs
mkdir -p gpurun_out
{ nproc; free -g; lscpu | head -20; nvidia-smi -L; cat /sys/fs/cgroup/memory.max 2>/dev/null; } > gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests.txt
timeout 900 python bench.py > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_c5.json 2> gpurun_out/bench_ref_c5.err; echo "ref rc=$?"
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/sanitize_$tool.txt 2>&1; echo "$tool rc=$?"
done
tail -5 gpurun_out/tests.txt; head -c 1500 gpurun_out/bench_c5_n1.json; tail -3 gpurun_out/bench_c5_n1.err; cat gpurun_out/bench_ref_c5.json; tail -4 gpurun_out/sanitize_*.txt
