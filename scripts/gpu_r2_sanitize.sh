#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/r2_sanitize_$tool.txt 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/r2_sanitize_$tool.txt
done
