#!/bin/bash
mkdir -p gpurun_out
python scripts/ps_trace.py > gpurun_out/ps_trace.txt 2>&1
python scripts/c3_truth.py > gpurun_out/c3_truth.txt 2>&1
N=256 R=32 python scripts/c3_truth.py > gpurun_out/c3_truth_256.txt 2>&1
cat gpurun_out/ps_trace.txt gpurun_out/c3_truth.txt gpurun_out/c3_truth_256.txt
