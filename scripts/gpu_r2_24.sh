#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ttm_triage.py > gpurun_out/ttm_triage.txt 2>&1; cat gpurun_out/ttm_triage.txt | tail -8
