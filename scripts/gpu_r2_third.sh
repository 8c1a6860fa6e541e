#!/bin/bash
mkdir -p gpurun_out
./probes/bf16_cross > gpurun_out/bf16_cross.txt 2>&1; echo "probe rc=$?"
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "orthonormalize or hals or tucker or gram_update or parafac or from_ttm or mttkrp_vs_oracle" > gpurun_out/tests3.txt 2>&1; echo "tests rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3.csv python scripts/prof_c3.py > gpurun_out/prof_c3.log 2>&1; echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c3.csv > gpurun_out/launches_c3_summary.txt 2>&1
cat gpurun_out/bf16_cross.txt; tail -n 15 gpurun_out/tests3.txt; cat gpurun_out/launches_c3_summary.txt
