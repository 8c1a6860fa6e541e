#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 3 -o gpurun_out/prof_hf_c5 python scripts/prof_mttkrp.py 2048 64 2 > gpurun_out/ncu_hf.log 2>&1; echo "exit $?"
python scripts/ncu_extract.py gpurun_out/prof_hf_c5.ncu-rep > gpurun_out/hf_c5_ncu.txt 2>&1; cat gpurun_out/hf_c5_ncu.txt | head -60
