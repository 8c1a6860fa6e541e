#!/bin/bash
mkdir -p gpurun_out
DTYPE=f64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_gemm_dmma -s 3 -c 3 -o gpurun_out/prof_dmma python scripts/prof_mttkrp.py 512 32 2 > gpurun_out/ncu_dmma.log 2>&1; echo "exit $?"
python scripts/ncu_extract.py gpurun_out/prof_dmma.ncu-rep | head -60
