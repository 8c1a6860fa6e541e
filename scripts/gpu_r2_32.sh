#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:recon_tc -s 2 -c 1 -o gpurun_out/prof_recon python scripts/prof_recon.py 32 0 > gpurun_out/ncu_recon.log 2>&1; echo "exit $?"
python scripts/ncu_extract.py gpurun_out/prof_recon.ncu-rep | head -30
