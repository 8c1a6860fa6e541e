#!/bin/bash
for dbg in 0 33 1 64 65 129 3 35; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1; done
echo "== R64 =="; for dbg in 0 33 64; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1; done
echo "== grid sweep =="; for g in 148 144 136 128 74; do TLB200_TC_GRID=$g timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1; done
