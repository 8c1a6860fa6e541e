#!/bin/bash
echo "== mma rate =="; timeout 300 ./probes/mma_rate 2>&1 | tail -30
echo "== ttm diag (default grid) =="; timeout 300 python scripts/ttm_diag.py 2>&1 | tail -12
echo "== ttm diag grid=74 =="; TLB200_TC_GRID=74 timeout 300 python scripts/ttm_diag.py 2>&1 | grep -E "^L=" 
echo "== ttm diag grid=1024 =="; TLB200_TC_GRID=1024 timeout 300 python scripts/ttm_diag.py 2>&1 | grep -E "^L="
echo "== ttm diag flush=64 =="; TLB200_TC_FLUSH=64 timeout 300 python scripts/ttm_diag.py 2>&1 | grep -E "^L="
echo "== ttm diag no-convert-stores (timing only) dbg=8 =="; TLB200_TC_DEBUG=8 timeout 300 python scripts/ttm_diag.py 2>&1 | grep -E "^L=" | head -2
