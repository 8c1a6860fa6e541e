"""from_ttm kernels for ncu: python scripts/prof_from_ttm.py [n] [rank]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(n, n, n, generator=g, device="cuda")
fs = [torch.rand(n, R, generator=g, device="cuda") for _ in range(3)]
t = tb.mode_dot(x, fs[2], 2, transpose=True)
for _ in range(4):
    for mode in range(2):
        tb.mttkrp_from_ttm(t, (None, fs), mode)
torch.cuda.synchronize()
