"""A few TTM (mode_dot) launches for ncu: python scripts/prof_ttm.py [n] [rows_out] [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
J = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(n, n, n, generator=g, device="cuda")
u = torch.randn(n, J, generator=g, device="cuda").t().contiguous().t()     # column-major, as HOOI passes it
for _ in range(6):
    tb.mode_dot(x, u, mode, transpose=True)
torch.cuda.synchronize()
print(tb.last_kernel_path())
