"""Minimal reconstruction / imputation driver for ncu: python scripts/prof_recon.py [rank=32] [impute=0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
impute = len(sys.argv) > 2 and sys.argv[2] == "1"
shape = (512, 1024, 1024)
g = torch.Generator(device="cuda").manual_seed(0)
fs = [torch.rand(s, R, generator=g, device="cuda") for s in shape]
w = torch.ones(R, device="cuda")
out = torch.empty(shape, device="cuda")
if impute:
    x = torch.rand(shape, generator=g, device="cuda")
    mask = (torch.rand(shape, generator=g, device="cuda") > 0.1).float()
for _ in range(3):
    if impute:
        tb.cp_impute(x, mask, (w, fs), out=out)
    else:
        tb.cp_to_tensor((w, fs), out=out)
torch.cuda.synchronize()
print("path", tb.last_kernel_path())
