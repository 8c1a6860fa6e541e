"""Minimal MTTKRP driver for ncu: python scripts/prof_mttkrp.py [n=1024] [rank=32] [reps=2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = torch.Generator(device="cuda").manual_seed(0)
dt = torch.float64 if os.environ.get("DTYPE", "f32") == "f64" else torch.float32
x = torch.rand(n, n, n, generator=g, device="cuda", dtype=dt)
fs = [torch.rand(n, R, generator=g, device="cuda", dtype=dt) for _ in range(3)]
w = torch.ones(R, device="cuda", dtype=dt)
hint = None
if dt == torch.float32 and os.environ.get("TLB200_DISABLE_HF", "0") in ("", "0"):
    hint = tb.RangeHint(x)      # the ALS drivers register it
for _ in range(reps):
    for mode in range(3):
        tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
torch.cuda.synchronize()
print("path", tb.last_kernel_path())
