"""Time the MTTKRP call per mode: python scripts/prof_time.py [n] [rank]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(n, n, n, generator=g, device="cuda")
fs = [torch.rand(n, R, generator=g, device="cuda") for _ in range(3)]
w = torch.ones(R, device="cuda")
out = []
for mode in range(3):
    for _ in range(3): tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out.append(f"mode{mode} {ms:.3f} ms {x.numel()*4/ms/1e6:.0f} GB/s")
print(f"n={n} R={R} dbg={os.environ.get('TLB200_TC_DEBUG','0')} path={tb.last_kernel_path()}: " + " | ".join(out))
