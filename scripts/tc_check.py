"""GPU check of the tcgen05 MTTKRP path: error vs an fp64 evaluation and vs the SIMT fp32
path, and per-mode bandwidth.  usage: python scripts/tc_check.py [n=512] [rank=32] [dist=uniform|randn]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dist = sys.argv[3] if len(sys.argv) > 3 else "uniform"
shape = (n, n, n) if len(sys.argv) <= 4 else tuple(int(s) for s in sys.argv[4].split("x"))
g = torch.Generator(device="cuda").manual_seed(0)
mk = (lambda *s: torch.rand(*s, generator=g, device="cuda")) if dist == "uniform" else (lambda *s: torch.randn(*s, generator=g, device="cuda"))
x = mk(*shape)
fs = [mk(s, R) for s in shape]
w = torch.rand(R, generator=g, device="cuda") + 0.5
x64 = x.double(); fs64 = [f.double() for f in fs]; w64 = w.double()
def rel(a, b): return float(torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double()))
print(f"shape {shape} rank {R} {dist} flush={os.environ.get('TLB200_TC_FLUSH','default')}")
for mode in range(len(shape)):
    tb.set_kernel_path("simt")
    truth = tb.unfolding_dot_khatri_rao(x64, (w64, fs64), mode)
    simt = tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
    tb.set_kernel_path("auto")
    tc = tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
    path = tb.last_kernel_path()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = x.numel() * 4 / ms / 1e6
    print(f"  mode {mode}: path={path:8s} err(tc,fp64)={rel(tc, truth):.3e} err(simt,fp64)={rel(simt, truth):.3e} "
          f"err(tc,simt)={rel(tc, simt):.3e}  {ms:8.3f} ms  {gbs:8.1f} GB/s")
