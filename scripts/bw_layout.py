"""Bandwidth of the bit-exact data-movement kernels (unfold / fold / khatri_rao) at C2-like sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(1024, 1024, 1024, generator=g, device="cuda")
for mode in range(3):
    ms = timeit(lambda: tb.unfold(x, mode, contiguous=True))
    print(f"unfold 1024^3 fp32 mode {mode} (contiguous copy): {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s (read+write)")
    ms_t = timeit(lambda: x.movedim(mode, 0).reshape(1024, -1).contiguous() if mode else x.reshape(1024, -1).clone())
    print(f"   torch movedim+reshape copy: {ms_t:.3f} ms")
u = tb.unfold(x, 1, contiguous=True)
ms = timeit(lambda: tb.fold(u, 1, x.shape))
print(f"fold mode 1: {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s")
x4 = x.view(256, 256, 256, 64)
for mode in (1, 2):
    ms = timeit(lambda: tb.unfold(x4, mode, contiguous=True))
    print(f"unfold (256,256,256,64) mode {mode}: {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s")
del u
fs = [torch.rand(1024, 32, generator=g, device="cuda") for _ in range(2)]
w = torch.rand(32, generator=g, device="cuda")
ms = timeit(lambda: tb.khatri_rao(fs, weights=w))
out_bytes = 1024 * 1024 * 32 * 4
print(f"khatri_rao 1024x32 (x) 1024x32 -> 1M x 32: {ms:.3f} ms  {out_bytes/ms/1e6:.0f} GB/s (output)")
fs3 = [torch.rand(256, 64, generator=g, device="cuda") for _ in range(3)]
ms = timeit(lambda: tb.khatri_rao(fs3))
print(f"khatri_rao 3 x (256x64) -> 16.8M x 64 (4.3 GB): {ms:.3f} ms  {256**3*64*4/ms/1e6:.0f} GB/s (output)")
