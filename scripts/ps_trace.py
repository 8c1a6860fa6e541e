"""Phase clocks of the factoring CTA of tlb200_subspace_iterate (C3 geometry: n = 512, p = 64)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
from tensorly_b200 import _lib
g = torch.Generator(device="cuda").manual_seed(0)
y = torch.rand(512, 4096, generator=g, device="cuda", dtype=torch.float64)
G = y @ y.T
u = tb.orthonormalize(torch.rand(512, 64, generator=g, device="cuda", dtype=torch.float64))
for _ in range(3):
    tb.subspace_iterate(G, u, 4)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); tb.subspace_iterate(G, u, 32); b.record(); torch.cuda.synchronize()
print(f"32 steps: {a.elapsed_time(b) * 1e3 / 32:.1f} us per step")
out = (ctypes.c_longlong * 8)()
lib = _lib.load()
lib.tlb200_debug_subspace_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
lib.tlb200_debug_subspace_trace(out)
t = list(out)
names = ["gemm+gram", "rendezvous+sum+load S", "cholesky"]
print("phase clocks (factoring CTA):", {n: t[i + 1] - t[i] for i, n in enumerate(names)})
print("inside gemm+gram:", {"k loop": t[6] - t[0], "fold + Z out": t[7] - t[6], "gram partial + fence": t[1] - t[7]})
