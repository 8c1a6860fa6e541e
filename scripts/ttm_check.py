"""GPU check of the tcgen05 TTM path: error vs fp64 and throughput.  usage: python scripts/ttm_check.py [n=512] [I=64]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
I = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, n, n, generator=g, device="cuda")
us = [torch.randn(n, I, generator=g, device="cuda").t().contiguous().t() for _ in range(3)]   # column-major (n, I)
def rel(a, b): return float(torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double()))
print(f"TTM {n}^3 x ({I},{n}) flush={os.environ.get('TLB200_TC_FLUSH','default')}")
for mode in range(3):
    truth = torch.tensordot(us[mode].double().t(), x.double(), dims=([1], [mode])).movedim(0, mode)
    tb.set_kernel_path("simt"); simt = tb.mode_dot(x, us[mode], mode, transpose=True)
    tb.set_kernel_path("auto"); tc = tb.mode_dot(x, us[mode], mode, transpose=True); path = tb.last_kernel_path()
    for nm, fn in (("simt", "simt"), ("auto", "auto")):
        tb.set_kernel_path(fn)
        for _ in range(2): tb.mode_dot(x, us[mode], mode, transpose=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): tb.mode_dot(x, us[mode], mode, transpose=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  mode {mode} {nm:5s}({tb.last_kernel_path():7s}) {ms:7.3f} ms  {2*I*x.numel()/ms/1e9:7.1f} TFLOP/s  {x.numel()*4/ms/1e6:7.1f} GB/s", end="")
        print(f"  err(vs fp64) {rel(tc if nm=='auto' else simt, truth):.2e}")
    del truth
tb.set_kernel_path("auto")
for skip in (None, 0, 1, 2):
    ref = x.double()
    for m in range(3):
        if m != skip: ref = torch.tensordot(us[m].double().t(), ref, dims=([1], [m])).movedim(0, m)
    out = tb.multi_mode_dot(x, us, skip=skip, transpose=True)
    for _ in range(2): tb.multi_mode_dot(x, us, skip=skip, transpose=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): tb.multi_mode_dot(x, us, skip=skip, transpose=True)
    e1.record(); torch.cuda.synchronize()
    print(f"  chain skip={skip}: {e0.elapsed_time(e1)/10:7.3f} ms  err {rel(out, ref):.2e}  path {tb.last_kernel_path()}")
# the bar to beat: torch (cuBLAS) on the same B200, as the reference's core tenalg would run it
for mode in range(3):
    xm = x.movedim(mode, 0).reshape(n, -1)
    for _ in range(2): (us[mode].t() @ xm)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): (us[mode].t() @ x.movedim(mode, 0).reshape(n, -1))
    e1.record(); torch.cuda.synchronize()
    print(f"  torch unfold+matmul mode {mode}: {e0.elapsed_time(e1)/10:7.3f} ms")
