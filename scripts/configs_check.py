"""Timing + sanity of the BASELINE configs C2..C4 on one GPU, next to the reference's own GPU path
(unmodified TensorLy, pytorch backend, `core` tenalg = cuBLAS + materialised KR + unfold copy)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
tl = tb.import_tensorly()
tl.set_backend("pytorch")
from tensorly.cp_tensor import CPTensor
from tensorly.decomposition import parafac, non_negative_parafac, tucker

def rel(a, b): return float(torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double()))
def sync_time(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r
def sweeps_per_s(run, n0, n1):
    run(n0); ta, _ = sync_time(lambda: run(n0)); tb_, _ = sync_time(lambda: run(n1)); return (n1 - n0) / max(tb_ - ta, 1e-9)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device="cuda").manual_seed(0)

if which in ("all", "c2"):
    x = torch.rand(1024, 1024, 1024, generator=g, device="cuda"); R = 32
    fs = [torch.rand(1024, R, generator=g, device="cuda") for _ in range(3)]
    def run_ref(n):
        init = CPTensor((torch.ones(R, device="cuda"), [f.clone() for f in fs]))
        return parafac(x, R, n_iter_max=n, init=init, tol=0, return_errors=True)
    tl.tenalg.set_backend("core")
    print("C2 reference GPU path (core tenalg, cuBLAS): %.1f sweeps/s" % sweeps_per_s(run_ref, 2, 6), flush=True)
    tb.use()
    print("C2 unmodified parafac on b200 tenalg:          %.1f sweeps/s" % sweeps_per_s(run_ref, 2, 12), flush=True)
    def run_ours(n): return tb.parafac(x, R, n_iter_max=n, init=(None, fs), tol=0, return_errors=True)
    print("C2 tensorly_b200.parafac (CUDA graph):         %.1f sweeps/s" % sweeps_per_s(run_ours, 3, 43), flush=True)
    del x

if which in ("all", "c3"):
    x = torch.rand(512, 512, 512, generator=g, device="cuda")
    # TTM chains first: after the SVD runs the caching allocator is fragmented and the 67 MB outputs of every
    # chain go through cudaMalloc (a 40x timing artifact that has nothing to do with the kernels)
    us = [torch.randn(512, 64, generator=g, device="cuda").t().contiguous().t() for _ in range(3)]
    tb.register()
    for name, be in (("core", "core"), ("b200", "b200")):
        tl.tenalg.set_backend(be)
        f = lambda: [tl.tenalg.multi_mode_dot(x, us, skip=k, transpose=True) for k in range(3)] + [tl.tenalg.multi_mode_dot(x, us, transpose=True)]
        for _ in range(5): f()           # warm the caching allocator: the first calls pay cudaMalloc for the outputs
        t, _ = sync_time(lambda: [f() for _ in range(10)])
        print(f"C3 TTM chains of one HOOI sweep (3 skip + 1 full) on {name}: {t/10*1e3:.3f} ms  ({77.5/(t/10)/1e3:.1f} TFLOP/s useful)", flush=True)
    def run_t(n): return tucker(x, [64, 64, 64], n_iter_max=n, init="random", random_state=1, tol=0)
    tl.tenalg.set_backend("core")
    print("C3 tucker reference GPU path (core):  %.2f sweeps/s" % sweeps_per_s(run_t, 1, 3), flush=True)
    tb.use()
    print("C3 tucker on b200 tenalg:             %.2f sweeps/s" % sweeps_per_s(run_t, 1, 3), flush=True)
    tb.use_gram_svd()
    print("C3 tucker on b200 tenalg + gram_svd:  %.2f sweeps/s" % sweeps_per_s(run_t, 1, 5), flush=True)
    tb.use_default_svd()
    del x

if which in ("all", "c4"):
    R = 64
    x = torch.rand(256, 256, 256, 256, generator=g, device="cuda")
    fs = [torch.rand(256, R, generator=g, device="cuda") for _ in range(4)]
    w = torch.ones(R, device="cuda")
    for mode in range(4):
        tb.set_kernel_path("auto"); out = tb.unfolding_dot_khatri_rao(x, (w, fs), mode); path = tb.last_kernel_path()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3): tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
        e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 3
        tb.set_kernel_path("simt"); ref = tb.unfolding_dot_khatri_rao(x, (w, fs), mode); tb.set_kernel_path("auto")
        print(f"C4 MTTKRP mode {mode}: path {path} {ms:.3f} ms {x.numel()*4/ms/1e6:.0f} GB/s  err vs SIMT fp32 {rel(out, ref):.2e}", flush=True)
    from tensorly_b200.cp_als import CPALS
    for dimtree in (True, False):
        st = CPALS(x, w, fs, update="mu", dimtree=dimtree)
        for _ in range(3): st.sweep(False)
        t, _ = sync_time(lambda: [st.sweep(False) for _ in range(10)])
        print("C4 tensorly_b200 NN-CP sweeps (CPALS, update='mu', %s): %.2f sweeps/s" % ("dimension tree" if dimtree else "one MTTKRP per mode", 10 / t), flush=True)
        del st
    tb.use()
    def run_nn_ref(n):
        init = CPTensor((torch.ones(R, device="cuda"), [f.clone() for f in fs]))
        return non_negative_parafac(x, R, n_iter_max=n, init=init, tol=0)
    print("C4 unmodified non_negative_parafac on b200 tenalg: %.2f sweeps/s" % sweeps_per_s(run_nn_ref, 1, 4), flush=True)
