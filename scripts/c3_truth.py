"""Which HOOI trajectory is right at C3 (512^3, ranks 64, random init)?  Exact HOOI with an fp64 SVD of the projected
unfolding (projections on the tlb200 TTM kernels) against (a) the own driver at several svd_iters and (b) the
unmodified reference driver with torch's fp32 SVD."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tensorly_b200 as tb
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
from bench import device_slab
n, R, sweeps = int(os.environ.get("N", 512)), int(os.environ.get("R", 64)), 6
shape, ranks = (n, n, n), [R, R, R]
x = device_slab(shape, 0, n, torch.float32, torch.device("cuda"), seed=3)
rs = np.random.RandomState(1); rs.random_sample(ranks)
init = [torch.as_tensor(rs.random_sample((s, r))).cuda().float() for s, r in zip(shape, ranks)]
nx2 = float(tb.sumsq(x))
def exact(dtype):
    fs = [f.clone() for f in init]
    errs = []
    for _ in range(sweeps):
        for k in range(3):
            y = tb.multi_mode_dot(x, fs, skip=k, transpose=True)
            unf = tb.unfold(y, k, contiguous=True).to(dtype)
            u, _, _ = torch.linalg.svd(unf, full_matrices=False)
            fs[k] = u[:, :ranks[k]].float().contiguous()
        core = tb.multi_mode_dot(x, fs, transpose=True)
        errs.append((abs(nx2 - float(tb.sumsq(core))) / nx2) ** 0.5)
    return errs
truth = exact(torch.float64)
print("exact HOOI, fp64 SVD :", ["%.7f" % e for e in truth])
e32 = exact(torch.float32)
print("exact HOOI, fp32 SVD :", ["%.7f" % e for e in e32], "max rel dev vs fp64 %.2e" % max(abs(a - b) / b for a, b in zip(e32, truth)))
for it in (4, 8, 16, 32):
    _, errs = tb.tucker(x, ranks, n_iter_max=sweeps, init="random", random_state=1, tol=0, return_errors=True, svd_iters=it)
    print(f"own driver svd_iters={it:2d}:", ["%.7f" % e for e in errs], "max rel dev vs fp64 %.2e" % max(abs(a - b) / b for a, b in zip(errs, truth)))
try:
    tl = tb.import_tensorly(); tl.set_backend("pytorch"); tb.use()
    from tensorly.decomposition import tucker
    _, errs = tucker(x, ranks, n_iter_max=sweeps, init="random", random_state=1, tol=0, return_errors=True)
    errs = [float(e) for e in errs]
    print("reference driver     :", ["%.7f" % e for e in errs], "max rel dev vs fp64 %.2e" % max(abs(a - b) / b for a, b in zip(errs, truth)))
except Exception as exc:
    print("reference driver unavailable:", exc)
