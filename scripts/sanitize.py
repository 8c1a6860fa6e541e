"""Small tcgen05-path MTTKRP / TTM / ALS runs for compute-sanitizer (memcheck, racecheck, synccheck):

    compute-sanitizer --tool memcheck python scripts/sanitize.py

Covers the three X layouts of tc_stream_kernel at rank 32 and 64 (two-line K-major, one-line K-major,
m-contiguous), the TTM engine (T >= 32 and T == 1), the dimension-tree kernels, the fused solve and one CUDA-graphed
ALS sweep.  Results are checked against torch fp64 so that a sanitizer-clean but wrong run still fails."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb

torch.manual_seed(0)
paths = set()
for shape in [(128, 64, 96), (64, 36, 100)]:
    for R in (32, 64):
        x = torch.randn(shape, device="cuda")
        fs = [torch.randn(s, R, device="cuda") for s in shape]
        w = torch.rand(R, device="cuda") + 0.5
        letters = "ijk"
        for mode in range(3):
            got = tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
            paths.add(tb.last_kernel_path())
            ops = [fs[m].double() for m in range(3) if m != mode]
            sub = ",".join(f"{letters[m]}r" for m in range(3) if m != mode)
            ref = torch.einsum(f"ijk,{sub}->{letters[mode]}r", x.double(), *ops) * w.double()
            err = float(torch.linalg.norm(got.double() - ref) / torch.linalg.norm(ref))
            assert err < 1e-5, (shape, R, mode, err)
        t = tb.mode_dot(x, fs[2], 2, transpose=True)
        paths.add(tb.last_kernel_path())
        ref = torch.einsum("ijk,kr->ijr", x.double(), fs[2].double())
        assert float(torch.linalg.norm(t.double() - ref) / torch.linalg.norm(ref)) < 1e-5
        y = tb.mode_dot(x, fs[1], 1, transpose=True)
        ref = torch.einsum("ijk,jr->irk", x.double(), fs[1].double())
        assert float(torch.linalg.norm(y.double() - ref) / torch.linalg.norm(ref)) < 1e-5
        for mode in range(2):
            a = tb.mttkrp_from_ttm(t, (w, fs), mode)
            b = tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
            assert float(torch.linalg.norm(a - b) / torch.linalg.norm(b)) < 1e-5
x = torch.rand((128, 96, 160), device="cuda")
fs = [torch.rand(s, 32, device="cuda") for s in x.shape]
cp, errs = tb.parafac(x, 32, n_iter_max=4, init=(None, fs), tol=0, return_errors=True)
torch.cuda.synchronize()
print("sanitize run ok; kernel paths:", sorted(paths), "errs", errs)
