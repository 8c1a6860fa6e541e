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
for shape, hinted in [((128, 64, 96), False), ((64, 36, 100), False), ((128, 64, 96), True), ((96, 64, 128), True)]:
    for R in (32, 64):
        x = torch.randn(shape, device="cuda")
        hint = tb.RangeHint(x) if hinted else None        # the fp16-split engine (64-element tiles: inner extent % 32 == 0)
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
        if hint is not None:
            hint.close()
# fp64 on DMMA (cp.async and register-staged variants: even / odd strides)
for shape in [(64, 48, 96), (33, 47, 51)]:
    x = torch.randn(shape, device="cuda", dtype=torch.float64)
    fs = [torch.randn(s, 32, device="cuda", dtype=torch.float64) for s in shape]
    for mode in range(3):
        got = tb.unfolding_dot_khatri_rao(x, (None, fs), mode)
        paths.add(tb.last_kernel_path())
        ops = [fs[m] for m in range(3) if m != mode]
        sub = ",".join(f"{'ijk'[m]}r" for m in range(3) if m != mode)
        ref = torch.einsum(f"ijk,{sub}->{'ijk'[mode]}r", x, *ops)
        assert float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref)) < 1e-12
# reconstruction / imputation on the tensor cores: TMA epilogue with two-line and one-line boxes, two contraction chunks,
# and the transposing epilogue for extents TMA cannot describe
for shape, R in [((256, 64, 64), 32), ((256, 66, 62), 24), ((192, 64, 96), 48), ((192, 75, 73), 20), ((128, 16, 24, 32), 40)]:
    fs = [torch.randn(s, R, device="cuda") for s in shape]
    w = torch.rand(R, device="cuda") + 0.5
    sub = ",".join(f"{'ijkl'[m]}r" for m in range(len(shape)))
    ref = torch.einsum(sub + "->" + "ijkl"[:len(shape)], fs[0].double() * w.double(), *[f.double() for f in fs[1:]])
    rec = tb.cp_to_tensor((w, fs))
    paths.add("recon-" + tb.last_kernel_path())
    assert float(torch.linalg.norm(rec.double() - ref) / torch.linalg.norm(ref)) < 1e-5, shape
    x = torch.randn(shape, device="cuda")
    mask = (torch.rand(shape, device="cuda") > 0.3).float()
    new, stats = tb.cp_impute(x, mask, (w, fs))
    want = x.double() * mask.double() + ref * (1 - mask.double())
    assert float(torch.linalg.norm(new.double() - want) / torch.linalg.norm(want)) < 1e-5, shape
    assert abs(float(stats[1]) - float((want ** 2).sum())) <= 1e-4 * float((want ** 2).sum())
    got = tb.cp_to_tensor((w, fs), mask=mask)
    assert float(torch.linalg.norm(got.double() - ref * mask.double()) / torch.linalg.norm(ref * mask.double())) < 1e-5
# HOOI power step (cp.async + DMMA + register Cholesky) and the own Tucker driver
y = torch.rand(96, 512, device="cuda", dtype=torch.float64)
u = tb.subspace_iterate(y @ y.T, tb.orthonormalize(torch.rand(96, 16, device="cuda", dtype=torch.float64)), 3)
# one-pass Cholesky QR: the defect is cond(Z)^2 x eps (1e-9..1e-8 here); the drivers finish with a two-pass orthonormalize
assert float(torch.linalg.norm(u.T @ u - torch.eye(16, device="cuda", dtype=torch.float64))) < 1e-6
u = tb.orthonormalize(u)
assert float(torch.linalg.norm(u.T @ u - torch.eye(16, device="cuda", dtype=torch.float64))) < 1e-10
tb.tucker(torch.rand((48, 40, 56), device="cuda"), [8, 8, 8], n_iter_max=2, init="random", random_state=1, tol=0)
x = torch.rand((128, 96, 160), device="cuda")
fs = [torch.rand(s, 32, device="cuda") for s in x.shape]
cp, errs = tb.parafac(x, 32, n_iter_max=4, init=(None, fs), tol=0, return_errors=True)
torch.cuda.synchronize()
print("sanitize run ok; kernel paths:", sorted(paths), "errs", errs)
