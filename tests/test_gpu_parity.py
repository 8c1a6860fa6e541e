"""GPU parity tests: the CUDA path (called through the C ABI) against the committed
reference fixtures (tests/golden, generated from the real TensorLy) and against the CPU
oracle on the same seeded inputs.

Gates (BASELINE.json north_star): unfold / khatri_rao bit-exact; MTTKRP / TTM relative
Frobenius error <= 1e-5 (fp32) / 1e-12 (fp64) against the numpy `core` result in the same
dtype; ALS reconstruction error within 1e-4 (relative) after a fixed number of sweeps.
"""
import itertools

import os

import numpy as np
import pytest
import torch

import tensorly_b200 as tb
from oracle import oracle as O
from conftest import rel_fro

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}
PATHS = ["simt", "auto"]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _reset_path():
    tb.set_kernel_path("auto")
    yield
    tb.set_kernel_path("auto")


# --------------------------------------------------------------------------- unfold / fold
def test_unfold_fold_golden(golden):
    g = golden("unfold")
    for case in g.cases():
        x = g[f"{case}/x"]
        xd = dev(x)
        for mode in range(x.ndim):
            u = tb.unfold(xd, mode, contiguous=True)
            assert u.is_contiguous()
            assert np.array_equal(host(u), g[f"{case}/unfold{mode}"])
            assert np.array_equal(host(tb.fold(u, mode, x.shape)), x)
            assert np.array_equal(host(tb.unfold(xd, mode)), g[f"{case}/unfold{mode}"])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(37, 65, 129), (64, 3, 5, 40), (5, 300, 1), (1, 17, 33), (2, 2, 2, 2, 2, 2, 3, 4),
                                   (130, 257), (96, 32, 64), (33, 40, 7)])
def test_unfold_fold_bit_exact(shape, dtype):
    rng = np.random.RandomState(0)
    x = rng.standard_normal(shape).astype(dtype)
    xd = dev(x)
    for mode in range(len(shape)):
        u = tb.unfold(xd, mode, contiguous=True)
        assert np.array_equal(host(u), np.ascontiguousarray(O.unfold(x, mode)))
        assert np.array_equal(host(tb.fold(u, mode, shape)), x)
    assert tb.last_kernel_path() == "copy"


def test_unfold_noncontiguous_input_and_negative_mode():
    x = np.random.RandomState(1).random_sample((6, 7, 8)).astype(np.float32)
    xd = dev(x).permute(2, 0, 1)          # non-contiguous view
    xn = np.transpose(x, (2, 0, 1))
    for mode in (-1, 0, 1):
        assert np.array_equal(host(tb.unfold(xd, mode)), O.unfold(xn, mode % 3))


# --------------------------------------------------------------------------- khatri_rao
def test_khatri_rao_golden_bit_exact(golden):
    g = golden("khatri_rao")
    for case in g.cases():
        mats = [dev(m) for m in g.arrays(case, "m")]
        w = dev(g[f"{case}/w"]) if g.has(f"{case}/w") else None
        mask = dev(g[f"{case}/mask"]) if g.has(f"{case}/mask") else None
        skip = int(g[f"{case}/skip"])
        out = tb.khatri_rao(mats, weights=w, skip_matrix=None if skip < 0 else skip, mask=mask)
        ref = g[f"{case}/out"]
        assert host(out).dtype == ref.dtype
        assert np.array_equal(host(out), ref), case


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_khatri_rao_bit_exact_vs_oracle(dtype):
    rng = np.random.RandomState(3)
    for rows, rank in (((1024, 64), 32), ((17, 9, 33), 37), ((5, 4, 3, 2, 6), 5), ((300, 1, 7), 64), ((2, 2), 1),
                       ((40, 50, 37), 32), ((9, 700), 64), ((3, 5, 7, 600), 8), ((515, 19), 6), ((64, 64, 16), 128)):
        mats = [(rng.standard_normal((r, rank))).astype(dtype) for r in rows]
        w = rng.standard_normal(rank).astype(dtype)
        for weights in (None, w):
            out = tb.khatri_rao([dev(m) for m in mats], weights=None if weights is None else dev(weights))
            assert np.array_equal(host(out), O.khatri_rao(mats, weights=weights))
        # column-major (strided) inputs, e.g. U straight out of an SVD
        strided = [dev(np.ascontiguousarray(m.T)).T for m in mats]
        assert not strided[0].is_contiguous() or strided[0].shape[1] == 1 or strided[0].shape[0] == 1
        assert np.array_equal(host(tb.khatri_rao(strided, weights=dev(w))), O.khatri_rao(mats, weights=w))
        total = int(np.prod(rows))
        mask = (rng.random_sample((total, 1)) > 0.3).astype(dtype) * dtype(1.5)
        got = tb.khatri_rao([dev(m) for m in mats], weights=dev(w), mask=dev(mask))
        assert np.array_equal(host(got), O.khatri_rao(mats, weights=w, mask=mask))


def test_khatri_rao_reference_semantics():
    rng = np.random.RandomState(4)
    mats = [dev(rng.random_sample((n, 3)).astype(np.float32)) for n in (4, 5, 2)]
    # single remaining matrix is returned as is, weights ignored (_khatri_rao.py:68-69)
    assert tb.khatri_rao([mats[0]], weights=dev(np.array([2.0, 2.0, 2.0], dtype=np.float32))) is mats[0]
    assert tb.khatri_rao(mats[:2], skip_matrix=1) is mats[0]
    with pytest.raises(ValueError):
        tb.khatri_rao([mats[0], dev(rng.random_sample((3, 4)).astype(np.float32))])
    with pytest.raises(ValueError):
        tb.khatri_rao([mats[0], dev(rng.random_sample((3, 3, 2)).astype(np.float32))])
    with pytest.warns(UserWarning):
        out = tb.khatri_rao([dev(np.arange(3, dtype=np.float32)), dev(np.arange(4, dtype=np.float32))])
    assert np.array_equal(host(out), np.outer(np.arange(3), np.arange(4)).reshape(-1, 1).astype(np.float32))


# --------------------------------------------------------------------------- MTTKRP
@pytest.mark.parametrize("path", PATHS)
def test_mttkrp_golden(golden, path):
    tb.set_kernel_path(path)
    g = golden("mttkrp")
    for case in g.cases():
        x = g[f"{case}/x"]
        fs = g.arrays(case, "f")
        w = g[f"{case}/w"] if g.has(f"{case}/w") else None
        xd, fd = dev(x), [dev(f) for f in fs]
        wd = None if w is None else dev(w)
        for mode in range(x.ndim):
            out = tb.unfolding_dot_khatri_rao(xd, (wd, fd), mode)
            assert out.shape == (x.shape[mode], fs[0].shape[1])
            err = rel_fro(host(out), g[f"{case}/out{mode}"])
            assert err <= TOL[x.dtype], (case, mode, err)


MTTKRP_CASES = [
    # shape, rank
    ((64, 48, 80), 32),
    ((128, 128, 128), 32),
    ((130, 70, 45), 10),
    ((256, 64, 96), 64),
    ((33, 257, 19), 7),
    ((40, 24, 20, 12), 16),
    ((12, 10, 9, 8, 6), 5),
    ((300, 200), 12),
    ((96, 160, 64), 100),
    ((128, 96, 160), 130),       # three column blocks on the tensor-core path
    ((16, 16, 4096), 32),
    ((4096, 16, 16), 32),
    ((48, 1, 40), 8),
    ((24, 20, 1, 36), 12),
]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,rank", MTTKRP_CASES)
def test_mttkrp_vs_oracle(shape, rank, dtype, path):
    tb.set_kernel_path(path)
    rng = np.random.RandomState(hash((shape, rank)) % (2 ** 31))
    x = rng.random_sample(shape).astype(dtype)
    fs = [rng.random_sample((s, rank)).astype(dtype) for s in shape]
    w = (rng.random_sample(rank) + 0.5).astype(dtype)
    xd, fd, wd = dev(x), [dev(f) for f in fs], dev(w)
    for mode in range(len(shape)):
        ref = O.unfolding_dot_khatri_rao(x, (w, fs), mode)
        out = host(tb.unfolding_dot_khatri_rao(xd, (wd, fd), mode))
        err = rel_fro(out, ref)
        assert err <= TOL[np.dtype(dtype)], (shape, rank, mode, tb.last_kernel_path(), err)
        assert tb.last_kernel_path() in ("simt", "dmma", "tcgen05", "tcgen05-f16")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,rank", [((64, 48, 80), 32), ((130, 70, 45), 10), ((256, 64, 96), 64), ((33, 257, 19), 7),
                                        ((40, 24, 20, 12), 16), ((12, 10, 9, 8, 6), 5), ((96, 160, 64), 100),
                                        ((512, 256, 128), 32),
                                        # singleton lead modes: the inner range collapses to one row
                                        ((48, 1, 40), 8), ((24, 20, 1, 36), 12), ((1, 30, 28), 6), ((20, 1, 1, 16), 4)])
def test_mttkrp_from_ttm_vs_oracle(shape, rank, dtype):
    """Dimension-tree reuse: the MTTKRP of every mode before the last, from T = X x_{N-1} F_{N-1}^T, meets the
    same gate against the oracle's full MTTKRP."""
    rng = np.random.RandomState(hash((shape, rank, 7)) % (2 ** 31))
    x = rng.random_sample(shape).astype(dtype)
    fs = [rng.random_sample((s, rank)).astype(dtype) for s in shape]
    w = (rng.random_sample(rank) + 0.5).astype(dtype)
    xd, wd = dev(x), dev(w)
    fd = [dev(f) for f in fs]
    fd[0] = dev(np.ascontiguousarray(fs[0].T)).T          # a column-major factor
    t = tb.mode_dot(xd, fd[-1], len(shape) - 1, transpose=True)
    assert tuple(t.shape) == tuple(shape[:-1]) + (rank,)
    for mode in range(len(shape) - 1):
        ref = O.unfolding_dot_khatri_rao(x, (w, fs), mode)
        out = host(tb.mttkrp_from_ttm(t, (wd, fd), mode))
        err = rel_fro(out, ref)
        assert err <= TOL[np.dtype(dtype)], (shape, rank, mode, err)
        out_now = host(tb.mttkrp_from_ttm(t, (None, fd), mode))
        assert rel_fro(out_now, O.unfolding_dot_khatri_rao(x, (None, fs), mode)) <= TOL[np.dtype(dtype)]
    with pytest.raises(ValueError):
        tb.mttkrp_from_ttm(t, (wd, fd), len(shape) - 1)
    with pytest.raises(ValueError):
        tb.mttkrp_from_ttm(t, (wd, fd[:-1]), 0)


@pytest.mark.parametrize("path", PATHS)
def test_mttkrp_zero_mean_data(path):
    """Zero-mean data does not average rounding bias away: a plain-TF32
    tensor-core product fails this test; the 3xTF32 split must pass it."""
    tb.set_kernel_path(path)
    rng = np.random.RandomState(11)
    shape, rank = (128, 96, 160), 32
    x = rng.standard_normal(shape).astype(np.float32)
    fs = [rng.standard_normal((s, rank)).astype(np.float32) for s in shape]
    xd, fd = dev(x), [dev(f) for f in fs]
    for mode in range(3):
        ref32 = O.unfolding_dot_khatri_rao(x, (None, fs), mode)
        truth = O.mttkrp_float64_truth(x, (None, fs), mode)
        out = host(tb.unfolding_dot_khatri_rao(xd, (None, fd), mode))
        assert rel_fro(out, ref32) <= 1e-5
        assert rel_fro(out, truth) <= 1e-5


def test_mttkrp_argument_conventions():
    rng = np.random.RandomState(5)
    shape, rank = (20, 30, 40), 6
    x = rng.random_sample(shape).astype(np.float32)
    fs = [rng.random_sample((s, rank)).astype(np.float32) for s in shape]
    xd = dev(x)
    # column-major factors (as LAPACK's U arrives), weights None, tuple or list container
    fcm = [dev(np.ascontiguousarray(f.T)).T for f in fs]
    ref = O.unfolding_dot_khatri_rao(x, (None, fs), 1)
    assert rel_fro(host(tb.unfolding_dot_khatri_rao(xd, [None, fcm], 1)), ref) <= 1e-5
    # the skipped factor is never read: garbage / wrong shape there is fine, as in the reference
    fbad = list(fcm)
    fbad[1] = None
    assert rel_fro(host(tb.unfolding_dot_khatri_rao(xd, (None, fbad), 1)), ref) <= 1e-5
    # inputs are not mutated
    before = [host(f).copy() for f in fcm]
    tb.unfolding_dot_khatri_rao(xd, (None, fcm), 0)
    for a, b in zip(before, fcm):
        assert np.array_equal(a, host(b))
    # non-contiguous tensor
    xp = dev(np.ascontiguousarray(np.transpose(x, (2, 0, 1)))).permute(1, 2, 0)
    assert rel_fro(host(tb.unfolding_dot_khatri_rao(xp, (None, fcm), 2)),
                   O.unfolding_dot_khatri_rao(x, (None, fs), 2)) <= 1e-5
    # 2-way tensors ignore the weights, exactly like the reference (_khatri_rao.py:68-69)
    x2 = rng.random_sample((30, 20)).astype(np.float64)
    f2 = [rng.random_sample((30, 4)), rng.random_sample((20, 4))]
    w2 = rng.random_sample(4) + 1
    ref2 = O.unfolding_dot_khatri_rao(x2, (w2, f2), 0)
    assert rel_fro(host(tb.unfolding_dot_khatri_rao(dev(x2), (dev(w2), [dev(f) for f in f2]), 0)), ref2) <= 1e-12
    with pytest.raises(ValueError):
        tb.unfolding_dot_khatri_rao(xd, (None, fcm[:2]), 0)
    with pytest.raises(ValueError):
        bad = list(fcm)
        bad[2] = dev(rng.random_sample((41, rank)).astype(np.float32))
        tb.unfolding_dot_khatri_rao(xd, (None, bad), 0)
    with pytest.raises(TypeError):
        tb.unfolding_dot_khatri_rao(xd, (None, [f.double() for f in fcm]), 0)


def test_mttkrp_edge_cases():
    """Ragged / degenerate inputs: storage that TMA cannot describe (falls back to the SIMT kernel under AUTO and
    refuses under a forced tcgen05 path), rank 1, unit extents, extents far from any tile size."""
    rng = np.random.RandomState(11)
    # a contiguous tensor whose storage starts 4 bytes off a 16-byte boundary
    shape, rank = (128, 64, 96), 32
    flat = torch.empty(int(np.prod(shape)) + 1, dtype=torch.float32, device="cuda")
    x = rng.random_sample(shape).astype(np.float32)
    xd = flat[1:].view(shape)
    xd.copy_(dev(x))
    assert xd.data_ptr() % 16 == 4 and xd.is_contiguous()
    fs = [rng.random_sample((s, rank)).astype(np.float32) for s in shape]
    fd = [dev(f) for f in fs]
    for mode in range(3):
        out = host(tb.unfolding_dot_khatri_rao(xd, (None, fd), mode))
        assert tb.last_kernel_path() == "simt"
        assert rel_fro(out, O.unfolding_dot_khatri_rao(x, (None, fs), mode)) <= 1e-5
    tb.set_kernel_path("tcgen05")
    with pytest.raises(RuntimeError):
        tb.unfolding_dot_khatri_rao(xd, (None, fd), 0)
    tb.set_kernel_path("auto")
    # the same tensor, aligned: tensor cores
    tb.unfolding_dot_khatri_rao(dev(x), (None, fd), 0)
    assert tb.last_kernel_path() == "tcgen05"
    # rank 1, unit extents, primes
    for shp, r in (((64, 48, 80), 1), ((1, 37, 41), 3), ((37, 1, 41), 3), ((37, 41, 1), 3), ((7, 11, 13, 5), 2), ((2, 3), 1)):
        for dtype in (np.float32, np.float64):
            xx = rng.standard_normal(shp).astype(dtype)
            ff = [rng.standard_normal((s_, r)).astype(dtype) for s_ in shp]
            ww = (rng.random_sample(r) + 0.5).astype(dtype)
            for mode in range(len(shp)):
                got = host(tb.unfolding_dot_khatri_rao(dev(xx), (dev(ww), [dev(f) for f in ff]), mode))
                ref = O.unfolding_dot_khatri_rao(xx, (ww, ff), mode)
                assert got.shape == ref.shape
                assert rel_fro(got, ref) <= TOL[np.dtype(dtype)] * 4, (shp, r, mode)


# --------------------------------------------------------------------------- full-size properties (BASELINE configs)
def _lowrank_tensor(shape, comps, gen):
    """X = sum_q a_q o b_q o c_q (+ ...) built on the device, returned with its fp64 factor vectors."""
    vs = [torch.rand(s, comps, generator=gen, device="cuda", dtype=torch.float64) + 0.1 for s in shape]
    letters = "ijkl"[:len(shape)]
    x = torch.einsum(",".join(f"{c}q" for c in letters) + "->" + letters, *[v.float() for v in vs])
    return x, vs


def test_full_size_c2_properties():
    """C2 (1024^3 fp32, rank 32) is far beyond what the CPU oracle can check directly, so use identities:
    (i) a tensor with known CP structure has a closed-form MTTKRP (fp64), (ii) MTTKRP is additive over mode-0 slabs
    (the multi-GPU partition), (iii) the dimension-tree MTTKRP equals the direct one, (iv) unfold/fold round-trip
    bit-exactly and match torch's own permuting copy, (v) khatri_rao equals the broadcast product bit for bit."""
    n, R = 1024, 32
    gen = torch.Generator(device="cuda").manual_seed(7)
    x, vs = _lowrank_tensor((n, n, n), 3, gen)
    x += 0.0                                                   # contiguous fp32 tensor, 4.3 GB
    fs = [torch.rand(n, R, generator=gen, device="cuda") for _ in range(3)]
    w = torch.rand(R, generator=gen, device="cuda") + 0.5
    f64 = [f.double() for f in fs]
    vs32 = [v.float().double() for v in vs]                   # the values that actually went into x
    for mode in range(3):
        others = [m for m in range(3) if m != mode]
        # M[i, r] = w_r * sum_q v_mode[i, q] * prod_{m != mode} (v_m[:, q] . F_m[:, r])
        inner = torch.ones(3, R, dtype=torch.float64, device="cuda")
        for m in others:
            inner = inner * (vs32[m].T @ f64[m])
        truth = (vs32[mode] @ inner) * w.double()[None, :]
        got = tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
        assert tb.last_kernel_path() == "tcgen05"
        # x itself carries one fp32 rounding per element (~6e-8 relative, averaged down by the sums)
        err = float(torch.linalg.norm(got.double() - truth) / torch.linalg.norm(truth))
        assert err <= 1e-5, (mode, err)
        if mode > 0:                                           # additivity over mode-0 slabs
            parts = sum(tb.unfolding_dot_khatri_rao(x[lo:lo + 256], (w, [fs[0][lo:lo + 256]] + fs[1:]), mode)
                        for lo in range(0, n, 256))
            assert float(torch.linalg.norm(parts.double() - truth) / torch.linalg.norm(truth)) <= 1e-5
    t = tb.mode_dot(x, fs[2], 2, transpose=True)
    for mode in range(2):
        a = tb.mttkrp_from_ttm(t, (w, fs), mode)
        b = tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
        assert float(torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double())) <= 5e-6
    del t
    for mode in range(3):
        u = tb.unfold(x, mode, contiguous=True)
        assert torch.equal(u, x.movedim(mode, 0).reshape(n, -1))
        assert torch.equal(tb.fold(u, mode, x.shape), x)
        del u
    kr = tb.khatri_rao(fs[:2], weights=w)
    assert torch.equal(kr, ((fs[0] * w)[:, None, :] * fs[1][None, :, :]).reshape(-1, R))


def test_full_size_c4_properties():
    """C4 (256^4 fp32, rank 64, 17 GB): closed-form MTTKRP of a structured tensor for all four modes (the 4-way
    P/Q table split, rank-64 engine) and the dimension-tree path."""
    n, R = 256, 64
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~40 GB of free device memory")
    gen = torch.Generator(device="cuda").manual_seed(9)
    x, vs = _lowrank_tensor((n, n, n, n), 2, gen)
    fs = [torch.rand(n, R, generator=gen, device="cuda") for _ in range(4)]
    w = torch.rand(R, generator=gen, device="cuda") + 0.5
    f64 = [f.double() for f in fs]
    vs32 = [v.float().double() for v in vs]
    t = tb.mode_dot(x, fs[3], 3, transpose=True)
    for mode in range(4):
        inner = torch.ones(2, R, dtype=torch.float64, device="cuda")
        for m in range(4):
            if m != mode:
                inner = inner * (vs32[m].T @ f64[m])
        truth = (vs32[mode] @ inner) * w.double()[None, :]
        got = tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
        assert tb.last_kernel_path() == "tcgen05"
        assert float(torch.linalg.norm(got.double() - truth) / torch.linalg.norm(truth)) <= 1e-5, mode
        if mode < 3:
            got = tb.mttkrp_from_ttm(t, (w, fs), mode)
            assert float(torch.linalg.norm(got.double() - truth) / torch.linalg.norm(truth)) <= 1e-5, mode


def test_full_size_c3_ttm_properties():
    """C3 (512^3 fp32, ranks 64): the HOOI projection of a tensor with known CP structure has a closed form."""
    n, J = 512, 64
    gen = torch.Generator(device="cuda").manual_seed(8)
    x, vs = _lowrank_tensor((n, n, n), 4, gen)
    us = [torch.randn(n, J, generator=gen, device="cuda").t().contiguous().t() for _ in range(3)]   # column-major, as HOOI passes them
    vs32 = [v.float().double() for v in vs]
    proj = [u.double().T @ v for u, v in zip(us, vs32)]                     # (J, comps)
    truth = torch.einsum("aq,bq,cq->abc", *proj)
    got = tb.multi_mode_dot(x, us, transpose=True)
    assert tb.last_kernel_path() == "tcgen05"
    assert float(torch.linalg.norm(got.double() - truth) / torch.linalg.norm(truth)) <= 1e-5
    for skip in range(3):
        got = tb.multi_mode_dot(x, us, skip=skip, transpose=True)
        ops = [vs32[m] if m == skip else proj[m] for m in range(3)]
        truth = torch.einsum("aq,bq,cq->abc", *ops)
        assert float(torch.linalg.norm(got.double() - truth) / torch.linalg.norm(truth)) <= 1e-5


# --------------------------------------------------------------------------- mode_dot / multi_mode_dot
@pytest.mark.parametrize("path", PATHS)
def test_mode_dot_golden(golden, path):
    tb.set_kernel_path(path)
    g = golden("mode_dot")
    assert np.array_equal(host(tb.mode_dot(dev(g["a/x"]), dev(g["a/m"]), 0)), g["a/out"])
    assert np.array_equal(host(tb.mode_dot(dev(g["a/x"]), dev(g["a/v"]), 2)), g["a/outv"])
    for case in ("b", "c", "d"):
        x = g[f"{case}/x"]
        xd = dev(x)
        for mode in range(x.ndim):
            m, v = g[f"{case}/m{mode}"], g[f"{case}/v{mode}"]
            out = tb.mode_dot(xd, dev(m), mode)
            assert tuple(out.shape) == g[f"{case}/out{mode}"].shape
            assert rel_fro(host(out), g[f"{case}/out{mode}"]) <= TOL[x.dtype]
            outT = tb.mode_dot(xd, dev(np.ascontiguousarray(m.T)), mode, transpose=True)
            assert rel_fro(host(outT), g[f"{case}/outT{mode}"]) <= TOL[x.dtype]
            outv = tb.mode_dot(xd, dev(v), mode)
            assert tuple(outv.shape) == g[f"{case}/outv{mode}"].shape
            assert rel_fro(host(outv), g[f"{case}/outv{mode}"]) <= TOL[x.dtype]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,J", [((64, 96, 128), 32), ((128, 128, 128), 64), ((50, 33, 70), 17), ((20, 12, 10, 16), 8),
                                     ((64, 96, 128), 100), ((128, 128, 128), 150),
                                     ((256, 40), 24), ((8, 512, 64), 64), ((64, 64, 3), 5)])
def test_mode_dot_vs_oracle(shape, J, dtype, path):
    tb.set_kernel_path(path)
    rng = np.random.RandomState(7)
    x = rng.standard_normal(shape).astype(dtype)
    xd = dev(x)
    for mode, s in enumerate(shape):
        m = rng.standard_normal((J, s)).astype(dtype)
        ref = O.mode_dot(x, m, mode)
        out = tb.mode_dot(xd, dev(m), mode)
        assert out.is_contiguous() and tuple(out.shape) == ref.shape
        assert rel_fro(host(out), ref) <= TOL[np.dtype(dtype)], (shape, mode, tb.last_kernel_path())
        # column-major U with transpose=True: how HOOI calls it (_tucker.py:194-196)
        u = dev(np.ascontiguousarray(m)).T      # (s, J) view with stride (1, s)
        assert rel_fro(host(tb.mode_dot(xd, u, mode, transpose=True)), ref) <= TOL[np.dtype(dtype)]


def test_mode_dot_many_short_items_rowwise():
    """Regression: a persistent CTA that runs several short work items back to back (last-mode
    TTM, 256 row tiles x 8 K chunks) once handed a shared-memory stage back to TMA before every
    warp had consumed it; a few rows per tile came out ~10% off, invisible in a Frobenius norm.
    Check every output row on its own."""
    rng = np.random.RandomState(17)
    x = rng.standard_normal((256, 128, 512)).astype(np.float32)
    m = rng.standard_normal((64, 512)).astype(np.float32)
    ref = O.mode_dot(x.astype(np.float64), m.astype(np.float64), 2).reshape(-1, 64)
    out = host(tb.mode_dot(dev(x), dev(m), 2)).reshape(-1, 64).astype(np.float64)
    row_err = np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert row_err.max() <= 1e-4, (int(np.argmax(row_err)), float(row_err.max()), tb.last_kernel_path())
    # same for MTTKRP with many short items: (32768, 8, 64), mode 0
    x = rng.standard_normal((32768, 8, 64)).astype(np.float32)
    fs = [rng.standard_normal((s, 64)).astype(np.float32) for s in x.shape]
    ref = O.mttkrp_float64_truth(x, (None, fs), 0)
    out = host(tb.unfolding_dot_khatri_rao(dev(x), (None, [dev(f) for f in fs]), 0)).astype(np.float64)
    row_err = np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert row_err.max() <= 1e-4, (int(np.argmax(row_err)), float(row_err.max()), tb.last_kernel_path())


def test_mode_dot_errors():
    x = dev(np.zeros((3, 4, 2), dtype=np.float32))
    with pytest.raises(ValueError):
        tb.mode_dot(x, dev(np.zeros((2, 5), dtype=np.float32)), 0)      # test_n_mode_product.py:65-72
    with pytest.raises(ValueError):
        tb.mode_dot(x, dev(np.zeros((5, 2), dtype=np.float32)), 0, transpose=True)
    with pytest.raises(ValueError):
        tb.mode_dot(x, dev(np.zeros(5, dtype=np.float32)), 1)
    with pytest.raises(ValueError):
        tb.mode_dot(x, dev(np.zeros((2, 2, 2), dtype=np.float32)), 1)


@pytest.mark.parametrize("path", PATHS)
def test_multi_mode_dot_golden(golden, path):
    tb.set_kernel_path(path)
    g = golden("multi_mode_dot")
    for case in g.cases():
        x = g[f"{case}/x"]
        fs = g.arrays(case, "f")
        tol = TOL[x.dtype]
        xd = dev(x)
        # column-major (I_n, R_n) factors with transpose=True, as HOOI passes them
        fd = [dev(np.ascontiguousarray(f.T)).T for f in fs]
        assert rel_fro(host(tb.multi_mode_dot(xd, fd, transpose=True)), g[f"{case}/full"]) <= tol
        for k in range(x.ndim):
            out = tb.multi_mode_dot(xd, fd, skip=k, transpose=True)
            assert tuple(out.shape) == g[f"{case}/skip{k}"].shape
            assert rel_fro(host(out), g[f"{case}/skip{k}"]) <= tol
        ms = [dev(np.ascontiguousarray(fs[2].T)), dev(np.ascontiguousarray(fs[0].T))]
        assert rel_fro(host(tb.multi_mode_dot(xd, ms, modes=[2, 0])), g[f"{case}/sub20"]) <= tol
        out = tb.multi_mode_dot(xd, [dev(g[f"{case}/vec0"]), dev(g[f"{case}/vec2"])], modes=[0, 2])
        assert tuple(out.shape) == g[f"{case}/vecs02"].shape
        assert rel_fro(host(out), g[f"{case}/vecs02"]) <= tol


def test_multi_mode_dot_vector_order_independence():
    # test_n_mode_product.py:139-153: contracting every mode with a vector gives the same
    # scalar whatever the order the pairs are listed in
    rng = np.random.RandomState(9)
    shape = (5, 6, 7, 4)
    x = rng.random_sample(shape)
    vs = [rng.random_sample(s) for s in shape]
    truth = np.einsum("ijkl,i,j,k,l->", x, *vs)
    xd, vd = dev(x), [dev(v) for v in vs]
    for perm in itertools.permutations(range(4)):
        out = tb.multi_mode_dot(xd, [vd[i] for i in perm], modes=list(perm))
        assert out.dim() == 0
        assert abs(float(out) - truth) <= 1e-10 * abs(truth)


# --------------------------------------------------------------------------- normal-equation kernels
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("rows,rank", [(1024, 32), (2048, 64), (100, 10), (37, 5), (513, 100)])
def test_gram_update_error_kernels(rows, rank, dtype):
    rng = np.random.RandomState(13)
    tol = 2e-5 if dtype == np.float32 else 1e-11
    fs = [rng.random_sample((rows, rank)).astype(dtype), rng.random_sample((50, rank)).astype(dtype),
          rng.random_sample((60, rank)).astype(dtype)]
    w = (rng.random_sample(rank) + 0.5).astype(dtype)
    grams = [f.T @ f for f in fs]
    gd = [tb.gram(dev(f)) for f in fs]
    for a, b in zip(gd, grams):
        assert rel_fro(host(a), b) <= tol
    # column-major factor
    assert rel_fro(host(tb.gram(dev(np.ascontiguousarray(fs[0].T)).T)), grams[0]) <= tol
    m = rng.random_sample((rows, rank)).astype(dtype)
    l2 = 0.1
    V = np.ones((rank, rank), dtype=dtype) * grams[1] * grams[2] + np.eye(rank, dtype=dtype) * dtype(l2)
    V = w[:, None] * V * w[None, :]
    ref = np.linalg.solve(V.T.astype(np.float64), m.T.astype(np.float64)).T
    out = tb.cp_update(gd, 0, dev(w), dev(m), l2_reg=l2)
    cond = np.linalg.cond(V.astype(np.float64))
    eps = np.finfo(dtype).eps
    assert rel_fro(host(out), ref) <= 20 * cond * eps
    # fused variant: same rows, plus the Gram of the updated factor from the same launch (twice: the ticket
    # counter in the workspace must come back to zero)
    for _ in range(2):
        g_new = torch.empty((rank, rank), dtype=gd[0].dtype, device="cuda")
        out2 = tb.cp_update(gd, 0, dev(w), dev(m), l2_reg=l2, gram_out=g_new)
        assert torch.equal(out2, out)
        o64 = host(out).astype(np.float64)
        assert rel_fro(host(g_new), o64.T @ o64) <= tol
    # MU update
    f0 = dev(fs[0].copy())
    acc = w[:, None] * (grams[1] * grams[2]) * w[None, :]
    e = np.finfo(dtype).eps
    ref_mu = fs[0] * np.clip(m, e, None) / np.clip(fs[0] @ acc, e, None)
    tb.nncp_update(gd, 0, dev(w), dev(m), f0, float(e))
    assert rel_fro(host(f0), ref_mu) <= (1e-5 if dtype == np.float32 else 1e-12)
    # error
    nx2 = dev(np.array([12345.678], dtype=dtype))
    errs = host(tb.cp_error(gd, dev(w), dev(m), dev(fs[0]), nx2))
    iprod = np.sum(m.astype(np.float64) * fs[0])
    ncp = np.sum((grams[0] * grams[1] * grams[2]).astype(np.float64) * np.outer(w, w))
    assert abs(errs[1] - iprod) <= tol * abs(iprod) * 10
    assert abs(errs[2] - ncp) <= tol * abs(ncp) * 10
    ref_err = np.sqrt(abs(12345.678 + ncp - 2 * iprod)) / np.sqrt(12345.678)
    assert abs(errs[0] - ref_err) <= 1e-4 * ref_err
    x = rng.standard_normal(100003).astype(dtype)
    assert abs(float(tb.sumsq(dev(x))) - float(np.sum(x.astype(np.float64) ** 2))) <= 1e-6 * float(np.sum(x.astype(np.float64) ** 2))


# --------------------------------------------------------------------------- ALS level
@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("tag", ["p32", "p64", "p4way"])
def test_parafac_driver_vs_reference(golden, tag, use_graph):
    g = golden("als")
    x = g[f"{tag}/x"]
    rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
    init = (None, [dev(f) for f in g.arrays(tag, "init")])
    cp, errs = tb.parafac(dev(x), rank, n_iter_max=iters, init=init, tol=0, return_errors=True, use_graph=use_graph)
    ref = g[f"{tag}/errors"]
    assert len(errs) == iters
    rel = np.max(np.abs(np.array(errs) - ref) / ref)
    assert rel <= 1e-4, rel                    # north-star ALS gate
    if x.dtype == np.float64:
        assert rel <= 1e-9
        for a, b in zip(cp[1], g.arrays(tag, "f")):
            assert rel_fro(host(a), b) <= 1e-6


@pytest.mark.parametrize("shape,rank", [((96, 80, 112), 16), ((40, 24, 20, 12), 8), ((48, 1, 40), 8), ((24, 20, 1, 36), 6)])
def test_parafac_dimtree_same_trajectory(shape, rank):
    """Two tensor passes per sweep (dimension tree) vs N passes: same ALS trajectory."""
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(shape, generator=g, device="cuda")
    fs = [torch.rand(s, rank, generator=g, device="cuda") for s in shape]
    w = torch.ones(rank, device="cuda")
    runs = []
    for dimtree in (False, True):
        st = tb.CPALS(x, w, fs, dimtree=dimtree)
        assert st.dimtree == dimtree
        errs = []
        for _ in range(6):
            st.sweep(True)
            errs.append(float(st.err[0]))
        runs.append((errs, [f.clone() for f in st.factors]))
    dev_err = max(abs(a - b) / b for a, b in zip(*[r[0] for r in runs]))
    assert dev_err <= 1e-5, dev_err
    for a, b in zip(runs[0][1], runs[1][1]):
        assert float(torch.linalg.norm(a - b) / torch.linalg.norm(b)) <= 1e-3


def test_parafac_config1_fp64(golden):
    """BASELINE config 1: rank 10, 100^3, float64, identical input and initial factors."""
    g = golden("als")
    x = O.random_tensor((100, 100, 100), 0)
    w, fs = O.random_cp_factors((100, 100, 100), 10, 1)
    iters = int(g["c1/iters"])
    cp, errs = tb.parafac(dev(x), 10, n_iter_max=iters, init=(None, [dev(f) for f in fs]), tol=0, return_errors=True)
    ref = g["c1/errors"]
    assert np.max(np.abs(np.array(errs) - ref) / ref) <= 1e-9
    assert rel_fro(host(cp[1][0])[:4], g["c1/f0_head"]) <= 1e-6
    # init="random" reproduces random_cp(random_state=1): same trajectory
    cp2, errs2 = tb.parafac(dev(x), 10, n_iter_max=3, init="random", random_state=1, tol=0, return_errors=True)
    assert np.max(np.abs(np.array(errs2) - ref[:3]) / ref[:3]) <= 1e-9


def test_parafac_lowrank_and_tol(golden):
    g = golden("als")
    x = g["lowrank/x"]
    init = (None, [dev(f) for f in g.arrays("lowrank", "init")])
    cp, errs = tb.parafac(dev(x), 5, n_iter_max=8, init=init, tol=0, return_errors=True)
    assert np.max(np.abs(np.array(errs) - g["lowrank/errors"])) <= 1e-8
    # early stop on tolerance: stops where the reference's rule stops
    ref = g["lowrank/errors"]
    dec = np.abs(np.diff(ref))
    tol = float(np.sqrt(dec[3] * dec[4]))          # between the 4th and 5th decrease
    stop = next(i for i in range(1, len(ref)) if abs(ref[i - 1] - ref[i]) < tol)
    _, errs2 = tb.parafac(dev(x), 5, n_iter_max=8, init=init, tol=tol, return_errors=True)
    assert len(errs2) == stop + 1
    # init tuple is not mutated
    for a, b in zip(init[1], g.arrays("lowrank", "init")):
        assert np.array_equal(host(a), b)


def test_non_negative_parafac_driver_vs_reference(golden):
    g = golden("als")
    x = g["nn/x"]
    init = (None, [dev(f) for f in g.arrays("nn", "init")])
    cp, errs = tb.non_negative_parafac(dev(x), 6, n_iter_max=10, init=init, tol=1e-30, return_errors=True)
    ref = g["nn/errors"]
    assert np.max(np.abs(np.array(errs) - ref) / ref) <= 1e-4
    for a, b in zip(cp[1], g.arrays("nn", "f")):
        assert rel_fro(host(a), b) <= 1e-3
        assert float(a.min()) >= 0.0


# --------------------------------------------------------------------------- unmodified TensorLy on the b200 backend
@pytest.fixture
def tl_b200():
    try:
        tl = tb.import_tensorly()
    except ImportError:
        pytest.skip("tensorly not importable on this box")
    prev_backend, prev_tenalg = tl.get_backend(), tl.tenalg.get_backend()
    tl.set_backend("pytorch")
    tb.use()
    yield tl
    tl.tenalg.set_backend(prev_tenalg)
    tl.set_backend(prev_backend)


def test_reference_parafac_runs_unmodified_on_backend(tl_b200, golden):
    tl = tl_b200
    from tensorly.cp_tensor import CPTensor
    from tensorly.decomposition import parafac
    g = golden("als")
    for tag in ("p32", "p64", "p4way"):
        x = g[f"{tag}/x"]
        rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
        init = CPTensor((torch.ones(rank, dtype=dev(x).dtype, device="cuda"), [dev(f) for f in g.arrays(tag, "init")]))
        cp, errs = parafac(dev(x), rank, n_iter_max=iters, init=init, tol=0, return_errors=True)
        ref = g[f"{tag}/errors"]
        got = np.array([float(e) for e in errs])
        assert np.max(np.abs(got - ref) / ref) <= 1e-4
        assert tb.last_kernel_path() in ("simt", "dmma", "tcgen05", "tcgen05-f16")


def test_reference_parafac_on_backend_with_dimension_tree_cache(tl_b200, golden):
    """The opt-in reuse behind the stateless API: same trajectories as the reference, T formed once per sweep,
    and never a stale hit (new tensor object / in-place edit of the last factor are misses)."""
    from tensorly.cp_tensor import CPTensor
    from tensorly.decomposition import parafac
    from tensorly.tenalg import unfolding_dot_khatri_rao as tl_mttkrp
    g = golden("als")
    tb.set_dimension_tree(True)
    try:
        for tag in ("p32", "p64", "p4way"):
            x = g[f"{tag}/x"]
            rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
            init = CPTensor((torch.ones(rank, dtype=dev(x).dtype, device="cuda"), [dev(f) for f in g.arrays(tag, "init")]))
            cp, errs = parafac(dev(x), rank, n_iter_max=iters, init=init, tol=0, return_errors=True)
            ref = g[f"{tag}/errors"]
            got = np.array([float(e) for e in errs])
            assert np.max(np.abs(got - ref) / ref) <= 1e-4
        # staleness: edit the last factor in place between two calls for the same tensor
        rng = np.random.RandomState(5)
        x = dev(rng.random_sample((40, 30, 20)).astype(np.float32))
        fs = [dev(rng.random_sample((s, 6)).astype(np.float32)) for s in (40, 30, 20)]
        m0 = tl_mttkrp(x, (None, fs), 0)
        assert rel_fro(host(tl_mttkrp(x, (None, fs), 1)), O.unfolding_dot_khatri_rao(host(x), (None, [host(f) for f in fs]), 1)) <= 1e-5
        fs[2].mul_(2.0)
        got = host(tl_mttkrp(x, (None, fs), 1))
        assert rel_fro(got, O.unfolding_dot_khatri_rao(host(x), (None, [host(f) for f in fs]), 1)) <= 1e-5
        assert rel_fro(host(m0), O.unfolding_dot_khatri_rao(host(x), (None, [host(fs[0]), host(fs[1]), host(fs[2]) / 2]), 0)) <= 1e-5
    finally:
        tb.set_dimension_tree(False)


def test_reference_nn_parafac_and_tucker_run_unmodified_on_backend(tl_b200, golden):
    from tensorly.cp_tensor import CPTensor
    from tensorly.decomposition import non_negative_parafac, partial_tucker, tucker
    g = golden("als")
    x = g["nn/x"]
    init = CPTensor((torch.ones(6, dtype=torch.float32, device="cuda"), [dev(f) for f in g.arrays("nn", "init")]))
    cp, errs = non_negative_parafac(dev(x), 6, n_iter_max=10, init=init, tol=1e-30, return_errors=True)
    ref = g["nn/errors"]
    assert np.max(np.abs(np.array([float(e) for e in errs]) - ref) / ref) <= 1e-4
    x = g["tucker/x"]
    ranks = [int(r) for r in g["tucker/ranks"]]
    (core, factors), errs = tucker(dev(x), ranks, n_iter_max=5, init="random", random_state=1, tol=0, return_errors=True)
    ref = g["tucker/errors"]
    assert np.max(np.abs(np.array([float(e) for e in errs]) - ref) / ref) <= 1e-4
    assert abs(float(torch.linalg.norm(core)) - float(g["tucker/core_norm"])) <= 1e-6 * float(g["tucker/core_norm"])
    (core2, factors2), errs2 = partial_tucker(dev(x), ranks[:2], modes=[0, 1], n_iter_max=3, init="svd", tol=0)
    assert tuple(core2.shape) == (ranks[0], ranks[1], x.shape[2])
    assert tb.last_kernel_path() in ("simt", "dmma", "tcgen05", "tcgen05-f16")


def test_reference_tucker_with_gram_svd_plugin(tl_b200, golden):
    """HOOI with the Gram + eigh SVD plug-in follows the reference trajectory."""
    from tensorly.decomposition import tucker
    g = golden("als")
    x = g["tucker/x"]
    ranks = [int(r) for r in g["tucker/ranks"]]
    tb.use_gram_svd()
    try:
        (core, factors), errs = tucker(dev(x), ranks, n_iter_max=5, init="random", random_state=1, tol=0, return_errors=True)
    finally:
        tb.use_default_svd()
    ref = g["tucker/errors"]
    assert np.max(np.abs(np.array([float(e) for e in errs]) - ref) / ref) <= 1e-4
    assert abs(float(torch.linalg.norm(core)) - float(g["tucker/core_norm"])) <= 1e-5 * float(g["tucker/core_norm"])


def test_reference_parafac_with_fast_solve_plugin(tl_b200, golden):
    """The sync-free `solve` hook (solve.py): unmodified parafac follows the reference trajectories, and the hook
    agrees with torch.linalg.solve on the systems the loop hands it (transposed views) and declines the rest."""
    from tensorly.cp_tensor import CPTensor
    from tensorly.decomposition import parafac
    g = golden("als")
    tb.use_fast_solve()
    try:
        for tag in ("p32", "p64", "p4way"):
            x = g[f"{tag}/x"]
            rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
            init = CPTensor((torch.ones(rank, dtype=dev(x).dtype, device="cuda"), [dev(f) for f in g.arrays(tag, "init")]))
            cp, errs = parafac(dev(x), rank, n_iter_max=iters, init=init, tol=0, return_errors=True)
            ref = g[f"{tag}/errors"]
            got = np.array([float(e) for e in errs])
            assert np.max(np.abs(got - ref) / ref) <= 1e-4
    finally:
        tb.use_default_solve()
    gen = torch.Generator(device="cuda").manual_seed(4)
    for n, cols, dt in ((32, 1000, torch.float32), (64, 300, torch.float32), (10, 77, torch.float64), (100, 50, torch.float64)):
        f = torch.rand(500, n, generator=gen, device="cuda", dtype=dt)
        v = f.T @ f + torch.eye(n, device="cuda", dtype=dt)
        m = torch.rand(cols, n, generator=gen, device="cuda", dtype=dt)
        got = tb.fast_solve(v.T, m.T)
        ref = torch.linalg.solve(v.T.double(), m.T.double())
        assert tuple(got.shape) == (n, cols)
        assert float(torch.linalg.norm(got.double() - ref) / torch.linalg.norm(ref)) <= (1e-3 if dt == torch.float32 else 1e-10)
    # not a small square CUDA system: handed to the previous solve untouched
    a = torch.rand(200, 200, device="cuda", dtype=torch.float64) + 200 * torch.eye(200, device="cuda", dtype=torch.float64)
    b = torch.rand(200, 3, device="cuda", dtype=torch.float64)
    assert torch.allclose(tb.fast_solve(a, b), torch.linalg.solve(a, b))


def test_reference_reconstruction_uses_backend(tl_b200):
    tl = tl_b200
    rng = np.random.RandomState(2)
    fs = [dev(rng.random_sample((s, 4))) for s in (6, 7, 8)]
    w = dev(rng.random_sample(4))
    full = tl.cp_to_tensor((w, fs))              # calls the dispatched khatri_rao
    ref = O.cp_to_tensor((host(w), [host(f) for f in fs]))
    assert rel_fro(host(full), ref) <= 1e-12
    assert tl.tenalg.get_backend() == "b200"
    with pytest.raises(ValueError):
        tl.tenalg.set_backend("no-such-backend")


# --------------------------------------------------------------------------- reconstruction / masked ALS
def test_cp_to_tensor_golden(golden):
    """tlb200_cp_to_tensor against the reference's cp_to_tensor outputs (tests/golden/round2.npz)."""
    g = golden("round2")
    for tag in ("rec_a", "rec_b", "rec_c", "rec_d", "rec_e", "rec_f"):
        fs = g.arrays(tag, "f")
        w = g[f"{tag}/w"] if g.has(f"{tag}/w") else None
        ref = g[f"{tag}/out"]
        out = tb.cp_to_tensor((dev(w) if w is not None else None, [dev(f) for f in fs]))
        assert tuple(out.shape) == ref.shape
        assert rel_fro(host(out), ref) <= TOL[ref.dtype], tag
        if len(fs) >= 2:
            # column-major factors (strided views) and an element-wise mask
            fT = [dev(np.ascontiguousarray(f.T)).T for f in fs]
            assert rel_fro(host(tb.cp_to_tensor((dev(w) if w is not None else None, fT))), ref) <= TOL[ref.dtype]
            mask = (np.random.RandomState(1).random_sample(ref.shape) > 0.4).astype(ref.dtype)
            got = tb.cp_to_tensor((dev(w) if w is not None else None, [dev(f) for f in fs]), mask=dev(mask))
            assert rel_fro(host(got), ref * mask) <= TOL[ref.dtype]
    with pytest.raises(ValueError):
        tb.cp_to_tensor((None, [dev(np.ones((4, 3))), dev(np.ones((5, 2)))]))
    with pytest.raises(ValueError):
        tb.cp_to_tensor((dev(np.ones(4)), [dev(np.ones((4, 3))), dev(np.ones((5, 3)))]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,rank", [((96, 80, 112), 16), ((130, 70, 45), 10), ((40, 24, 20, 12), 33), ((300, 200), 64),
                                        ((257, 129), 100),
                                        # large enough for the tcgen05 variant (fp32, rank <= 64): aligned, ragged rows,
                                        # odd column count (scalar stores), 4-way, 2-way
                                        ((256, 96, 80), 32), ((200, 130, 52), 20), ((300, 77, 61), 64),
                                        ((64, 32, 24, 40), 48), ((2000, 1100), 7),
                                        # two contraction chunks through the TMA epilogue (one / two lines per box)
                                        ((256, 64, 128), 64), ((192, 80, 96), 48), ((100, 90, 124), 40),
                                        # 4-way instances of the tensor-core kernel (two-level column odometer)
                                        ((128, 32, 36, 28), 24), ((96, 20, 24, 32), 64)])
def test_cp_to_tensor_and_impute_vs_oracle(shape, rank, dtype):
    rng = np.random.RandomState(17)
    fs = [(rng.random_sample((s, rank)) - 0.4).astype(dtype) for s in shape]
    w = (rng.random_sample(rank) + 0.5).astype(dtype)
    ref = O.cp_to_tensor((w, fs))
    out = tb.cp_to_tensor((dev(w), [dev(f) for f in fs]))
    assert rel_fro(host(out), ref) <= TOL[np.dtype(dtype)]
    big = shape[0] >= 64 and np.prod(shape[1:]) >= 1024 and np.prod(shape) >= 2 ** 20
    assert tb.last_kernel_path() == ("tcgen05" if (big and dtype == np.float32 and rank <= 64) else "simt")
    if big:
        # every row of the tensor on its own (a wrong tile or swizzle shows up as a few bad rows, not in the norm)
        got = host(out).reshape(shape[0], -1).astype(np.float64)
        want = ref.reshape(shape[0], -1).astype(np.float64)
        rows = np.linalg.norm(got - want, axis=1) / np.maximum(np.linalg.norm(want, axis=1), 1e-30)
        assert rows.max() <= 10 * TOL[np.dtype(dtype)], int(rows.argmax())
        mask2 = (rng.random_sample(shape) > 0.5).astype(dtype)
        got2 = tb.cp_to_tensor((dev(w), [dev(f) for f in fs]), mask=dev(mask2))
        assert rel_fro(host(got2), ref * mask2) <= TOL[np.dtype(dtype)]
    x = rng.standard_normal(shape).astype(dtype)
    mask = (rng.random_sample(shape) > 0.3).astype(dtype)
    new_ref, nrm, unnorm = O.cp_impute(x, mask, (w, fs))
    xd = dev(x)
    new, stats = tb.cp_impute(xd, dev(mask), (dev(w), [dev(f) for f in fs]))
    assert np.array_equal(host(xd), x)                            # out-of-place by default
    assert rel_fro(host(new), new_ref) <= TOL[np.dtype(dtype)]
    st = host(stats).astype(np.float64)
    tol = 1e-5 if dtype == np.float32 else 1e-11
    assert abs(st[1] - float(nrm) ** 2) <= tol * float(nrm) ** 2
    assert abs(st[2] - float(unnorm) ** 2) <= tol * float(unnorm) ** 2
    assert abs(st[0] - float(unnorm) / float(nrm)) <= tol * float(unnorm) / float(nrm)
    # observed entries are kept bit for bit, in place too
    tb.cp_impute(xd, dev(mask), (dev(w), [dev(f) for f in fs]), out=xd)
    assert np.array_equal(host(xd)[mask == 1], x[mask == 1])


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("tag", ["mask64", "mask32", "mask4way"])
def test_masked_parafac_driver_vs_reference(golden, tag, use_graph):
    """parafac(mask=...) of the own driver (fused imputation + error pass) against the reference's trajectories."""
    g = golden("round2")
    x, mask = g[f"{tag}/x"], g[f"{tag}/mask"]
    rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
    xd = dev(x)
    init = (None, [dev(f) for f in g.arrays(tag, "init")])
    cp, errs = tb.parafac(xd, rank, n_iter_max=iters, init=init, tol=0, return_errors=True, mask=dev(mask), use_graph=use_graph)
    ref = g[f"{tag}/errors"]
    rel = np.max(np.abs(np.array(errs) - ref) / ref)
    assert rel <= (1e-9 if x.dtype == np.float64 else 1e-4), rel
    assert np.array_equal(host(xd), x)                            # the caller's tensor is not modified
    for a, b in zip(cp[1], g.arrays(tag, "f")):
        assert rel_fro(host(a), b) <= (1e-6 if x.dtype == np.float64 else 2e-2)


def test_non_negative_parafac_svd_init_vs_reference(golden):
    """init='svd' of non_negative_parafac is NNDSVDA (svd_interface(non_negative=True)), not |U|."""
    from tensorly_b200.cp_als import _svd_init
    g = golden("round2")
    x = g["nnsvd/x"]
    rank = int(g["nnsvd/rank"])
    _, fs = _svd_init(dev(x), rank, None, non_negative=True)
    for a, b in zip(fs, g.arrays("nnsvd", "init")):
        assert rel_fro(np.abs(host(a)), b) <= 1e-8
    _, errs = tb.non_negative_parafac(dev(x), rank, n_iter_max=8, init="svd", tol=1e-30, return_errors=True)
    ref = g["nnsvd/errors"]
    assert np.max(np.abs(np.array(errs) - ref) / ref) <= 1e-8


# --------------------------------------------------------------------------- own HOOI driver
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("rows,rank", [(512, 64), (300, 20), (64, 64), (1000, 7), (130, 33)])
def test_orthonormalize_and_symeig(rows, rank, dtype):
    rng = np.random.RandomState(5)
    # an ill-conditioned block: one dominant direction (like G U on a tensor with a large mean), cond ~ 1e5-1e6,
    # i.e. cond^2 far beyond 1 / eps(fp32): the Gram, its Cholesky factor and the inverse are formed in fp64
    z = rng.standard_normal((rows, rank)) + 1e4 * np.outer(rng.random_sample(rows), np.ones(rank))
    z = z.astype(dtype)
    q = host(tb.orthonormalize(dev(z))).astype(np.float64)              # two passes: orthonormal to output rounding
    tol = 5e-6 if dtype == np.float32 else 1e-12
    assert np.linalg.norm(q.T @ q - np.eye(rank)) <= tol * rank
    q1 = host(tb.orthonormalize(dev(z), passes=1)).astype(np.float64)   # one pass: defect ~ cond^2 * 1e-16
    cond = np.linalg.cond(z.astype(np.float64))
    assert np.linalg.norm(q1.T @ q1 - np.eye(rank)) <= max(tol * rank, 1e3 * cond ** 2 * 1e-16)    # incl. the 2^-43 shift
    tol = 5e-5 if dtype == np.float32 else 1e-9
    # same span: projecting z on q reproduces z
    z64 = z.astype(np.float64)
    assert np.linalg.norm(q @ (q.T @ z64) - z64) / np.linalg.norm(z64) <= tol
    qs = host(tb.orthonormalize(dev(np.ascontiguousarray(z.T)).T)).astype(np.float64)      # column-major input
    assert np.linalg.norm(qs - q) <= tol * rank
    a = rng.standard_normal((rank, rank))
    a = (a @ a.T + np.diag(rng.random_sample(rank))).astype(dtype)
    w, v = tb.symeig(dev(a))
    w, v = host(w).astype(np.float64), host(v).astype(np.float64)
    ref = np.linalg.eigvalsh(a.astype(np.float64))[::-1]
    etol = 1e-5 if dtype == np.float32 else 1e-12
    assert np.max(np.abs(w - ref)) <= etol * np.abs(ref).max()
    assert np.all(np.diff(w) <= 0)
    assert np.linalg.norm(a.astype(np.float64) @ v - v * w) <= etol * 10 * np.abs(ref).max() * rank
    assert np.linalg.norm(v.T @ v - np.eye(rank)) <= etol * rank


def test_own_tucker_vs_reference(golden):
    """tensorly_b200.tucker (TTM chains + Gram + warm-started subspace iteration, no SVD in the loop) against the
    reference's HOOI trajectories: random init (small modes: the exact Rayleigh-Ritz path), SVD init on a
    low-rank + noise tensor and on a uniform random tensor."""
    g, g2 = golden("als"), golden("round2")
    x = g["tucker/x"]
    ranks = [int(r) for r in g["tucker/ranks"]]
    tk, errs = tb.tucker(dev(x), ranks, n_iter_max=5, init="random", random_state=1, tol=0, return_errors=True)
    core, factors = tk
    ref = g["tucker/errors"]
    assert np.max(np.abs(np.array(errs) - ref) / ref) <= 1e-8
    assert abs(float(torch.linalg.norm(core)) - float(g["tucker/core_norm"])) <= 1e-8 * float(g["tucker/core_norm"])
    for f in factors:
        ff = host(f)
        assert np.linalg.norm(ff.T @ ff - np.eye(ff.shape[1])) <= 1e-8
    for tag, iters, tol in (("tucker_svd", 6, 1e-8), ("tucker_svd32", 8, 1e-4)):
        x = g2[f"{tag}/x"]
        ranks = [int(r) for r in g2[f"{tag}/ranks"]]
        tk, errs = tb.tucker(dev(x), ranks, n_iter_max=iters, init="svd", tol=0, return_errors=True)
        ref = g2[f"{tag}/errors"]
        assert np.max(np.abs(np.array(errs) - ref) / ref) <= tol, tag
        if tag == "tucker_svd":
            rec = tb.tucker_to_tensor(tk)
            assert rel_fro(host(rec), g2["tucker_svd/rec"]) <= 1e-8
    # partial tucker: modes 0 and 2 only
    (core, factors), errs = tb.partial_tucker(dev(g["tucker/x"]), [4, 6], modes=[0, 2], n_iter_max=4, init="svd", tol=0)
    assert tuple(core.shape) == (4, 22, 6) and len(factors) == 2 and len(errs) == 4
    assert errs[-1] <= errs[0] + 1e-12


@pytest.mark.parametrize("shape,ranks", [((160, 150, 140), [32, 32, 32]), ((200, 64, 300), [40, 20, 64]), ((100, 90, 80, 70), [10, 12, 8, 9])])
def test_own_tucker_large_modes_vs_oracle(shape, ranks):
    """Modes wider than 64 (oversampled subspace iteration + Rayleigh-Ritz, or plain subspace iteration at rank 64)
    on uniform random data — the flat-spectrum worst case — against the oracle's exact-SVD HOOI: 1e-4 on the errors."""
    rng = np.random.RandomState(0)
    x = rng.random_sample(shape).astype(np.float32)
    rs = np.random.RandomState(1)
    rs.random_sample(ranks)
    init = [rs.random_sample((s, r)) for s, r in zip(shape, ranks)]
    (_, _), ref = O.tucker_hooi(x, ranks, init, n_iter_max=5)
    tk, errs = tb.tucker(dev(x), ranks, n_iter_max=5, init="random", random_state=1, tol=0, return_errors=True)
    dev_ = np.max(np.abs(np.array(errs) - np.array(ref, dtype=np.float64)) / np.array(ref, dtype=np.float64))
    assert dev_ <= 1e-4, (errs, ref)


# --------------------------------------------------------------------------- HALS
def test_hals_nnls_golden(golden):
    """tlb200_hals_update against the reference's hals_nnls outputs (plain, and with sparsity / ridge / epsilon)."""
    g = golden("round2")
    for tag in ("hals_a", "hals_b", "hals_c"):
        UtM, UtU, V0 = g[f"{tag}/UtM"], g[f"{tag}/UtU"], g[f"{tag}/V0"]
        # fp32: the reference's own result is only reproducible to its rounding (summation order of the rank-length
        # dot products over up to 100 Gauss-Seidel passes): the oracle port matches it to 1e-4, the kernel to 5e-4
        tol = 1e-9 if UtM.dtype == np.float64 else 5e-4
        v0 = dev(V0)
        out = tb.hals_nnls(dev(UtM), dev(UtU), v0, n_iter_max=100)
        assert np.array_equal(host(v0), V0)                      # input untouched
        assert rel_fro(host(out), g[f"{tag}/V"]) <= tol, tag
        assert float(out.min()) >= 0.0
        out = tb.hals_nnls(dev(UtM), dev(UtU), dev(V0), n_iter_max=20, sparsity_coefficient=0.05, ridge_coefficient=0.1,
                           epsilon=1e-6)
        assert rel_fro(host(out), g[f"{tag}/V_sparse"]) <= tol, tag
        assert float(out.min()) >= 1e-6 * (1 - 1e-6)
    with pytest.raises(NotImplementedError):
        tb.hals_nnls(dev(np.ones((100, 5))), dev(np.eye(100)), dev(np.ones((100, 5))))      # rank 100 in fp64


@pytest.mark.parametrize("tag", ["nnhals64", "nnhals32"])
def test_non_negative_parafac_hals_driver_vs_reference(golden, tag):
    g = golden("round2")
    x = g[f"{tag}/x"]
    rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
    init = (None, [dev(f) for f in g.arrays(tag, "init")])
    cp, errs = tb.non_negative_parafac_hals(dev(x), rank, n_iter_max=iters, init=init, tol=1e-30, return_errors=True)
    ref = g[f"{tag}/errors"]
    assert len(errs) == iters
    assert np.max(np.abs(np.array(errs) - ref) / ref) <= (1e-8 if x.dtype == np.float64 else 1e-4)
    for a, b in zip(cp[1], g.arrays(tag, "f")):
        assert rel_fro(host(a), b) <= (1e-6 if x.dtype == np.float64 else 1e-2)
        assert float(a.min()) >= 0.0


def test_hals_large_factor_many_ctas():
    """A factor with thousands of rows: several CTAs, the grid-wide stopping statistic, iteration count reported."""
    rng = np.random.RandomState(8)
    R, n = 32, 5000
    U = rng.random_sample((300, R)).astype(np.float32)
    M = rng.random_sample((300, n)).astype(np.float32)
    V0 = rng.random_sample((R, n)).astype(np.float32)
    UtM, UtU = U.T @ M, U.T @ U
    ref = O.hals_nnls(UtM, UtU, V0, n_iter_max=30)
    iters = torch.zeros(1, dtype=torch.int32, device="cuda")
    f = dev(np.ascontiguousarray(V0.T))
    tb.hals_update([dev(UtU)], -1, None, dev(np.ascontiguousarray(UtM.T)), f, n_iter_max=30, iters_out=iters)
    assert rel_fro(host(f).T, ref) <= 2e-4
    assert 1 <= int(iters[0]) <= 30


def test_own_tucker_c3_sized_random_init_vs_exact_fp64_hooi():
    """256^3, ranks 48 (no room to oversample beyond 64 columns) from the reference's random init: the own driver
    against exact HOOI with an fp64 SVD of the projected unfolding (library SVD as the checker) — 1e-4 on every
    sweep, including the first one, where the all-positive random factors make the projected tensor almost rank one."""
    n, R, sweeps = 256, 48, 4
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand((n, n, n), generator=g, device="cuda")
    rs = np.random.RandomState(1)
    rs.random_sample([R, R, R])
    fs = [torch.as_tensor(rs.random_sample((n, R))).cuda().float() for _ in range(3)]
    nx2 = float(tb.sumsq(x))
    exact = []
    for _ in range(sweeps):
        for k in range(3):
            y = tb.multi_mode_dot(x, fs, skip=k, transpose=True)
            u, _, _ = torch.linalg.svd(tb.unfold(y, k, contiguous=True).double(), full_matrices=False)
            fs[k] = u[:, :R].float().contiguous()
        core = tb.multi_mode_dot(x, fs, transpose=True)
        exact.append((abs(nx2 - float(tb.sumsq(core))) / nx2) ** 0.5)
    _, errs = tb.tucker(x, [R, R, R], n_iter_max=sweeps, init="random", random_state=1, tol=0, return_errors=True)
    dev_ = max(abs(a - b) / b for a, b in zip(errs, exact))
    assert dev_ <= 1e-4, (errs, exact)


# --------------------------------------------------------------------------- fused right-hand sides (round 2)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,rank", [((96, 80, 112), 16), ((300, 64, 150), 32), ((256, 200, 260), 64), ((40, 24, 20, 12), 8),
                                        ((130, 70, 45), 10)])
def test_fused_update_equals_unfused(shape, rank, dtype):
    """mttkrp_partials + cp_update_fused (+ cp_error_iprod) against mttkrp + cp_update + cp_error: same factor rows, Gram,
    summed MTTKRP, inner product and error — the partial sums are only added in a different (fixed) order."""
    rng = np.random.RandomState(23)
    x = dev(rng.random_sample(shape).astype(dtype))
    fs = [dev(rng.random_sample((s, rank)).astype(dtype)) for s in shape]
    w = torch.ones(rank, dtype=x.dtype, device="cuda")
    grams = [tb.gram(f) for f in fs]
    nx2 = tb.sumsq(x)
    tol = 2e-5 if dtype == np.float32 else 1e-11
    t = tb.mode_dot(x, fs[-1], len(shape) - 1, transpose=True)
    for mode in range(len(shape)):
        m = tb.unfolding_dot_khatri_rao(x, (None, fs), mode)
        g_ref = torch.empty((rank, rank), dtype=x.dtype, device="cuda")
        f_ref = tb.cp_update(grams, mode, w, m, gram_out=g_ref)
        kinds = ["tensor"] + (["tree"] if mode < len(shape) - 1 and len(shape) >= 3 else [])
        for kind in kinds:
            # the partials live in the per-stream workspace: produce them right before the call that consumes them
            part = (tb._ops.mttkrp_partials(x, (None, fs), mode) if kind == "tensor"
                    else tb._ops.mttkrp_from_ttm_partials(t, (None, fs), mode))
            g_new = torch.empty((rank, rank), dtype=x.dtype, device="cuda")
            m_out = torch.empty_like(m)
            ip = torch.zeros(1, dtype=x.dtype, device="cuda")
            reg_lu = rank <= (64 if dtype == np.float32 else 32)       # <M, F> exists on the register-LU path only
            if not reg_lu:
                with pytest.raises(NotImplementedError):
                    tb._ops.cp_update_fused(grams, mode, w, part, gram_out=g_new, m_out=m_out, iprod_out=ip)
            f_new = tb._ops.cp_update_fused(grams, mode, w, part, gram_out=g_new, m_out=m_out, iprod_out=ip if reg_lu else None)
            assert rel_fro(host(m_out), host(m)) <= tol
            cond = float(torch.linalg.cond(g_ref.double()))
            assert rel_fro(host(f_new), host(f_ref)) <= max(tol, 50 * cond * np.finfo(dtype).eps)
            assert rel_fro(host(g_new), host(g_ref)) <= max(tol, 50 * cond * np.finfo(dtype).eps)
            if not reg_lu:
                continue
            ip_ref = float((m.double() * f_ref.double()).sum())
            assert abs(float(ip) - ip_ref) <= max(tol, 50 * cond * np.finfo(dtype).eps) * abs(ip_ref)
            new_grams = [g_new if i == mode else g for i, g in enumerate(grams)]
            e1 = host(tb._ops.cp_error_iprod(new_grams, w, ip, nx2))
            e2 = host(tb.cp_error(new_grams, w, m_out, f_new, nx2))
            assert abs(e1[2] - e2[2]) <= tol * abs(e2[2]) and abs(e1[1] - e2[1]) <= 10 * tol * abs(e2[1])
            if mode == len(shape) - 1:
                # the same launch can finish the error itself (last mode): identical inputs, same three scalars
                part = tb._ops.mttkrp_partials(x, (None, fs), mode)
                e3 = torch.zeros(3, dtype=x.dtype, device="cuda")
                g3 = torch.empty_like(g_new)
                ip3 = torch.zeros_like(ip)
                tb._ops.cp_update_fused(grams, mode, w, part, gram_out=g3, iprod_out=ip3, norm_x2=nx2, err_out=e3)
                e3 = host(e3)
                assert abs(e3[1] - e1[1]) <= 1e-6 * abs(e1[1]) and abs(e3[2] - e1[2]) <= 1e-6 * abs(e1[2])
                assert abs(e3[0] - e1[0]) <= 1e-5 * max(abs(e1[0]), 1e-3)


def test_fused_sweep_same_trajectory_as_unfused(monkeypatch):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((200, 180, 160), generator=g, device="cuda")
    fs = [torch.rand(s, 32, generator=g, device="cuda") for s in x.shape]
    w = torch.ones(32, device="cuda")
    runs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TLB200_FUSED_UPDATE", flag)
        st = tb.CPALS(x, w, fs)
        assert st._fuse == (flag == "1")
        errs = []
        for _ in range(6):
            st.sweep(True)
            errs.append(float(st.err[0]))
        runs.append((errs, [f.clone() for f in st.factors], tb.launch_count()))
    assert max(abs(a - b) / b for a, b in zip(runs[0][0], runs[1][0])) <= 1e-5
    for a, b in zip(runs[0][1], runs[1][1]):
        assert float(torch.linalg.norm(a - b) / torch.linalg.norm(b)) <= 1e-3


# ---- fp16-split tensor-core engine (range hint) ---------------------------------------------------------------
def _kr_ref64(x, fs, mode):
    """MTTKRP in fp64 on the device (einsum over the other modes)."""
    xd = x.double()
    letters = "abcde"[: x.dim()]
    ops, subs = [xd], [letters]
    for i, f in enumerate(fs):
        if i != mode:
            ops.append(f.double()); subs.append(letters[i] + "r")
    return torch.einsum(",".join(subs) + "->" + letters[mode] + "r", *ops)


@pytest.mark.parametrize("shape,rank", [((256, 192, 320), 64), ((256, 192, 320), 32), ((130, 96, 200), 48),
                                        ((64, 48, 40, 64), 64), ((256, 192, 320), 96)])
@pytest.mark.parametrize("data", ["uniform", "zero_mean", "wide"])
def test_fp16_engine_mttkrp_matches_fp64(shape, rank, data):
    """With a registered range hint MTTKRP runs on the fp16-split engine; same 1e-5 gate as 3xTF32 — on uniform data,
    zero-mean data, and data / factor columns spanning many orders of magnitude (per-tensor and per-column scales)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(shape, generator=g, device="cuda")
    fs = [torch.rand(s, rank, generator=g, device="cuda") for s in shape]
    if data != "uniform":
        x = x - 0.5
        fs = [f - 0.5 for f in fs]
    if data == "wide":
        x = x * 3.7e-10
        fs = [f * torch.logspace(-6, 5, rank, device="cuda")[None, :] for f in fs]
    x = x.contiguous()
    hint = tb.RangeHint(x)
    try:
        assert float(hint.absmax) == float(x.abs().max())
        for mode in range(len(shape)):
            got = tb.unfolding_dot_khatri_rao(x, (None, fs), mode)
            path = tb.last_kernel_path()
            ref = _kr_ref64(x, fs, mode)
            # column-wise: every rank-one component is held to the gate on its own (they differ by many orders of magnitude in "wide")
            err = float(((got.double() - ref).norm(dim=0) / ref.norm(dim=0)).max())
            assert err <= 1e-5, (mode, path, err)
            if shape == (256, 192, 320) and rank <= 64:          # every mode of this shape streams 64-element tiles
                assert path == "tcgen05-f16", (mode, path)
    finally:
        hint.close()
    tb.unfolding_dot_khatri_rao(x, (None, fs), 0)
    assert tb.last_kernel_path() != "tcgen05-f16"


@pytest.mark.parametrize("shape,mode,rows", [((320, 256, 192), 2, 64), ((320, 256, 192), 0, 64), ((320, 256, 192), 1, 48),
                                             ((96, 2048), 1, 64)])
def test_fp16_engine_mode_dot_matches_fp64(shape, mode, rows):
    g = torch.Generator(device="cuda").manual_seed(12)
    x = (torch.rand(shape, generator=g, device="cuda") - 0.5).contiguous()
    m = (torch.rand(rows, shape[mode], generator=g, device="cuda") - 0.5) * torch.logspace(-6, 6, rows, device="cuda")[:, None]
    want = torch.tensordot(m.double(), x.double(), dims=([1], [mode])).movedim(0, mode)
    base = tb.mode_dot(x, m, mode)
    hint = tb.RangeHint(x)
    try:
        got = tb.mode_dot(x, m, mode)
        path = tb.last_kernel_path()
    finally:
        hint.close()
    sl = [slice(None)] * len(shape)
    for i in (0, rows // 2, rows - 1):        # row-wise: the rows of m differ by 1e12 in scale
        sl[mode] = i
        e = float((got[tuple(sl)].double() - want[tuple(sl)]).norm() / want[tuple(sl)].norm())
        e0 = float((base[tuple(sl)].double() - want[tuple(sl)]).norm() / want[tuple(sl)].norm())
        assert e <= 1e-5, (i, path, e, e0)


def test_fp16_engine_in_the_driver():
    """CPALS registers the hint itself at rank > 32: same trajectory as the 3xTF32 sweeps."""
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.rand((256, 224, 192), generator=g, device="cuda")
    fs = [torch.rand(s, 64, generator=g, device="cuda") for s in x.shape]
    w = torch.ones(64, device="cuda")
    runs = []
    for off in ("0", "1"):
        os.environ["TLB200_HF_MIN_RANK"] = "1" if off == "0" else "1000"
        try:
            st = tb.CPALS(x, w, fs)
            assert (st._range_hint is not None) == (off == "0")
            errs = []
            for _ in range(5):
                st.sweep(True)
                errs.append(float(st.err[0]))
            runs.append(errs)
            del st
        finally:
            os.environ.pop("TLB200_HF_MIN_RANK", None)
    assert max(abs(a - b) / b for a, b in zip(*runs)) <= 1e-5, runs


def test_range_hint_survives_a_second_owner():
    """Two drivers on one tensor: the hint stays registered until the last of them lets go."""
    g = torch.Generator(device="cuda").manual_seed(14)
    x = torch.rand((256, 192, 320), generator=g, device="cuda")
    fs = [torch.rand(s, 64, generator=g, device="cuda") for s in x.shape]
    a = tb.RangeHint(x)
    b = tb.RangeHint(x)
    b.close()
    tb.unfolding_dot_khatri_rao(x, (None, fs), 1)
    assert tb.last_kernel_path() == "tcgen05-f16"
    c = tb.RangeHint(x)
    a.close()
    tb.unfolding_dot_khatri_rao(x, (None, fs), 1)
    assert tb.last_kernel_path() == "tcgen05-f16"
    c.close()
    tb.unfolding_dot_khatri_rao(x, (None, fs), 1)
    assert tb.last_kernel_path() == "tcgen05"


def test_backend_registers_the_range_hint_on_second_sight(tl_b200):
    """Behind the stateless tenalg API: the first MTTKRP on a tensor runs 3xTF32, the second call that sees the same
    object at the same `_version` has measured max |x| and runs the fp16 split; an in-place edit drops the hint."""
    from tensorly.tenalg import unfolding_dot_khatri_rao as tl_mttkrp
    g = torch.Generator(device="cuda").manual_seed(15)
    x = torch.rand((256, 192, 320), generator=g, device="cuda")
    fs = [torch.rand(s, 32, generator=g, device="cuda") for s in x.shape]
    ref = _kr_ref64(x, fs, 1)
    paths = []
    for _ in range(3):
        got = tl_mttkrp(x, (None, fs), 1)
        paths.append(tb.last_kernel_path())
        assert float((got.double() - ref).norm() / ref.norm()) <= 1e-5
    assert paths == ["tcgen05", "tcgen05-f16", "tcgen05-f16"], paths
    x.mul_(1e6)                       # _version changes: the old max |x| must not be used
    got = tl_mttkrp(x, (None, fs), 1)
    assert tb.last_kernel_path() == "tcgen05"
    assert float((got.double() - 1e6 * ref).norm() / (1e6 * ref).norm()) <= 1e-5
    got = tl_mttkrp(x, (None, fs), 1)
    assert tb.last_kernel_path() == "tcgen05-f16"
    assert float((got.double() - 1e6 * ref).norm() / (1e6 * ref).norm()) <= 1e-5
    y = x.clone()
    tl_mttkrp(y, (None, fs), 1)       # another tensor: the hint for x is withdrawn
    assert tb.last_kernel_path() == "tcgen05"
