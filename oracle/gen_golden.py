"""Generate tests/golden/*.npz by running the REAL reference (tensorly v0.9.0).

Run in the build container only (needs /root/reference or baseline/_ref):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

The fixtures hold inputs (or the seeds that make them) and the reference's outputs on
the numpy backend with the `core` tenalg; `tests/test_oracle.py` pins the oracle
restatement against them and the `-m gpu` tests compare the CUDA path with them on
the GPU box, where the reference source is not available.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
    if os.path.isdir(os.path.join(cand, "tensorly")):
        sys.path.insert(0, cand)
        break
import tensorly as tl  # noqa: E402
from tensorly import random as tlrandom  # noqa: E402
from tensorly.cp_tensor import CPTensor  # noqa: E402
from tensorly.decomposition import non_negative_parafac, parafac, tucker  # noqa: E402
from tensorly.tenalg import khatri_rao, mode_dot, multi_mode_dot, unfolding_dot_khatri_rao  # noqa: E402

tl.set_backend("numpy")
tl.tenalg.set_backend("core")
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path)/1024:.1f} KiB, {len(arrays)} arrays")


def gen_unfold():
    d = {}
    x = np.arange(24, dtype=np.float64).reshape(3, 4, 2)
    d["a/x"] = x
    for m in range(3):
        d[f"a/unfold{m}"] = np.ascontiguousarray(tl.unfold(x, m))
    rng = np.random.RandomState(7)
    for tag, shape, dt in (("b", (3, 4, 5, 2), np.float32), ("c", (5, 1, 6), np.float64), ("d", (2, 3, 4, 5, 3), np.float32)):
        x = rng.random_sample(shape).astype(dt)
        d[f"{tag}/x"] = x
        for m in range(len(shape)):
            u = np.ascontiguousarray(tl.unfold(x, m))
            d[f"{tag}/unfold{m}"] = u
            assert np.array_equal(tl.fold(u, m, shape), x)
    save("unfold", **d)


def gen_khatri_rao():
    d = {}
    rng = np.random.RandomState(11)
    cases = {
        "a": dict(rows=(3, 3), rank=3, dt=np.float64, weights=False, mask=False, skip=None),
        "b": dict(rows=(4, 5, 3), rank=6, dt=np.float32, weights=True, mask=False, skip=None),
        "c": dict(rows=(4, 5, 3, 2), rank=3, dt=np.float32, weights=True, mask=False, skip=1),
        "d": dict(rows=(6, 7), rank=5, dt=np.float64, weights=True, mask=True, skip=None),
        "e": dict(rows=(33, 17, 9), rank=37, dt=np.float32, weights=True, mask=False, skip=None),
        "f": dict(rows=(2, 3, 4, 2, 3), rank=4, dt=np.float64, weights=False, mask=False, skip=4),
    }
    for tag, c in cases.items():
        mats = [(rng.random_sample((r, c["rank"])) - 0.3).astype(c["dt"]) for r in c["rows"]]
        w = (rng.random_sample(c["rank"]) + 0.5).astype(c["dt"]) if c["weights"] else None
        rows_after = [r for i, r in enumerate(c["rows"]) if i != c["skip"]]
        mask = (rng.random_sample(int(np.prod(rows_after))) > 0.3).astype(c["dt"]) if c["mask"] else None
        out = khatri_rao(mats, weights=w, skip_matrix=c["skip"], mask=mask)
        for i, m in enumerate(mats):
            d[f"{tag}/m{i}"] = m
        if w is not None:
            d[f"{tag}/w"] = w
        if mask is not None:
            d[f"{tag}/mask"] = mask
        d[f"{tag}/skip"] = np.array(-1 if c["skip"] is None else c["skip"])
        d[f"{tag}/out"] = out
    save("khatri_rao", **d)


def gen_mttkrp():
    d = {}
    rng = np.random.RandomState(3)
    cases = {
        "a": dict(shape=(10, 10, 10, 4), rank=5, dt=np.float64, weights=True),   # reference test shape
        "b": dict(shape=(7, 8, 9), rank=4, dt=np.float32, weights=False),
        "c": dict(shape=(33, 20, 41), rank=32, dt=np.float32, weights=True),
        "d": dict(shape=(16, 5, 3, 4, 6), rank=3, dt=np.float64, weights=True),
        "e": dict(shape=(130, 70), rank=9, dt=np.float32, weights=True),
        "f": dict(shape=(40, 36, 44), rank=64, dt=np.float32, weights=False),
        "g": dict(shape=(24, 20, 12, 16), rank=10, dt=np.float32, weights=True),
    }
    for tag, c in cases.items():
        x = (rng.random_sample(c["shape"]) - (0.5 if tag in "cf" else 0.0)).astype(c["dt"])
        fs = [(rng.random_sample((s, c["rank"])) - 0.25).astype(c["dt"]) for s in c["shape"]]
        w = (rng.random_sample(c["rank"]) + 0.5).astype(c["dt"]) if c["weights"] else None
        d[f"{tag}/x"] = x
        for i, f in enumerate(fs):
            d[f"{tag}/f{i}"] = f
        if w is not None:
            d[f"{tag}/w"] = w
        for m in range(len(c["shape"])):
            d[f"{tag}/out{m}"] = unfolding_dot_khatri_rao(x, (w, fs), m)
    save("mttkrp", **d)


def gen_mode_dot():
    d = {}
    rng = np.random.RandomState(5)
    x = np.arange(24, dtype=np.float64).reshape(3, 4, 2)
    d["a/x"] = x
    u = np.array([[1, 2], [3, 4], [5, 6]], dtype=np.float64).T  # (2,3) on mode 0
    d["a/m"] = u
    d["a/out"] = mode_dot(x, u, 0)
    v = np.array([1.0, 2.0])
    d["a/v"] = v
    d["a/outv"] = mode_dot(x, v, 2)
    cases = {
        "b": dict(shape=(6, 7, 8), J=5, dt=np.float32),
        "c": dict(shape=(9, 4, 3, 5), J=11, dt=np.float64),
        "d": dict(shape=(40, 33, 21), J=16, dt=np.float32),
    }
    for tag, c in cases.items():
        x = (rng.random_sample(c["shape"]) - 0.5).astype(c["dt"])
        d[f"{tag}/x"] = x
        for m, s in enumerate(c["shape"]):
            mat = (rng.random_sample((c["J"], s)) - 0.5).astype(c["dt"])
            vec = (rng.random_sample(s) - 0.5).astype(c["dt"])
            d[f"{tag}/m{m}"] = mat
            d[f"{tag}/v{m}"] = vec
            d[f"{tag}/out{m}"] = np.ascontiguousarray(mode_dot(x, mat, m))
            d[f"{tag}/outT{m}"] = np.ascontiguousarray(mode_dot(x, np.ascontiguousarray(mat.T), m, transpose=True))
            d[f"{tag}/outv{m}"] = np.ascontiguousarray(mode_dot(x, vec, m))
    save("mode_dot", **d)


def gen_multi_mode_dot():
    d = {}
    rng = np.random.RandomState(9)
    cases = {
        "a": dict(shape=(5, 6, 7), ranks=(3, 4, 2), dt=np.float64),
        "b": dict(shape=(12, 10, 9, 8), ranks=(4, 3, 5, 2), dt=np.float32),
        "c": dict(shape=(32, 40, 24), ranks=(8, 16, 8), dt=np.float32),
    }
    for tag, c in cases.items():
        x = (rng.random_sample(c["shape"]) - 0.5).astype(c["dt"])
        # factors as (I_n, R_n), used with transpose=True like HOOI (_tucker.py:194-196)
        fs = [np.asfortranarray((rng.random_sample((s, r)) - 0.5).astype(c["dt"])) for s, r in zip(c["shape"], c["ranks"])]
        d[f"{tag}/x"] = x
        for i, f in enumerate(fs):
            d[f"{tag}/f{i}"] = np.ascontiguousarray(f)
        d[f"{tag}/full"] = np.ascontiguousarray(multi_mode_dot(x, fs, transpose=True))
        for k in range(len(c["shape"])):
            d[f"{tag}/skip{k}"] = np.ascontiguousarray(multi_mode_dot(x, fs, skip=k, transpose=True))
        # subset of modes, not transposed, given out of order
        ms = [np.ascontiguousarray(fs[2].T), np.ascontiguousarray(fs[0].T)]
        d[f"{tag}/sub20"] = np.ascontiguousarray(multi_mode_dot(x, ms, modes=[2, 0]))
        # vectors on modes 0 and 2
        vs = [(rng.random_sample(c["shape"][0]) - 0.5).astype(c["dt"]), (rng.random_sample(c["shape"][2]) - 0.5).astype(c["dt"])]
        d[f"{tag}/vec0"] = vs[0]
        d[f"{tag}/vec2"] = vs[1]
        d[f"{tag}/vecs02"] = np.ascontiguousarray(multi_mode_dot(x, vs, modes=[0, 2]))
    save("multi_mode_dot", **d)


def gen_als():
    d = {}
    # parafac: identical inputs and identical initial factors (random_cp semantics)
    for tag, shape, rank, dt, iters in (
        ("p32", (30, 25, 20), 4, np.float32, 10),
        ("p64", (30, 25, 20), 4, np.float64, 10),
        ("c1", (100, 100, 100), 10, np.float64, 10),     # BASELINE config 1
        ("p4way", (12, 10, 9, 8), 3, np.float64, 8),
    ):
        x = tlrandom.random_tensor(shape, random_state=0).astype(dt)
        init = tlrandom.random_cp(shape, rank, random_state=1, normalise_factors=False)
        init = CPTensor((init.weights.astype(dt), [f.astype(dt) for f in init.factors]))
        cp, errs = parafac(x, rank, n_iter_max=iters, init=init.cp_copy(), tol=0, return_errors=True)
        d[f"{tag}/shape"] = np.array(shape)
        d[f"{tag}/rank"] = np.array(rank)
        d[f"{tag}/iters"] = np.array(iters)
        d[f"{tag}/errors"] = np.array([float(e) for e in errs])
        if tag != "c1":
            d[f"{tag}/x"] = x
            for i, f in enumerate(init.factors):
                d[f"{tag}/init{i}"] = f
            for i, f in enumerate(cp.factors):
                d[f"{tag}/f{i}"] = f
        else:
            d[f"{tag}/x_checksum"] = np.array([x.sum(), (x * x).sum(), x[3, 5, 7]])
            d[f"{tag}/f0_head"] = cp.factors[0][:4]
    # low-rank + noise tensor so that the error is informative
    rng = np.random.RandomState(21)
    shape, rank = (28, 24, 26), 5
    gt = tlrandom.random_cp(shape, rank, random_state=4, normalise_factors=False)
    x = (tl.cp_to_tensor(gt) + 0.01 * rng.standard_normal(shape)).astype(np.float64)
    init = tlrandom.random_cp(shape, rank, random_state=1, normalise_factors=False)
    cp, errs = parafac(x, rank, n_iter_max=8, init=init.cp_copy(), tol=0, return_errors=True)
    d["lowrank/x"] = x
    for i, f in enumerate(init.factors):
        d[f"lowrank/init{i}"] = f
    d["lowrank/errors"] = np.array([float(e) for e in errs])

    # non-negative parafac (MU), 4-way like config 4 in miniature; tol tiny-but-truthy so
    # that errors are evaluated (_nn_cp.py:140)
    shape, rank = (12, 10, 9, 8), 6
    x = tlrandom.random_tensor(shape, random_state=0).astype(np.float32)
    init = tlrandom.random_cp(shape, rank, random_state=1, normalise_factors=False)
    init = CPTensor((init.weights.astype(np.float32), [f.astype(np.float32) for f in init.factors]))
    cp, errs = non_negative_parafac(x, rank, n_iter_max=10, init=init.cp_copy(), tol=1e-30, return_errors=True)
    d["nn/x"] = x
    for i, f in enumerate(init.factors):
        d[f"nn/init{i}"] = f
    for i, f in enumerate(cp.factors):
        d[f"nn/f{i}"] = f
    d["nn/errors"] = np.array([float(e) for e in errs])

    # tucker HOOI, random init (QR-free: _tucker.py:81-93 uses random_sample factors)
    shape, ranks = (20, 22, 24), [4, 5, 6]
    x = tlrandom.random_tensor(shape, random_state=0).astype(np.float64)
    (core, factors), errs = tucker(x, ranks, n_iter_max=5, init="random", random_state=1, tol=0, return_errors=True)
    d["tucker/x"] = x
    d["tucker/ranks"] = np.array(ranks)
    d["tucker/errors"] = np.array([float(e) for e in errs])
    d["tucker/core_norm"] = np.array(float(tl.norm(core, 2)))
    rs = tl.check_random_state(1)
    rs.random_sample(ranks)  # initialize_tucker draws the core first (_tucker.py:83-86)
    init_factors = [np.array(rs.random_sample((s, r))) for s, r in zip(shape, ranks)]
    for i, f in enumerate(init_factors):
        d[f"tucker/init{i}"] = f
    save("als", **d)


def gen_round2():
    """Round-2 fixtures (SURVEY 8(f) n4 + the NNDSVD initialisation): reconstruction, masked ALS, HALS."""
    from tensorly.decomposition import non_negative_parafac_hals
    from tensorly.decomposition._cp import initialize_cp
    from tensorly.solvers.nnls import hals_nnls
    d = {}
    rng = np.random.RandomState(31)
    # cp_to_tensor (tensorly/cp_tensor.py:433-485)
    for tag, shape, rank, dt, wts in (("rec_a", (9, 8, 7), 4, np.float64, True), ("rec_b", (40, 33, 21), 16, np.float32, False),
                                      ("rec_c", (12, 10, 9, 8), 5, np.float32, True), ("rec_d", (130, 70), 9, np.float64, True),
                                      ("rec_e", (7,), 3, np.float64, True), ("rec_f", (33, 5, 3, 4, 6), 40, np.float32, True)):
        fs = [(rng.random_sample((s, rank)) - 0.3).astype(dt) for s in shape]
        w = (rng.random_sample(rank) + 0.5).astype(dt) if wts else None
        for i, f in enumerate(fs):
            d[f"{tag}/f{i}"] = f
        if w is not None:
            d[f"{tag}/w"] = w
        d[f"{tag}/out"] = np.ascontiguousarray(tl.cp_to_tensor((w, fs)))
    # masked parafac (decomposition/_cp.py:195-207, :442-445): 20 % of the entries missing
    for tag, shape, rank, dt, iters in (("mask64", (18, 15, 12), 3, np.float64, 8), ("mask32", (30, 25, 20), 4, np.float32, 6),
                                        ("mask4way", (10, 9, 8, 7), 3, np.float64, 5)):
        gt = tlrandom.random_cp(shape, rank, random_state=5, normalise_factors=False)
        x = (tl.cp_to_tensor(gt) + 0.01 * rng.standard_normal(shape)).astype(dt)
        mask = (rng.random_sample(shape) > 0.2).astype(dt)
        x = x * mask                      # missing entries start at zero
        init = tlrandom.random_cp(shape, rank, random_state=1, normalise_factors=False)
        init = CPTensor((init.weights.astype(dt), [f.astype(dt) for f in init.factors]))
        cp, errs = parafac(x, rank, n_iter_max=iters, init=init.cp_copy(), tol=0, return_errors=True, mask=mask)
        d[f"{tag}/x"] = x
        d[f"{tag}/mask"] = mask
        d[f"{tag}/rank"] = np.array(rank)
        d[f"{tag}/iters"] = np.array(iters)
        for i, f in enumerate(init.factors):
            d[f"{tag}/init{i}"] = f
        for i, f in enumerate(cp.factors):
            d[f"{tag}/f{i}"] = f
        d[f"{tag}/errors"] = np.array([float(e) for e in errs])
    # hals_nnls (solvers/nnls.py:5-175) on its own
    for tag, r, n, dt in (("hals_a", 6, 40, np.float64), ("hals_b", 32, 300, np.float32), ("hals_c", 64, 129, np.float32)):
        U = rng.random_sample((80, r)).astype(dt)
        M = rng.random_sample((80, n)).astype(dt)
        V0 = rng.random_sample((r, n)).astype(dt)
        UtM, UtU = U.T @ M, U.T @ U
        d[f"{tag}/UtM"], d[f"{tag}/UtU"], d[f"{tag}/V0"] = UtM, UtU, V0
        d[f"{tag}/V"] = hals_nnls(UtM.copy(), UtU.copy(), V0.copy(), n_iter_max=100)
        d[f"{tag}/V_sparse"] = hals_nnls(UtM.copy(), UtU.copy(), V0.copy(), n_iter_max=20, sparsity_coefficient=0.05,
                                         ridge_coefficient=0.1, epsilon=1e-6)
    # non_negative_parafac_hals (decomposition/_nn_cp.py:186-379)
    for tag, shape, rank, dt, iters in (("nnhals64", (14, 12, 10), 4, np.float64, 6), ("nnhals32", (24, 20, 16), 5, np.float32, 5)):
        x = tlrandom.random_tensor(shape, random_state=0).astype(dt)
        init = tlrandom.random_cp(shape, rank, random_state=1, normalise_factors=False)
        init = CPTensor((init.weights.astype(dt), [f.astype(dt) for f in init.factors]))
        cp, errs = non_negative_parafac_hals(x, rank, n_iter_max=iters, init=init.cp_copy(), tol=1e-30, return_errors=True)
        d[f"{tag}/x"] = x
        d[f"{tag}/rank"] = np.array(rank)
        d[f"{tag}/iters"] = np.array(iters)
        for i, f in enumerate(init.factors):
            d[f"{tag}/init{i}"] = f
        for i, f in enumerate(cp.factors):
            d[f"{tag}/f{i}"] = f
        d[f"{tag}/errors"] = np.array([float(e) for e in errs])
    # non_negative_parafac with the default init='svd' (NNDSVDA through svd_interface, tenalg/svd.py:68-135)
    shape, rank = (16, 14, 12), 4
    x = tlrandom.random_tensor(shape, random_state=3).astype(np.float64)
    kt = initialize_cp(x, rank, init="svd", non_negative=True)
    cp, errs = non_negative_parafac(x, rank, n_iter_max=8, init="svd", tol=1e-30, return_errors=True)
    d["nnsvd/x"] = x
    d["nnsvd/rank"] = np.array(rank)
    for i, f in enumerate(kt.factors):
        d[f"nnsvd/init{i}"] = np.array(f)
    d["nnsvd/errors"] = np.array([float(e) for e in errs])
    # tucker with init='svd' (the default) for the own HOOI driver; tucker_to_tensor of the result
    shape, ranks = (24, 20, 22), [5, 4, 6]
    gtc = rng.standard_normal(ranks)
    gtf = [np.linalg.qr(rng.standard_normal((s, r)))[0] for s, r in zip(shape, ranks)]
    x = (multi_mode_dot(gtc, gtf) + 0.05 * rng.standard_normal(shape)).astype(np.float64)
    (core, factors), errs = tucker(x, ranks, n_iter_max=6, init="svd", tol=0, return_errors=True)
    d["tucker_svd/x"] = x
    d["tucker_svd/ranks"] = np.array(ranks)
    d["tucker_svd/errors"] = np.array([float(e) for e in errs])
    d["tucker_svd/rec"] = np.ascontiguousarray(tl.tucker_to_tensor((core, factors)))
    x2 = tlrandom.random_tensor((26, 24, 28), random_state=0).astype(np.float32)
    (core, factors), errs = tucker(x2, [6, 5, 7], n_iter_max=8, init="svd", tol=0, return_errors=True)
    d["tucker_svd32/x"] = x2
    d["tucker_svd32/ranks"] = np.array([6, 5, 7])
    d["tucker_svd32/errors"] = np.array([float(e) for e in errs])
    save("round2", **d)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":      # the round-1 fixtures stay byte-identical
        gen_round2()
        sys.exit(0)
    gen_unfold()
    gen_khatri_rao()
    gen_mttkrp()
    gen_mode_dot()
    gen_multi_mode_dot()
    gen_als()
