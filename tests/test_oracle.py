"""Pin the CPU oracle (oracle/oracle.py) against (1) the golden vectors the reference's own
tests hold for the hot path and (2) outputs of the real reference committed under
tests/golden/ by oracle/gen_golden.py.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from conftest import rel_fro


# ---- (1) golden vectors restated from the reference's tests ---------------------------
def test_unfold_fold_reference_vectors():
    # tensorly/tests/test_core.py:14-60 — arange(24).reshape(3,4,2)
    x = np.arange(24).reshape(3, 4, 2)
    u0 = np.array([[0, 1, 2, 3, 4, 5, 6, 7], [8, 9, 10, 11, 12, 13, 14, 15], [16, 17, 18, 19, 20, 21, 22, 23]])
    u1 = np.array([[0, 1, 8, 9, 16, 17], [2, 3, 10, 11, 18, 19], [4, 5, 12, 13, 20, 21], [6, 7, 14, 15, 22, 23]])
    u2 = np.array([[0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22], [1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23]])
    for mode, u in enumerate((u0, u1, u2)):
        assert np.array_equal(O.unfold(x, mode), u)
        assert np.array_equal(O.fold(u, mode, x.shape), x)


def test_khatri_rao_reference_vectors():
    # tensorly/tenalg/tests/test_khatri_rao.py:34-51 — the classic 3x3 example
    t1 = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    t2 = np.array([[1, 4, 7], [2, 5, 8], [3, 6, 9]])
    true = np.array([[1, 8, 21], [2, 10, 24], [3, 12, 27], [4, 20, 42], [8, 25, 48], [12, 30, 54],
                     [7, 32, 63], [14, 40, 72], [21, 48, 81]])
    assert np.array_equal(O.khatri_rao([t1, t2]), true)
    # :131-140 — skip_matrix equivalences and single-matrix passthrough
    rng = np.random.RandomState(0)
    mats = [rng.random_sample((n, 3)) for n in (4, 5, 2, 3)]
    for skip in range(4):
        kept = [m for i, m in enumerate(mats) if i != skip]
        assert np.array_equal(O.khatri_rao(mats, skip_matrix=skip), O.khatri_rao(kept))
    assert O.khatri_rao([mats[0]]) is mats[0]
    # :23-32 — ValueError cases
    with pytest.raises(ValueError):
        O.khatri_rao([rng.random_sample((3, 4)), rng.random_sample((3, 5))])
    with pytest.raises(ValueError):
        O.khatri_rao([rng.random_sample((3, 4)), rng.random_sample((3, 4, 2))])


def test_mode_dot_reference_vectors():
    # tensorly/tenalg/tests/test_n_mode_product.py:21-52
    x = np.array([[[1, 13], [4, 16], [7, 19], [10, 22]], [[2, 14], [5, 17], [8, 20], [11, 23]],
                  [[3, 15], [6, 18], [9, 21], [12, 24]]])
    u = np.array([[1, 3, 5], [2, 4, 6]])
    true = np.array([[[22, 130], [49, 157], [76, 184], [103, 211]], [[28, 172], [64, 208], [100, 244], [136, 280]]])
    assert np.array_equal(O.mode_dot(x, u, 0), true)
    v = np.array([1, 2, 3, 4])
    true_v = np.array([[70, 190], [80, 200], [90, 210]])
    assert np.array_equal(O.mode_dot(x, v, 1), true_v)
    with pytest.raises(ValueError):
        O.mode_dot(x, np.ones((2, 5)), 0)
    with pytest.raises(ValueError):
        O.mode_dot(x, np.ones(5), 1)


def test_multi_mode_dot_identities():
    # test_n_mode_product.py:99-153 — kron identity, skip, vector order independence
    rng = np.random.RandomState(1)
    x = rng.random_sample((3, 4, 5))
    us = [rng.random_sample((2, 3)), rng.random_sample((3, 4)), rng.random_sample((4, 5))]
    full = O.multi_mode_dot(x, us)
    ref = us[0] @ O.unfold(x, 0) @ np.kron(us[1], us[2]).T
    np.testing.assert_allclose(O.unfold(full, 0), ref, rtol=1e-12)
    np.testing.assert_allclose(O.multi_mode_dot(x, us, skip=1), O.mode_dot(O.mode_dot(x, us[0], 0), us[2], 2), rtol=1e-12)
    vs = [rng.random_sample(3), rng.random_sample(4), rng.random_sample(5)]
    np.testing.assert_allclose(O.multi_mode_dot(x, vs), np.einsum("ijk,i,j,k->", x, *vs), rtol=1e-12)


def test_mttkrp_identity():
    # tenalg/tests/test_unfolding_dot_khatri_rao.py:11-31
    rng = np.random.RandomState(2)
    shape, rank = (10, 10, 10, 4), 5
    x = rng.random_sample(shape)
    fs = [rng.random_sample((s, rank)) for s in shape]
    w = rng.random_sample(rank)
    for mode in range(4):
        full = O.cp_to_tensor((w, fs))
        true = O.unfold(x, mode) @ O.khatri_rao(fs, weights=w, skip_matrix=mode)
        np.testing.assert_allclose(O.unfolding_dot_khatri_rao(x, (w, fs), mode), true, rtol=1e-12)
        assert full.shape == shape


# ---- (2) outputs of the real reference ------------------------------------------------
def test_unfold_vs_reference(golden):
    g = golden("unfold")
    for case in g.cases():
        x = g[f"{case}/x"]
        for mode in range(x.ndim):
            u = g[f"{case}/unfold{mode}"]
            assert np.array_equal(O.unfold(x, mode), u)
            assert np.array_equal(O.fold(u, mode, x.shape), x)


def test_khatri_rao_vs_reference(golden):
    g = golden("khatri_rao")
    for case in g.cases():
        mats = g.arrays(case, "m")
        w = g[f"{case}/w"] if g.has(f"{case}/w") else None
        mask = g[f"{case}/mask"] if g.has(f"{case}/mask") else None
        skip = int(g[f"{case}/skip"])
        out = O.khatri_rao(mats, weights=w, skip_matrix=None if skip < 0 else skip, mask=mask)
        assert out.dtype == g[f"{case}/out"].dtype
        assert np.array_equal(out, g[f"{case}/out"])  # bit-exact


def test_mttkrp_vs_reference(golden):
    g = golden("mttkrp")
    for case in g.cases():
        x = g[f"{case}/x"]
        fs = g.arrays(case, "f")
        w = g[f"{case}/w"] if g.has(f"{case}/w") else None
        tol = 1e-12 if x.dtype == np.float64 else 1e-5
        for mode in range(x.ndim):
            assert rel_fro(O.unfolding_dot_khatri_rao(x, (w, fs), mode), g[f"{case}/out{mode}"]) <= tol


def test_mode_dot_vs_reference(golden):
    g = golden("mode_dot")
    assert np.array_equal(O.mode_dot(g["a/x"], g["a/m"], 0), g["a/out"])
    assert np.array_equal(O.mode_dot(g["a/x"], g["a/v"], 2), g["a/outv"])
    for case in ("b", "c", "d"):
        x = g[f"{case}/x"]
        tol = 1e-12 if x.dtype == np.float64 else 1e-5
        for mode in range(x.ndim):
            m, v = g[f"{case}/m{mode}"], g[f"{case}/v{mode}"]
            assert rel_fro(O.mode_dot(x, m, mode), g[f"{case}/out{mode}"]) <= tol
            assert rel_fro(O.mode_dot(x, np.ascontiguousarray(m.T), mode, transpose=True), g[f"{case}/outT{mode}"]) <= tol
            assert rel_fro(O.mode_dot(x, v, mode), g[f"{case}/outv{mode}"]) <= tol


def test_multi_mode_dot_vs_reference(golden):
    g = golden("multi_mode_dot")
    for case in g.cases():
        x = g[f"{case}/x"]
        fs = g.arrays(case, "f")
        tol = 1e-12 if x.dtype == np.float64 else 1e-5
        assert rel_fro(O.multi_mode_dot(x, fs, transpose=True), g[f"{case}/full"]) <= tol
        for k in range(x.ndim):
            assert rel_fro(O.multi_mode_dot(x, fs, skip=k, transpose=True), g[f"{case}/skip{k}"]) <= tol
        ms = [np.ascontiguousarray(fs[2].T), np.ascontiguousarray(fs[0].T)]
        assert rel_fro(O.multi_mode_dot(x, ms, modes=[2, 0]), g[f"{case}/sub20"]) <= tol
        out = O.multi_mode_dot(x, [g[f"{case}/vec0"], g[f"{case}/vec2"]], modes=[0, 2])
        assert out.shape == g[f"{case}/vecs02"].shape
        assert rel_fro(out, g[f"{case}/vecs02"]) <= tol


def test_generators_and_parafac_vs_reference(golden):
    g = golden("als")
    for tag in ("p32", "p64", "p4way"):
        shape = tuple(int(s) for s in g[f"{tag}/shape"])
        rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
        x = g[f"{tag}/x"]
        assert np.array_equal(O.random_tensor(shape, 0, x.dtype), x)               # generator restatement
        w, fs = O.random_cp_factors(shape, rank, 1, x.dtype)
        for a, b in zip(fs, g.arrays(tag, "init")):
            assert np.array_equal(a, b)
        (_, factors), errs = O.parafac(x, (w, fs), n_iter_max=iters)
        ref = g[f"{tag}/errors"]
        tol = 1e-10 if x.dtype == np.float64 else 1e-4
        assert np.max(np.abs(np.array(errs, dtype=np.float64) - ref) / ref) <= tol
        for a, b in zip(factors, g.arrays(tag, "f")):
            assert rel_fro(a, b) <= (1e-8 if x.dtype == np.float64 else 5e-3)


def test_c1_config_vs_reference(golden):
    """BASELINE config 1: parafac rank 10 on random 100x100x100 float64 (numpy reference run)."""
    g = golden("als")
    x = O.random_tensor((100, 100, 100), 0)
    chk = g["c1/x_checksum"]
    assert x.sum() == chk[0] and x[3, 5, 7] == chk[2]
    init = O.random_cp_factors((100, 100, 100), 10, 1)
    (_, factors), errs = O.parafac(x, init, n_iter_max=int(g["c1/iters"]))
    ref = g["c1/errors"]
    assert np.max(np.abs(np.array(errs) - ref) / ref) <= 1e-10
    assert rel_fro(factors[0][:4], g["c1/f0_head"]) <= 1e-7


def test_lowrank_nn_tucker_vs_reference(golden):
    g = golden("als")
    x = g["lowrank/x"]
    init = (np.ones(5), g.arrays("lowrank", "init"))
    _, errs = O.parafac(x, init, n_iter_max=8)
    ref = g["lowrank/errors"]
    assert np.max(np.abs(np.array(errs) - ref)) <= 1e-9
    x = g["nn/x"]
    init = (np.ones(6, dtype=np.float32), g.arrays("nn", "init"))
    (_, factors), errs = O.non_negative_parafac(x, init, n_iter_max=10)
    ref = g["nn/errors"]
    assert np.max(np.abs(np.array(errs, dtype=np.float64) - ref) / ref) <= 1e-4
    for a, b in zip(factors, g.arrays("nn", "f")):
        assert rel_fro(a, b) <= 1e-4
    x = g["tucker/x"]
    ranks = [int(r) for r in g["tucker/ranks"]]
    (core, _), errs = O.tucker_hooi(x, ranks, g.arrays("tucker", "init"), n_iter_max=5)
    assert np.max(np.abs(np.array(errs) - g["tucker/errors"])) <= 1e-10
    assert abs(O.tensor_norm(core) - float(g["tucker/core_norm"])) <= 1e-9


def test_gram_svd_plugin_matches_lapack_svd():
    """tensorly_b200.gram_svd (host-side plug-in, library GEMM + eigh): singular triplets of short-fat and
    tall-skinny matrices agree with LAPACK's SVD; full_matrices=True defers to the previous svd."""
    import torch
    from tensorly_b200.svd import gram_svd
    rng = np.random.RandomState(3)
    for shape in [(20, 300), (300, 20), (16, 16)]:
        a = torch.as_tensor(rng.standard_normal(shape))
        u, s, vh = gram_svd(a, full_matrices=False)
        _, s_ref, _ = torch.linalg.svd(a, full_matrices=False)
        k = min(shape)
        assert u.shape == (shape[0], k) and s.shape == (k,) and vh.shape == (k, shape[1])
        assert float((s - s_ref).abs().max() / s_ref.max()) < 1e-12
        assert float(torch.linalg.norm((u * s) @ vh - a) / torch.linalg.norm(a)) < 1e-10
        assert float(torch.linalg.norm(u.T @ u - torch.eye(k, dtype=a.dtype))) < 1e-8
    u, s, vh = gram_svd(torch.as_tensor(rng.standard_normal((6, 9))), full_matrices=True)
    assert u.shape == (6, 6) and vh.shape == (9, 9)


# --------------------------------------------------------------------------- round-2 fixtures
def test_cp_to_tensor_vs_reference(golden):
    g = golden("round2")
    for tag in ("rec_a", "rec_b", "rec_c", "rec_d", "rec_e", "rec_f"):
        fs = g.arrays(tag, "f")
        w = g[f"{tag}/w"] if g.has(f"{tag}/w") else None
        out = O.cp_to_tensor((w, fs))
        ref = g[f"{tag}/out"]
        assert out.shape == ref.shape
        assert np.array_equal(out, ref), tag          # same expression, same BLAS: bit for bit


def test_masked_parafac_vs_reference(golden):
    g = golden("round2")
    for tag in ("mask64", "mask32", "mask4way"):
        x, mask = g[f"{tag}/x"], g[f"{tag}/mask"]
        rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
        init = (np.ones(rank, dtype=x.dtype), g.arrays(tag, "init"))
        (_, factors), errs, _ = O.parafac_masked(x, mask, init, n_iter_max=iters)
        ref = g[f"{tag}/errors"]
        tol = 1e-9 if x.dtype == np.float64 else 1e-4
        assert np.max(np.abs(np.array(errs, dtype=np.float64) - ref) / ref) <= tol, tag
        for a, b in zip(factors, g.arrays(tag, "f")):
            assert rel_fro(a, b) <= (1e-7 if x.dtype == np.float64 else 5e-3)


def test_hals_vs_reference(golden):
    g = golden("round2")
    for tag in ("hals_a", "hals_b", "hals_c"):
        UtM, UtU, V0 = g[f"{tag}/UtM"], g[f"{tag}/UtU"], g[f"{tag}/V0"]
        tol = 1e-10 if UtM.dtype == np.float64 else 1e-4
        assert rel_fro(O.hals_nnls(UtM, UtU, V0, n_iter_max=100), g[f"{tag}/V"]) <= tol
        v = O.hals_nnls(UtM, UtU, V0, n_iter_max=20, sparsity_coefficient=0.05, ridge_coefficient=0.1, epsilon=1e-6)
        assert rel_fro(v, g[f"{tag}/V_sparse"]) <= tol
    for tag in ("nnhals64", "nnhals32"):
        x = g[f"{tag}/x"]
        rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
        init = (np.ones(rank, dtype=x.dtype), g.arrays(tag, "init"))
        (_, factors), errs = O.non_negative_parafac_hals(x, init, n_iter_max=iters)
        ref = g[f"{tag}/errors"]
        assert np.max(np.abs(np.array(errs, dtype=np.float64) - ref) / ref) <= (1e-9 if x.dtype == np.float64 else 1e-4)
        for a, b in zip(factors, g.arrays(tag, "f")):
            assert rel_fro(a, b) <= (1e-7 if x.dtype == np.float64 else 5e-3)
