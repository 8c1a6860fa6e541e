"""Host logic of the own HOOI driver (tensorly_b200.tucker / partial_tucker) with oracle-backed ops on the CPU:
which projections are formed, exact Rayleigh-Ritz path for small modes, oversampled subspace iteration for wide
ones, error assembly — against the reference's trajectories (tests/golden) and the oracle's exact-SVD HOOI."""
import numpy as np
import torch

import tensorly_b200 as tb
from oracle import oracle as O
from oracle_ops import OracleOps


def test_hooi_small_modes_follow_reference_exactly(golden):
    g, g2 = golden("als"), golden("round2")
    x = g["tucker/x"]
    ranks = [int(r) for r in g["tucker/ranks"]]
    (core, fs), errs = tb.tucker(torch.from_numpy(x.copy()), ranks, n_iter_max=5, init="random", random_state=1, tol=0,
                                 return_errors=True, ops=OracleOps)
    ref = g["tucker/errors"]
    assert np.max(np.abs(np.array(errs) - ref) / ref) <= 1e-10
    assert abs(float(torch.linalg.norm(core)) - float(g["tucker/core_norm"])) <= 1e-9 * float(g["tucker/core_norm"])
    for tag, iters in (("tucker_svd", 6), ("tucker_svd32", 8)):
        x = g2[f"{tag}/x"]
        ranks = [int(r) for r in g2[f"{tag}/ranks"]]
        (_, _), errs = tb.tucker(torch.from_numpy(x.copy()), ranks, n_iter_max=iters, init="svd", tol=0, return_errors=True,
                                 ops=OracleOps)
        ref = g2[f"{tag}/errors"]
        assert np.max(np.abs(np.array(errs) - ref) / ref) <= (1e-10 if x.dtype == np.float64 else 1e-4)


def test_hooi_wide_modes_within_gate_of_exact_svd_hooi():
    shape, ranks = (100, 90, 80), [10, 12, 8]
    x = np.random.RandomState(0).random_sample(shape)
    rs = np.random.RandomState(1)
    rs.random_sample(ranks)
    init = [rs.random_sample((s, r)) for s, r in zip(shape, ranks)]
    (_, _), ref = O.tucker_hooi(x, ranks, init, n_iter_max=4)
    (core, fs), errs = tb.tucker(torch.from_numpy(x.copy()), ranks, n_iter_max=4, init="random", random_state=1, tol=0,
                                 return_errors=True, ops=OracleOps)
    assert np.max(np.abs(np.array(errs) - np.array(ref)) / np.array(ref)) <= 1e-4
    for f in fs:
        assert np.linalg.norm(f.numpy().T @ f.numpy() - np.eye(f.shape[1])) <= 1e-8
    # early stop on tol, partial modes, int rank
    (_, _), errs2 = tb.partial_tucker(torch.from_numpy(x.copy()), [10, 8], modes=[0, 2], n_iter_max=50, init="svd", tol=1e-3,
                                      ops=OracleOps)
    assert 3 <= len(errs2) < 50
