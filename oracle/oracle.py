"""CPU oracle for the TensorLy dense-decomposition hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the reference
algorithms (tensorly/tensorly v0.9.0, numpy backend + `core` tenalg).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl
reference` legs may import it, and only as the checker or the timed CPU baseline.
Nothing under `tensorly_b200/` imports it: the product path is CUDA-only and fails
loudly when the extension is missing.

Parity status: PINNED.  `oracle/gen_golden.py` ran the real reference (imported from
/root/reference in the build container) and committed its outputs to
`tests/golden/*.npz`; `tests/test_oracle.py` checks this restatement against those
fixtures (bit-exact for unfold/fold/khatri_rao, and for every function when the same
numpy build is used) and against the golden vectors of the reference's own tests.

Each function cites the reference lines it follows (paths relative to the checkout).
"""
from __future__ import annotations

import math
import numpy as np


# --------------------------------------------------------------------------- #
# base.py
# --------------------------------------------------------------------------- #
def unfold(tensor: np.ndarray, mode: int) -> np.ndarray:
    """tensorly/base.py:39-53 — reshape(moveaxis(tensor, mode, 0), (shape[mode], -1))."""
    return np.reshape(np.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))


def fold(unfolded: np.ndarray, mode: int, shape) -> np.ndarray:
    """tensorly/base.py:56-79 — inverse of unfold."""
    full_shape = list(shape)
    mode_dim = full_shape.pop(mode)
    full_shape.insert(0, mode_dim)
    return np.moveaxis(np.reshape(unfolded, full_shape), 0, mode)


# --------------------------------------------------------------------------- #
# tenalg/core_tenalg/_khatri_rao.py
# --------------------------------------------------------------------------- #
def khatri_rao(matrices, weights=None, skip_matrix=None, mask=None) -> np.ndarray:
    """tensorly/tenalg/core_tenalg/_khatri_rao.py:64-109.

    Left fold of broadcast multiplies: res <- reshape(res[:,None,:] * e[None,:,:]),
    with `weights` folded into the first matrix (:96-99) and the result multiplied by
    the mask column or by 1 (:107-109).  A single remaining matrix is returned as is,
    ignoring weights (:68-69).
    """
    if skip_matrix is not None:
        matrices = [m for i, m in enumerate(matrices) if i != skip_matrix]
    if len(matrices) == 1:
        return matrices[0]
    if np.ndim(matrices[0]) == 2:
        n_columns = matrices[0].shape[1]
    else:
        n_columns = 1
        matrices = [np.reshape(m, (-1, 1)) for m in matrices]
    for i, m in enumerate(matrices):
        if np.ndim(m) != 2:
            raise ValueError(f"All the matrices must have exactly 2 dimensions! Matrix {i} has {np.ndim(m)}.")
        if m.shape[1] != n_columns:
            raise ValueError(f"All matrices must have same number of columns! Matrix {i} has {m.shape[1]} != {n_columns}.")
    res = None
    for i, e in enumerate(matrices[1:]):
        if not i:
            res = matrices[0] if weights is None else matrices[0] * np.reshape(weights, (1, -1))
        s1, s2 = res.shape
        s3, s4 = e.shape
        res = np.reshape(np.reshape(res, (s1, 1, s2)) * np.reshape(e, (1, s3, s4)), (-1, n_columns))
    m = np.reshape(mask, (-1, 1)) if mask is not None else 1
    return res * m


# --------------------------------------------------------------------------- #
# tenalg/core_tenalg/mttkrp.py
# --------------------------------------------------------------------------- #
def unfolding_dot_khatri_rao(tensor, cp_tensor, mode) -> np.ndarray:
    """tensorly/tenalg/core_tenalg/mttkrp.py:47-49 — dot(unfold(X, mode), conj(KR))."""
    weights, factors = cp_tensor
    kr = khatri_rao(factors, weights=weights, skip_matrix=mode)
    return np.dot(unfold(tensor, mode), np.conj(kr))


def mttkrp_float64_truth(tensor, cp_tensor, mode) -> np.ndarray:
    """Same contraction evaluated in float64 (for reporting error against 'truth')."""
    weights, factors = cp_tensor
    t64 = np.asarray(tensor, dtype=np.float64)
    f64 = [np.asarray(f, dtype=np.float64) for f in factors]
    w64 = None if weights is None else np.asarray(weights, dtype=np.float64)
    return unfolding_dot_khatri_rao(t64, (w64, f64), mode)


# --------------------------------------------------------------------------- #
# tenalg/core_tenalg/n_mode_product.py
# --------------------------------------------------------------------------- #
def mode_dot(tensor, matrix_or_vector, mode, transpose=False) -> np.ndarray:
    """tensorly/tenalg/core_tenalg/n_mode_product.py:5-76."""
    fold_mode = mode
    new_shape = list(tensor.shape)
    if np.ndim(matrix_or_vector) == 2:
        dim = 0 if transpose else 1
        if matrix_or_vector.shape[dim] != tensor.shape[mode]:
            raise ValueError(
                f"shapes {tensor.shape} and {matrix_or_vector.shape} not aligned in mode-{mode} multiplication"
            )
        if transpose:
            matrix_or_vector = np.conj(np.transpose(matrix_or_vector))
        new_shape[mode] = matrix_or_vector.shape[0]
        vec = False
    elif np.ndim(matrix_or_vector) == 1:
        if matrix_or_vector.shape[0] != tensor.shape[mode]:
            raise ValueError(
                f"shapes {tensor.shape} and {matrix_or_vector.shape} not aligned for mode-{mode} multiplication"
            )
        if len(new_shape) > 1:
            new_shape.pop(mode)
        else:
            new_shape = ()
        vec = True
    else:
        raise ValueError("Can only take n_mode_product with a vector or a matrix.")
    res = np.dot(matrix_or_vector, unfold(tensor, mode))
    if vec:
        return np.reshape(res, new_shape)  # vec_to_tensor, base.py:21-36
    return fold(res, fold_mode, new_shape)


def multi_mode_dot(tensor, matrix_or_vec_list, modes=None, skip=None, transpose=False) -> np.ndarray:
    """tensorly/tenalg/core_tenalg/n_mode_product.py:113-135.

    Pairs are sorted by mode (:122); `skip` indexes the *sorted list* (:124); a vector
    operand decrements the later mode numbers (:132-133).
    """
    if modes is None:
        modes = range(len(matrix_or_vec_list))
    decrement = 0
    res = tensor
    factors_modes = sorted(zip(matrix_or_vec_list, modes), key=lambda x: x[1])
    for i, (m, mode) in enumerate(factors_modes):
        if (skip is not None) and (i == skip):
            continue
        if transpose:
            res = mode_dot(res, np.conj(np.transpose(m)), mode - decrement)
        else:
            res = mode_dot(res, m, mode - decrement)
        if np.ndim(m) == 1:
            decrement += 1
    return res


# --------------------------------------------------------------------------- #
# cp_tensor.py helpers
# --------------------------------------------------------------------------- #
def cp_to_tensor(cp_tensor) -> np.ndarray:
    """tensorly/cp_tensor.py:433-485 (mask=None branch): fold(F0*w . KR(skip 0)^T)."""
    weights, factors = cp_tensor
    shape = tuple(f.shape[0] for f in factors)
    if weights is None:
        weights = np.ones(factors[0].shape[1], dtype=factors[0].dtype)
    if len(shape) == 1:
        return np.sum(weights * factors[0], axis=1)
    full = np.dot(factors[0] * weights, np.transpose(khatri_rao(factors, skip_matrix=0)))
    return fold(full, 0, shape)


def cp_norm(cp_tensor) -> float:
    """tensorly/cp_tensor.py:614-644 — sqrt(sum((w w^T) o prod_n F_n^T conj(F_n)))."""
    weights, factors = cp_tensor
    rank = factors[0].shape[1]
    norm = np.ones((rank, rank), dtype=factors[0].dtype)
    for f in factors:
        norm = norm * np.dot(np.transpose(f), np.conj(f))
    if weights is not None:
        norm = norm * (np.reshape(weights, (-1, 1)) * np.reshape(weights, (1, -1)))
    return np.sqrt(np.sum(norm))


def tensor_norm(tensor) -> float:
    """backend/core.py:736-737 (order=2): sqrt(sum(abs(t)**2))."""
    return np.sqrt(np.sum(np.abs(tensor) ** 2))


# --------------------------------------------------------------------------- #
# decomposition/_cp.py — the ALS loop for fixed initial factors
# --------------------------------------------------------------------------- #
def parafac(tensor, init, n_iter_max=10, l2_reg=0.0, return_errors=True):
    """tensorly/decomposition/_cp.py:394-440 with init=(weights, factors), tol=0,
    no mask / sparsity / linesearch / orthogonalise / normalize_factors.

    Per mode: Gram-Hadamard (:411-422), MTTKRP (:423), solve (:425-428); per sweep the
    fast error via the last mode's MTTKRP (:217-225).
    """
    weights, factors = init
    factors = [np.array(f, copy=True) for f in factors]
    rank = factors[0].shape[1]
    dtype = tensor.dtype
    weights = np.ones(rank, dtype=dtype) if weights is None else np.array(weights, dtype=dtype)
    norm_tensor = tensor_norm(tensor)
    Id = np.eye(rank, dtype=dtype) * l2_reg if l2_reg else 0
    rec_errors = []
    for _ in range(n_iter_max):
        mttkrp = None
        for mode in range(tensor.ndim):
            pinv = np.ones((rank, rank), dtype=dtype)
            for i, f in enumerate(factors):
                if i != mode:
                    pinv = pinv * np.dot(np.conj(np.transpose(f)), f)
            pinv = pinv + Id
            pinv = np.reshape(weights, (-1, 1)) * pinv * np.reshape(weights, (1, -1))
            mttkrp = unfolding_dot_khatri_rao(tensor, (weights, factors), mode)
            factors[mode] = np.transpose(np.linalg.solve(np.conj(np.transpose(pinv)), np.transpose(mttkrp)))
        if return_errors:
            factors_norm = cp_norm((weights, factors))
            iprod = np.sum(np.sum(mttkrp * np.conj(factors[-1]), axis=0))
            unnorm = np.sqrt(np.abs(norm_tensor ** 2 + factors_norm ** 2 - 2 * iprod))
            rec_errors.append(unnorm / norm_tensor)
    return (weights, factors), rec_errors


def cp_impute(tensor, mask, cp_tensor):
    """The mask branch of error_calc, tensorly/decomposition/_cp.py:195-207: returns the imputed tensor,
    its norm and the unnormalised error ||tensor_new - rec||."""
    low_rank = cp_to_tensor(cp_tensor)
    tensor = tensor * mask + low_rank * (1 - mask)
    return tensor, tensor_norm(tensor), tensor_norm(tensor - low_rank)


def parafac_masked(tensor, mask, init, n_iter_max=10):
    """tensorly/decomposition/_cp.py:394-478 with a mask (1 = observed), init=(weights, factors), tol=0,
    return_errors=True: the factor updates run on the current (imputed) tensor, then error_calc re-imputes the
    missing entries from the new factors, recomputes the tensor norm and reports ||tensor - rec|| / ||tensor||."""
    weights, factors = init
    factors = [np.array(f, copy=True) for f in factors]
    rank = factors[0].shape[1]
    dtype = tensor.dtype
    weights = np.ones(rank, dtype=dtype) if weights is None else np.array(weights, dtype=dtype)
    tensor = np.array(tensor, copy=True)
    rec_errors = []
    for _ in range(n_iter_max):
        for mode in range(tensor.ndim):
            pinv = np.ones((rank, rank), dtype=dtype)
            for i, f in enumerate(factors):
                if i != mode:
                    pinv = pinv * np.dot(np.conj(np.transpose(f)), f)
            pinv = np.reshape(weights, (-1, 1)) * pinv * np.reshape(weights, (1, -1))
            mttkrp = unfolding_dot_khatri_rao(tensor, (weights, factors), mode)
            factors[mode] = np.transpose(np.linalg.solve(np.conj(np.transpose(pinv)), np.transpose(mttkrp)))
        tensor, norm_tensor, unnorm = cp_impute(tensor, mask, (weights, factors))
        rec_errors.append(unnorm / norm_tensor)
    return (weights, factors), rec_errors, tensor


def hals_nnls(UtM, UtU, V, n_iter_max=500, tol=1e-8, sparsity_coefficient=None, ridge_coefficient=None, epsilon=0.0):
    """tensorly/solvers/nnls.py:139-173 for a given V (nonzero_rows=False, exact=False, no callback).  Note the
    reference's stopping statistic: `tl.norm(V - newV) ** 2` subtracts the new ROW from the whole matrix
    (broadcast), i.e. sum_{l, j} (V[l, j] - newV[j])^2 — restated as is."""
    V = np.array(V, copy=True)
    rank = UtM.shape[0]
    rec_error0 = None
    for iteration in range(n_iter_max):
        rec_error = 0
        for k in range(rank):
            if UtU[k, k]:
                num = UtM[k, :] - np.dot(UtU[k, :], V) + UtU[k, k] * V[k, :]
                den = UtU[k, k]
                if sparsity_coefficient is not None:
                    num = num - sparsity_coefficient
                if ridge_coefficient is not None:
                    den = den + 2 * ridge_coefficient
                newV = np.clip(num / den, epsilon, None)
                rec_error += np.sqrt(np.sum(np.abs(V - newV) ** 2)) ** 2
                V[k, :] = newV
        if iteration == 0:
            rec_error0 = rec_error
        if rec_error < tol * rec_error0:
            break
    return V


def non_negative_parafac_hals(tensor, init, n_iter_max=10, return_errors=True):
    """tensorly/decomposition/_nn_cp.py:307-379 (nn_modes='all', no sparsity, exact=False, fixed_modes=[],
    normalize_factors=False, tol truthy) for init=(weights, factors)."""
    weights, factors = init
    factors = [np.array(f, copy=True) for f in factors]
    rank = factors[0].shape[1]
    dtype = tensor.dtype
    weights = np.ones(rank, dtype=dtype) if weights is None else np.array(weights, dtype=dtype)
    norm_tensor = tensor_norm(tensor)
    rec_errors = []
    for _ in range(n_iter_max):
        mttkrp = None
        for mode in range(tensor.ndim):
            pinv = np.ones((rank, rank), dtype=dtype)
            for i, f in enumerate(factors):
                if i != mode:
                    pinv = pinv * np.dot(np.transpose(f), f)
            pinv = np.reshape(weights, (-1, 1)) * pinv * np.reshape(weights, (1, -1))
            mttkrp = unfolding_dot_khatri_rao(tensor, (weights, factors), mode)
            nn = hals_nnls(np.transpose(mttkrp), pinv, np.transpose(factors[mode]), n_iter_max=100)
            factors[mode] = np.transpose(nn)
        if return_errors:
            factors_norm = cp_norm((weights, factors))
            iprod = np.sum(np.sum(mttkrp * factors[-1], axis=0))
            rec_errors.append(np.sqrt(np.abs(norm_tensor ** 2 + factors_norm ** 2 - 2 * iprod)) / norm_tensor)
    return (weights, factors), rec_errors


def non_negative_parafac(tensor, init, n_iter_max=10, return_errors=True):
    """tensorly/decomposition/_nn_cp.py:107-152 (multiplicative updates, no mask,
    fixed_modes=[], normalize_factors=False) for init=(weights, factors)."""
    weights, factors = init
    factors = [np.array(f, copy=True) for f in factors]
    rank = factors[0].shape[1]
    dtype = tensor.dtype
    weights = np.ones(rank, dtype=dtype) if weights is None else np.array(weights, dtype=dtype)
    eps = np.finfo(dtype).eps  # tl.eps(tensor.dtype), _nn_cp.py:82
    norm_tensor = tensor_norm(tensor)
    rec_errors = []
    for _ in range(n_iter_max):
        mttkrp = None
        for mode in range(tensor.ndim):
            accum = 1
            for i, e in enumerate([i for i in range(len(factors)) if i != mode]):
                if i:
                    accum = accum * np.dot(np.transpose(factors[e]), factors[e])
                else:
                    accum = np.dot(np.transpose(factors[e]), factors[e])
            accum = np.reshape(weights, (-1, 1)) * accum * np.reshape(weights, (1, -1))
            mttkrp = unfolding_dot_khatri_rao(tensor, (weights, factors), mode)
            numerator = np.clip(mttkrp, eps, None)
            denominator = np.clip(np.dot(factors[mode], accum), eps, None)
            factors[mode] = factors[mode] * numerator / denominator
        if return_errors:
            factors_norm = cp_norm((weights, factors))
            iprod = np.sum(np.sum(mttkrp * np.conj(factors[-1]), axis=0))
            rec_errors.append(np.sqrt(np.abs(norm_tensor ** 2 + factors_norm ** 2 - 2 * iprod)) / norm_tensor)
    return (weights, factors), rec_errors


# --------------------------------------------------------------------------- #
# decomposition/_tucker.py — HOOI for fixed initial factors
# --------------------------------------------------------------------------- #
def svd_flip_u(U: np.ndarray) -> np.ndarray:
    """tenalg/svd.py:13-65 restricted to U (u_based_decision=True): make the largest-
    magnitude entry of every column positive."""
    U = np.array(U, copy=True)
    idx = np.argmax(np.abs(U), axis=0)
    signs = np.sign(U[idx, np.arange(U.shape[1])])
    return U * signs


def truncated_svd_u(matrix: np.ndarray, n_eigenvecs: int) -> np.ndarray:
    """tenalg/svd.py:211-235 + svd_interface flip (:366-447): leading left singular
    vectors of `matrix`."""
    U, _, _ = np.linalg.svd(matrix, full_matrices=False)
    return svd_flip_u(U[:, :n_eigenvecs])


def tucker_hooi(tensor, rank, init_factors, n_iter_max=5):
    """tensorly/decomposition/_tucker.py:187-207 for explicit initial factors:
    per mode, project on all other factors (multi_mode_dot(..., skip=index,
    transpose=True), :194-196), take the leading left singular vectors of the mode
    unfolding (:197-201); then core = multi_mode_dot(X, factors, transpose=True) (:204)
    and rec_error = sqrt(|‖X‖² − ‖core‖²|)/‖X‖ (:207)."""
    factors = [np.array(f, copy=True) for f in init_factors]
    modes = list(range(tensor.ndim))
    norm_tensor = tensor_norm(tensor)
    rec_errors = []
    core = None
    for _ in range(n_iter_max):
        for index, mode in enumerate(modes):
            core_approx = multi_mode_dot(tensor, factors, modes=modes, skip=index, transpose=True)
            factors[index] = truncated_svd_u(unfold(core_approx, mode), rank[index])
        core = multi_mode_dot(tensor, factors, modes=modes, transpose=True)
        rec_errors.append(math.sqrt(abs(norm_tensor ** 2 - tensor_norm(core) ** 2)) / norm_tensor)
    return (core, factors), rec_errors


# --------------------------------------------------------------------------- #
# random/base.py — input generators (bit-identical to the reference's)
# --------------------------------------------------------------------------- #
def random_tensor(shape, seed, dtype=np.float64) -> np.ndarray:
    """tensorly/random/base.py:12-15 — RandomState(seed).random_sample(shape)."""
    return np.random.RandomState(seed).random_sample(shape).astype(dtype, copy=False)


def random_cp_factors(shape, rank, seed, dtype=np.float64):
    """tensorly/random/base.py:103-105,113-114 — one random_sample((I_n, R)) per mode in
    mode order from a single RandomState; weights = ones."""
    rns = np.random.RandomState(seed)
    factors = [rns.random_sample((s, rank)).astype(dtype, copy=False) for s in shape]
    weights = np.ones(rank, dtype=dtype)
    return weights, factors
