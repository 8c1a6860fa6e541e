"""Truncated-SVD plug-in for HOOI / initialize_* (SURVEY.md section 8(f) n1).

`partial_tucker` calls `svd_interface(unfold(core, mode), n_eigenvecs=rank)` once per mode and sweep and ignores
its `svd=` argument inside the loop (tensorly/decomposition/_tucker.py:197-201), so the only non-invasive way to
change the SVD is the array backend's own `svd` entry — the same hook tensorly/plugins.py:79-82 uses for einsum:

    tensorly_b200.use_gram_svd()        # tl.backend.BackendManager.register_backend_method("svd", gram_svd)
    tensorly_b200.use_default_svd()     # puts the previous function back

For a short-fat (m x n, m <= n) matrix the left singular vectors are the eigenvectors of the m x m Gram matrix
A A^T: one GEMM over A, one `eigh` of an m x m matrix (512 x 512 at C3) and one GEMM for V replace a full
LAPACK/cuSOLVER SVD of A.  Same algebra as the reference's own `symeig_svd` (tensorly/tenalg/svd.py:238-285).
Squaring halves the attainable relative accuracy of the SMALL singular values (below sqrt(eps) * sigma_max they
are noise); HOOI only keeps the leading `rank` vectors, for which the parity tests hold the reference's
trajectory to 1e-4.  The GEMMs and eigh are library calls (cuBLAS / cuSOLVER through torch) — plain dense linear
algebra outside the hand-written hot path.
"""
from __future__ import annotations

import torch

_previous = None


def gram_svd(matrix, full_matrices=True, **kwargs):
    """Drop-in for the backend's `svd(matrix, full_matrices=...)`: returns (U, S, Vh) with min(m, n) columns/rows.
    Falls back to the previous `svd` when full matrices are requested or the input is not a real 2-D tensor."""
    prev = _previous if _previous is not None else torch.linalg.svd
    if (full_matrices or not torch.is_tensor(matrix) or matrix.dim() != 2 or not matrix.is_floating_point()
            or matrix.numel() == 0):
        return prev(matrix, full_matrices=full_matrices, **kwargs)
    m, n = matrix.shape
    a = matrix if m <= n else matrix.transpose(0, 1)       # a is short-fat
    gram = a @ a.transpose(0, 1)
    lam, vec = torch.linalg.eigh(gram)                      # ascending
    lam = torch.flip(lam, dims=(0,))
    u = torch.flip(vec, dims=(1,))
    tiny = torch.finfo(matrix.dtype).eps
    s = torch.sqrt(torch.clamp(lam, min=tiny * tiny))
    vh = (u.transpose(0, 1) @ a) / s.unsqueeze(1)
    if m <= n:
        return u, s, vh
    return vh.transpose(0, 1), s, u.transpose(0, 1)


def use_gram_svd():
    """Route the current TensorLy array backend's `svd` through gram_svd (idempotent)."""
    global _previous
    from .backend import import_tensorly
    tl = import_tensorly()
    current = getattr(tl.backend.BackendManager.current_backend(), "svd")
    if current is gram_svd:
        return
    _previous = current
    tl.backend.BackendManager.register_backend_method("svd", gram_svd)


def use_default_svd():
    """Undo use_gram_svd()."""
    global _previous
    if _previous is None:
        return
    from .backend import import_tensorly
    tl = import_tensorly()
    tl.backend.BackendManager.register_backend_method("svd", _previous)
    _previous = None
