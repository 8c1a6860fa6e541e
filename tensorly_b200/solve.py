"""Sync-free small-system `solve` plug-in for the unmodified reference drivers.

The reference's ALS loop solves its R x R normal equations with `tl.solve(V^T, M^T)` per mode
(tensorly/decomposition/_cp.py:425-428); on the pytorch backend that is `torch.linalg.solve`, whose error check
reads a status word back to the host — one device synchronisation per mode, plus cuSOLVER's launch chain — which
caps the unmodified `parafac` at about half the speed of the kernels underneath it.  Like the SVD plug-in
(svd.py) this goes through the array backend's own sanctioned hook (the one tensorly/plugins.py:79-82 uses):

    tensorly_b200.use_fast_solve()      # BackendManager.register_backend_method("solve", fast_solve)
    tensorly_b200.use_default_solve()   # puts the previous function back

`fast_solve(A, B)` runs tlb200_cp_update's LU (partial pivoting, same arithmetic as the own driver) for a square
CUDA fp32/fp64 system of order <= 128 with a matrix right-hand side; anything else goes to the previous `solve`.
A singular matrix yields inf/nan instead of the LinAlgError torch raises (there is no host round trip to raise
it from).
"""
from __future__ import annotations

import torch

from . import _ops

_previous = None
MAX_ORDER = 128


def fast_solve(a, b, *args, **kwargs):
    prev = _previous if _previous is not None else torch.linalg.solve
    if (args or kwargs or not torch.is_tensor(a) or not torch.is_tensor(b) or not a.is_cuda or not b.is_cuda
            or a.dim() != 2 or b.dim() != 2 or a.shape[0] != a.shape[1] or a.shape[0] != b.shape[0]
            or a.shape[0] > MAX_ORDER or a.shape[0] < 1 or b.shape[1] < 1
            or a.dtype not in (torch.float32, torch.float64) or b.dtype != a.dtype
            or a.requires_grad or b.requires_grad):
        return prev(a, b, *args, **kwargs)
    # cp_update computes solve(V^T, M^T)^T: V = A^T and M = B^T are free views when the caller passes transposes
    # of contiguous matrices, as the reference loop does
    v = a.transpose(0, 1)
    if not v.is_contiguous():
        v = v.contiguous()
    m = b.transpose(0, 1)
    x_t = _ops.cp_update([None, v], 0, None, m)
    return x_t.transpose(0, 1)


def use_fast_solve():
    """Route the current TensorLy array backend's `solve` through fast_solve (idempotent)."""
    global _previous
    from .backend import import_tensorly
    tl = import_tensorly()
    current = getattr(tl.backend.BackendManager.current_backend(), "solve")
    if current is fast_solve:
        return
    _previous = current
    tl.backend.BackendManager.register_backend_method("solve", fast_solve)


def use_default_solve():
    """Undo use_fast_solve()."""
    global _previous
    if _previous is None:
        return
    from .backend import import_tensorly
    tl = import_tensorly()
    tl.backend.BackendManager.register_backend_method("solve", _previous)
    _previous = None
