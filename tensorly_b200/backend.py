"""Registration of the "b200" tenalg backend with an unmodified TensorLy.

The plug-in point is tensorly.tenalg's backend switch (tensorly/tenalg/__init__.py:15-87,
tensorly/tenalg/base_tenalg.py:4-33): a subclass of TenalgBackend created with
`backend_name="b200"` registers itself, its 11 dispatched functions are attached with
`register_method`, and the name is appended to `available_backend_names`.  After
`tensorly_b200.use()` the reference's own parafac / non_negative_parafac / tucker /
partial_tucker run on the CUDA kernels without a single line of TensorLy changed:

    import tensorly as tl, tensorly_b200
    tl.set_backend("pytorch")
    tensorly_b200.use()            # == register() + tl.tenalg.set_backend("b200")
    cp = tl.decomposition.parafac(x_cuda, rank=32)

The five hot-path functions are ours; the six functions outside the hot path are
delegated to the reference's `core` implementations (the dispatcher does getattr for all
11 names, tenalg/__init__.py:40-53).
"""
from __future__ import annotations

import os
import sys
import threading

from . import _ops

BACKEND_NAME = "b200"

# ---- optional dimension-tree reuse behind the stateless tenalg API --------------------------------------------
# The reference's ALS loops call unfolding_dot_khatri_rao(tensor, (weights, factors), mode) for mode = 0, 1, ..,
# N-1 with the SAME tensor object and — until the last mode — the same last-factor object
# (tensorly/decomposition/_cp.py:407-428).  With the cache on, the mode-0 call forms
# T = tensor x_{N-1} factors[-1]^T once and the calls for 0 <= mode < N-1 are answered from T
# (tlb200_mttkrp_from_ttm) for as long as `tensor` and `factors[-1]` are the very same objects at the same
# torch `_version` (so an in-place edit or a new tensor is a miss, never a stale hit; the cache holds references,
# so their storage cannot be recycled under it).  Off by default: tensorly_b200.use(dimension_tree=True) or
# TLB200_BACKEND_DIMTREE=1 turn it on.
_dimtree = {"on": os.environ.get("TLB200_BACKEND_DIMTREE", "0") == "1"}
_cache = threading.local()


def set_dimension_tree(flag: bool) -> None:
    """Enable/disable the per-sweep reuse of the last-mode contraction in `unfolding_dot_khatri_rao`."""
    _dimtree["on"] = bool(flag)
    _cache.entry = None


# ---- automatic range hint behind the stateless tenalg API ------------------------------------------------------
# The fp16-split engine (rank 33..64, include/tlb200.h: tlb200_hint_tensor_absmax) needs max |tensor|.  An ALS loop
# passes the same tensor object every call, so the SECOND call that sees the same object at the same torch
# `_version` pays one pass for max |x| and registers it; any other tensor, or an in-place edit, drops the hint
# before the kernel runs (a one-off call never pays).  TLB200_BACKEND_AUTOHINT=0 turns this off.
_autohint = {"on": os.environ.get("TLB200_BACKEND_AUTOHINT", "1") != "0"}


def _track_range_hint(tensor, rank):
    e = getattr(_cache, "hint", None)
    if e is not None and e[0]() is tensor and e[1] == tensor._version:
        if e[3] is None:
            e[2] += 1
            if e[2] >= 2:
                e[3] = _ops.RangeHint(tensor, hold=False)    # withdrawn when the tensor object dies
        return
    if e is not None and e[3] is not None:
        e[3].close()
    _cache.hint = None
    import torch
    if torch.is_tensor(tensor) and _ops.RangeHint.applies(tensor, rank):
        import weakref
        _cache.hint = [weakref.ref(tensor), tensor._version, 1, None]


def _unfolding_dot_khatri_rao(tensor, cp_tensor, mode):
    if _autohint["on"]:
        try:
            rank = cp_tensor[1][0].shape[1]
        except Exception:
            rank = 0
        _track_range_hint(tensor, rank)
    if not _dimtree["on"]:
        return _ops.unfolding_dot_khatri_rao(tensor, cp_tensor, mode)
    import torch
    weights, factors = cp_tensor
    factors = list(factors)
    ndim = tensor.dim() if torch.is_tensor(tensor) else 0
    if ndim >= 3 and len(factors) == ndim and -ndim <= mode < ndim and torch.is_tensor(factors[-1]):
        mode %= ndim
        last = factors[-1]
        if mode < ndim - 1:
            e = getattr(_cache, "entry", None)
            hit = (e is not None and e[0] is tensor and e[1] == tensor._version and e[2] is last and e[3] == last._version)
            if not hit and mode == 0:        # a sweep starts: one pass over the tensor serves modes 0 .. N-2
                t = _ops.mode_dot(tensor, last, ndim - 1, transpose=True)
                e = (tensor, tensor._version, last, last._version, t)
                _cache.entry = e
                hit = True
            if hit:
                return _ops.mttkrp_from_ttm(e[4], (weights, factors), mode)
        else:
            _cache.entry = None              # the last factor is about to change: free T
    return _ops.unfolding_dot_khatri_rao(tensor, cp_tensor, mode)


_OURS = {
    "unfolding_dot_khatri_rao": _unfolding_dot_khatri_rao,
    "khatri_rao": _ops.khatri_rao,
    "mode_dot": _ops.mode_dot,
    "multi_mode_dot": _ops.multi_mode_dot,
}
_DELEGATED = ("kronecker", "inner", "outer", "batched_outer", "higher_order_moment", "tensordot")

_backend_cls = None


def import_tensorly():
    """Import TensorLy: an installed package first, then the unmodified reference install
    the harness keeps under baseline/_ref (pip --target), then /root/reference."""
    try:
        import tensorly  # noqa: F401
        return sys.modules["tensorly"]
    except ImportError:
        pass
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "tensorly")):
            sys.path.insert(0, cand)
            try:
                import tensorly  # noqa: F401
                return sys.modules["tensorly"]
            except ImportError:
                sys.path.remove(cand)
    raise ImportError("tensorly is not importable; install tensorly to use the b200 tenalg backend "
                      "(the tensorly_b200 functions themselves do not need it)")


def register():
    """Create and register the b200 tenalg backend class (idempotent). Returns the class."""
    global _backend_cls
    if _backend_cls is not None:
        return _backend_cls
    tl = import_tensorly()
    from tensorly.tenalg import core_tenalg
    from tensorly.tenalg.base_tenalg import TenalgBackend

    class B200TenalgBackend(TenalgBackend, backend_name=BACKEND_NAME):
        """Hand-written sm_100a kernels for the dense-decomposition hot path."""

    for name, fn in _OURS.items():
        B200TenalgBackend.register_method(name, fn)
    for name in _DELEGATED:
        B200TenalgBackend.register_method(name, getattr(core_tenalg, name))
    B200TenalgBackend.register_method("_tt_matrix_to_tensor", core_tenalg.tt_matrix_to_tensor)
    if BACKEND_NAME not in tl.tenalg.available_backend_names:
        tl.tenalg.available_backend_names.append(BACKEND_NAME)
    _backend_cls = B200TenalgBackend
    return B200TenalgBackend


def use(dimension_tree=None):
    """register() and select the backend (process-wide default + this thread).  `dimension_tree=True/False`
    also switches the per-sweep reuse described above (None leaves it as it is)."""
    register()
    if dimension_tree is not None:
        set_dimension_tree(dimension_tree)
    tl = import_tensorly()
    tl.tenalg.set_backend(BACKEND_NAME)
    return tl
