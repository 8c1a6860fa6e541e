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

from . import _ops

BACKEND_NAME = "b200"
_OURS = {
    "unfolding_dot_khatri_rao": _ops.unfolding_dot_khatri_rao,
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


def use():
    """register() and select the backend (process-wide default + this thread)."""
    register()
    tl = import_tensorly()
    tl.tenalg.set_backend(BACKEND_NAME)
    return tl
