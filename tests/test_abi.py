"""CPU-only checks of the C ABI boundary: the shared library loads, exports every symbol
that include/tlb200.h declares, validates arguments before touching the device, and its
host-side MTTKRP planning is sane.  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

import tensorly_b200 as tb
from tensorly_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tlb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tlb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tlb200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_introspection():
    lib = _lib.load()
    assert lib.tlb200_version() >= 100
    assert lib.tlb200_build_arch() == b"sm_100a"
    assert b"invalid" in lib.tlb200_status_string(_lib.TLB200_EINVAL)


def test_plan_three_way():
    # mode 0: X[j, (a, b)], K-contiguous; P = F1, Q = F2
    p = tb.mttkrp_plan((1024, 1024, 1024), 0, 32)
    assert (p.A, p.J, p.B) == (1024, 1024, 1024) and (p.sa, p.sj, p.sb) == (1024, 1024 * 1024, 1)
    assert (p.p_first, p.p_count, p.q_first, p.q_count) == (1, 1, 2, 1)
    # middle mode: X[a, j, b]
    p = tb.mttkrp_plan((10, 20, 30), 1, 4, torch.float64)
    assert (p.A, p.J, p.B) == (10, 20, 30) and (p.sa, p.sj, p.sb) == (600, 30, 1)
    assert (p.p_first, p.p_count, p.q_first, p.q_count) == (0, 1, 2, 1)
    # last mode: X[(a, b), j], j contiguous
    p = tb.mttkrp_plan((10, 20, 30), 2, 4)
    assert (p.A, p.J, p.B) == (10, 30, 20) and (p.sa, p.sj, p.sb) == (600, 1, 30)
    assert p.rank_padded >= 4 and p.splits >= 1


def test_plan_never_materialises_the_full_khatri_rao():
    # config 4: 256^4, rank 64 — the two tables together must be tiny next to the tensor
    shape = (256, 256, 256, 256)
    for mode in range(4):
        p = tb.mttkrp_plan(shape, mode, 64)
        assert p.A * p.B == 256 ** 3
        assert p.A + p.B <= 256 + 256 ** 2
        assert p.J == 256
    # 2-way: no P table at all
    p = tb.mttkrp_plan((130, 70), 0, 9)
    assert p.p_count == 0 and p.A == 1 and p.B == 70
    p = tb.mttkrp_plan((130, 70), 1, 9)
    assert p.p_count == 0 and p.A == 1 and p.B == 130 and p.sj == 1 and p.sb == 70


def test_plan_rejects_bad_arguments():
    with pytest.raises(ValueError):
        tb.mttkrp_plan((10,), 0, 4)           # MTTKRP needs >= 2 modes
    with pytest.raises(ValueError):
        tb.mttkrp_plan((10, 10), 2, 4)        # mode out of range
    with pytest.raises(ValueError):
        tb.mttkrp_plan((10, 10), 0, 0)        # rank < 1
    with pytest.raises(ValueError):
        tb.mttkrp_plan((10, 0, 10), 0, 3)     # empty extent


def test_c_abi_validates_before_launching():
    lib = _lib.load()
    shape = _lib.i64_array((4, 5, 6))
    # null pointers / bad dtype / bad mode are refused with EINVAL without any device work
    assert lib.tlb200_unfold(None, shape, 3, 1, _lib.F32, None, None) == _lib.TLB200_EINVAL
    assert lib.tlb200_unfold(None, shape, 3, 7, _lib.F32, None, None) == _lib.TLB200_EINVAL
    assert lib.tlb200_unfold(None, shape, 3, 1, 9, None, None) == _lib.TLB200_EINVAL
    assert lib.tlb200_mttkrp_workspace_bytes(shape, 3, 0, 4, _lib.F32, _lib.PATH_AUTO) > 0
    assert lib.tlb200_mttkrp_workspace_bytes(shape, 3, 5, 4, _lib.F32, _lib.PATH_AUTO) == 0
    assert lib.tlb200_gram_workspace_bytes(100, 8, _lib.F64) >= 8 * 8 * 8
    st = lib.tlb200_mode_dot(None, shape, 3, 0, None, 3, 1, 1, _lib.F32, None, None, 0, _lib.PATH_AUTO, None)
    assert st == _lib.TLB200_EINVAL
    # dimension-tree MTTKRP: lead_shape = the first N-1 extents; only modes < N-1 exist; >= 3-way problems only
    lead = _lib.i64_array((4, 5))
    assert lib.tlb200_mttkrp_from_ttm_workspace_bytes(lead, 2, 0, 8, _lib.F32) > 0
    assert lib.tlb200_mttkrp_from_ttm_workspace_bytes(lead, 2, 2, 8, _lib.F32) == 0
    assert lib.tlb200_mttkrp_from_ttm_workspace_bytes(lead, 1, 0, 8, _lib.F32) == 0
    st = lib.tlb200_mttkrp_from_ttm(None, lead, 2, 0, None, None, None, 8, None, _lib.F32, None, 8, None, 0, None)
    assert st == _lib.TLB200_EINVAL
    # fused update + Gram: 256 bytes for the ticket counter + one R x R partial per 64 rows
    assert lib.tlb200_cp_update_gram_workspace_bytes(1024, 32, _lib.F32) >= 256 + 16 * 32 * 32 * 4
    assert lib.tlb200_cp_update_gram_workspace_bytes(-1, 32, _lib.F32) == 0
    st = lib.tlb200_cp_update_gram(None, 3, 0, 8, None, 0.0, None, 8, 10, _lib.F32, None, 8, None, None, 0, None)
    assert st == _lib.TLB200_EINVAL


def test_range_hint_entry_points_validate_and_the_shape_only_plan_stays_tf32():
    """tlb200_hint_tensor_absmax is host-only bookkeeping (a pointer table): it validates its arguments, accepts a
    withdrawal of something never registered, and the shape-only plan never reports the fp16-split engine."""
    import ctypes
    import tensorly_b200 as tb
    lib = _lib.load()
    assert lib.tlb200_hint_tensor_absmax(None, None) == _lib.TLB200_EINVAL
    assert lib.tlb200_hint_tensor_absmax(ctypes.c_void_p(0x1000), None) == 0          # nothing registered: a no-op
    assert lib.tlb200_tensor_absmax(None, 10, _lib.F32, None, None) == _lib.TLB200_EINVAL
    assert lib.tlb200_tensor_absmax(ctypes.c_void_p(0x1000), 10, _lib.F64, ctypes.c_void_p(0x2000), None) != 0   # fp32 only
    assert tb.mttkrp_plan((256, 192, 320), 1, 64).f16 == 0
    # the fp16 engine blocks the contraction differently: the workspace query covers both plans
    shape = _lib.i64_array((256, 192, 320))
    assert lib.tlb200_mttkrp_workspace_bytes(shape, 3, 1, 64, _lib.F32, _lib.PATH_AUTO) > 0


def test_plan_reports_column_block_passes_field():
    """rank_passes is part of the plan struct: 1 on the SIMT path (all a GPU-less box can plan), ceil(rank / 64)
    column blocks when the tcgen05 engine is available."""
    import tensorly_b200 as tb
    pl = tb.mttkrp_plan((64, 48, 80), 1, 100)
    assert pl.rank_passes == (2 if pl.path == _lib.PATH_TCGEN05 else 1)
    assert pl.A * pl.J * pl.B == 64 * 48 * 80
    assert tb.mttkrp_plan((64, 48, 80), 1, 100, path="simt").rank_passes == 1


def test_host_mirror_refuses_cpu_tensors_loudly():
    """No CPU fallback: the product path raises instead of computing on the host."""
    x = torch.rand(4, 5, 6)
    fs = [torch.rand(s, 3) for s in x.shape]
    with pytest.raises(TypeError, match="CUDA"):
        tb.unfolding_dot_khatri_rao(x, (None, fs), 0)
    with pytest.raises(TypeError, match="CUDA"):
        tb.mode_dot(x, torch.rand(2, 4), 0)
    with pytest.raises(TypeError, match="CUDA"):
        tb.khatri_rao(fs)
    with pytest.raises(TypeError, match="CUDA"):
        tb.unfold(x, 1)
    with pytest.raises(TypeError):
        tb.unfold([[1.0, 2.0]], 0)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtlb200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_shard_bounds_partition():
    for extent, world in ((2048, 8), (10, 4), (7, 8), (1024, 3)):
        spans = [tb.shard_bounds(extent, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == extent
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and b - a >= d - c >= 0


def test_tensorly_backend_registration():
    tl = pytest.importorskip("tensorly") if False else None
    try:
        tl = tb.import_tensorly()
    except ImportError:
        pytest.skip("tensorly not importable here")
    cls = tb.register()
    assert cls.backend_name == "b200" and "b200" in tl.tenalg.available_backend_names
    prev = tl.tenalg.get_backend()
    try:
        tl.tenalg.set_backend("b200")
        assert tl.tenalg.get_backend() == "b200"
        # all 11 dispatched names resolve on the new backend
        for name in ("mode_dot", "multi_mode_dot", "kronecker", "khatri_rao", "inner", "outer", "batched_outer",
                     "higher_order_moment", "_tt_matrix_to_tensor", "unfolding_dot_khatri_rao", "tensordot"):
            assert callable(getattr(cls, name))
        import inspect
        sig = inspect.signature(tl.tenalg.unfolding_dot_khatri_rao)
        assert list(sig.parameters) == ["tensor", "cp_tensor", "mode"]
    finally:
        tl.tenalg.set_backend(prev)
