"""Peer-memory all-reduce for the sharded drivers (one process per GPU on one NVLink / NVSwitch node).

`P2PComm` owns one symmetric device buffer per rank, exchanges the CUDA IPC handles once through the
torch.distributed process group (plumbing), and afterwards every all-reduce of a small tensor is ONE launch of
tlb200_allreduce_oneshot (csrc/comm.cu): push to every peer over NVLink, publish a flag, wait, sum in rank order.
No NCCL call is left inside the ALS sweep, so the sharded sweep can be captured in a CUDA graph like the single-GPU
one, and its result is bit-identical on every rank.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64}


class P2PComm:
    def __init__(self, dist, group, device: torch.device, max_payload_bytes: int = 8 << 20):
        self.dist, self.group, self.device = dist, group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.max_payload = int(max_payload_bytes)
        lib = _lib.load()
        self.lib = lib
        self.nbytes = lib.tlb200_comm_buffer_bytes(self.world, self.max_payload)
        if self.nbytes == 0:
            raise RuntimeError(f"p2p all-reduce supports up to 16 ranks, got {self.world}")
        self.local = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.check(lib.tlb200_comm_alloc(self.nbytes, ctypes.byref(self.local), handle), "comm_alloc")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self.peers = []
            for p, h in enumerate(handles):
                if p == self.rank:
                    self.peers.append(self.local.value)
                    continue
                ptr = ctypes.c_void_p()
                _lib.check(lib.tlb200_comm_open(ctypes.create_string_buffer(h, 64), ctypes.byref(ptr)), "comm_open")
                self.peers.append(ptr.value)
        self._bufs = _lib.ptr_array(self.peers)
        # everyone has mapped everyone before the first kernel touches a peer
        dist.barrier(group=group)

    def fits(self, t: torch.Tensor) -> bool:
        return (t.is_cuda and t.dtype in _DTYPES and t.is_contiguous() and t.device == self.device
                and t.numel() * t.element_size() <= self.max_payload)

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks (same bits on every rank)."""
        with torch.cuda.device(self.device):
            st = self.lib.tlb200_allreduce_oneshot(t.data_ptr(), t.data_ptr(), t.numel(), _DTYPES[t.dtype], self._bufs,
                                                   self.world, self.rank, self.max_payload,
                                                   torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "allreduce_oneshot")
        return t

    def all_reduce_partials(self, part, tail, out: torch.Tensor) -> torch.Tensor:
        """out[: rows * rank] = sum over ranks of (sum over splits of the MTTKRP partials), out[rows * rank:] = sum over
        ranks of `tail` (a small contiguous tensor, or None) — the split-K reduction rides in the push phase of the
        exchange (tlb200_allreduce_partials)."""
        rows, rank = part.shape
        n_tail = 0 if tail is None else tail.numel()
        if out.numel() != rows * rank + n_tail or not out.is_contiguous():
            raise ValueError("all_reduce_partials: out must be contiguous with rows * rank + tail elements")
        with torch.cuda.device(self.device):
            st = self.lib.tlb200_allreduce_partials(ctypes.byref(part.info), tail.data_ptr() if tail is not None else None,
                                                    n_tail, out.data_ptr(), _DTYPES[out.dtype], self._bufs, self.world,
                                                    self.rank, self.max_payload,
                                                    torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "allreduce_partials")
        return out

    def close(self) -> None:
        if getattr(self, "peers", None) is None:
            return
        torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)            # nobody unmaps while a peer may still push
        for p, ptr in enumerate(self.peers):
            if p != self.rank:
                self.lib.tlb200_comm_close(ctypes.c_void_p(ptr))
        self.lib.tlb200_comm_free(self.local)
        self.peers = None
