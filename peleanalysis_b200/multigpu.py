"""One-process-per-GPU plumbing above the C ABI: torch.distributed moves what has to move between processes --
CUDA-IPC handles once per field (peer links), and the packed ghost slabs of whatever is NOT peer-linked per fill.
No compute happens here."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import capi


class _DevArray:
    """Raw device pointer -> torch tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def map_peers(field: capi.Field) -> None:
    """Collective: every rank maps every other rank's slabs of `field` (PA_HIER_PEER_LINKS hierarchies)."""
    H = field.hier
    if H.nranks == 1:
        return
    mine = field.ipc_handles()
    got = [None] * H.nranks
    dist.all_gather_object(got, mine)
    for r, h in enumerate(got):
        if r != H.rank:
            field.map_peer(r, h)
    dist.barrier()


def stream_barrier() -> None:
    """Cross-rank ordering point on the current CUDA stream without a host sync: a 1-element all-reduce.  Peers'
    writes enqueued before it are visible to kernels this rank enqueues after it."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.zeros(1, device="cuda")
        dist.all_reduce(t)


class SlabExchange:
    """NCCL send/recv of the packed ghost slabs for ncomp components of one field (everything the neighbour links do
    not cover: ragged same-level neighbours and coarse cells of coarse-fine faces owned by other ranks)."""

    def __init__(self, field: capi.Field, ncomp: int):
        H = field.hier
        self.field, self.ncomp, self.rank, self.world = field, ncomp, H.rank, H.nranks
        self.soff = self.roff = [0]
        self.empty = True
        if self.world == 1:
            return
        sp, rp = C.c_void_p(), C.c_void_p()
        so = (C.c_int64 * (self.world + 1))()
        ro = (C.c_int64 * (self.world + 1))()
        capi.check(capi.lib().pa_exchange_buffers(field.f, ncomp, C.byref(sp), C.byref(rp), so, ro))
        self.soff, self.roff = list(so), list(ro)
        self.send_t = torch.as_tensor(_DevArray(sp.value, max(self.soff[-1], 1)), device="cuda")
        self.recv_t = torch.as_tensor(_DevArray(rp.value, max(self.roff[-1], 1)), device="cuda")
        # a rank with nothing to send or receive still takes part if any peer does (batch_isend_irecv is pairwise)
        self.empty = (self.soff[-1] == 0 and self.roff[-1] == 0)

    def run(self, comp: int = 0) -> None:
        if self.world == 1 or self.empty:
            return
        f = self.field
        capi.check(capi.lib().pa_exchange_pack(f.f, comp, self.ncomp))
        ops = []
        for p in range(self.world):
            if p == self.rank:
                continue
            if self.roff[p + 1] > self.roff[p]:
                ops.append(dist.P2POp(dist.irecv, self.recv_t[self.roff[p]:self.roff[p + 1]], p))
            if self.soff[p + 1] > self.soff[p]:
                ops.append(dist.P2POp(dist.isend, self.send_t[self.soff[p]:self.soff[p + 1]], p))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        capi.check(capi.lib().pa_exchange_mark_received(f.f, comp, self.ncomp))
