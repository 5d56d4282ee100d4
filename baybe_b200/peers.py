"""Peer-memory arg-max reduction across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL for the plumbing).  Each rank owns a 64-byte device buffer with
two packed-key slots and two arrival counters; the buffers are exchanged ONCE with CUDA IPC (torch's own storage
sharing, the mechanism ``torch.multiprocessing`` uses) so that every rank holds NVLink-mapped pointers to every
peer's slots.  ``PeerReduce.allreduce_best`` then enqueues ``bb_allreduce_best`` -- one warp that folds the local
key into every rank's slot with ``atomicMax.sys`` and waits for ``world`` arrivals -- on the caller's stream:
per greedy round the step is [scoring kernel][one-warp kernel], with no host code and no NCCL call in between.
Round 1 issued an 8-byte ``dist.all_reduce`` from the host every step, which halved 8-GPU efficiency.
"""
from __future__ import annotations

import ctypes as C

import torch

from baybe_b200 import _lib

__all__ = ["PeerReduce", "get_peer_reduce"]


class PeerReduce:
    """Slots of this rank + mapped slots of all peers; create it collectively (every rank, same order)."""

    kind = "peer"

    def __init__(self, device: torch.device):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerReduce needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > _lib.MAX_PEERS:
            raise NotImplementedError(f"peer reduction supports up to {_lib.MAX_PEERS} GPUs of one box")
        self.device = torch.device(device)
        self.lib = _lib.load()
        with torch.cuda.device(self.device):
            # one private allocation (not a slice of a cached block): [key0 key1 | cnt0 cnt1 | pad]
            self.buf = torch.zeros(8, dtype=torch.int64, device=self.device)
            _lib.check(self.lib.bb_peer_slots_init(C.c_void_p(self.buf.data_ptr()),
                                                   C.c_void_p(self.buf.data_ptr() + 16),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "bb_peer_slots_init")
            torch.cuda.current_stream().synchronize()
        storage = self.buf.untyped_storage()
        info = storage._share_cuda_()  # (device, ipc handle, bytes, offset, ref-counter handle, ..., event handle, sync)
        offset = self.buf.storage_offset() * self.buf.element_size()
        infos: list = [None] * self.world
        dist.all_gather_object(infos, (info, offset))
        self._peer_storages = []
        group = _lib.PeerGroup()
        group.rank, group.world = self.rank, self.world
        for r, (inf, off) in enumerate(infos):
            if r == self.rank:
                base = self.buf.data_ptr()
            else:
                # cudaIpcOpenMemHandle(cudaIpcMemLazyEnablePeerAccess) must run with THIS rank's device current: it
                # is the opening device that gets peer access to the exporter.  torch opens under a guard for the
                # device index in the tuple -- the exporter's index as shipped; opened that way the pointer is only
                # valid for the exporter's device and a kernel on this device faults (first 2-GPU run, round 2).
                inf = (self.device.index,) + tuple(inf[1:])
                st = torch.UntypedStorage._new_shared_cuda(*inf)
                self._peer_storages.append(st)
                base = st.data_ptr() + off
            group.d_key[r] = base
            group.d_count[r] = base + 16
        self.group = group
        self.epoch = 0
        self.out = torch.empty(1, dtype=torch.int64, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier()  # every rank has mapped every peer before the first reduction

    def allreduce_best(self, key: torch.Tensor) -> torch.Tensor:
        """Global maximum of the ranks' packed keys (device tensor of shape [1]); enqueued on the current stream.
        The returned tensor is overwritten by the next call."""
        if key.device != self.device or key.dtype != torch.int64 or key.numel() != 1:
            raise ValueError("key must be a 1-element int64 tensor on this rank's device")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bb_allreduce_best(C.byref(self.group), C.c_void_p(key.data_ptr()),
                                                  C.c_uint32(self.epoch & 0xFFFFFFFF), C.c_void_p(self.out.data_ptr()),
                                                  C.c_void_p(self.status.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "bb_allreduce_best")
        self.epoch += 1
        return self.out

    def check(self) -> None:
        """Raise if a reduction timed out waiting for a peer (synchronises)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("bb_allreduce_best: a peer rank did not arrive within the time-out")


class NcclReduce:
    """Escape hatch (BB_PEER_REDUCE=0): the same reduction as one host-issued ``ncclAllReduce(MAX, int64)`` -- the
    round-1 path, kept for boxes on which CUDA IPC peer mappings are unavailable.  Not the default."""

    kind = "nccl"

    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.out = torch.empty(1, dtype=torch.int64, device=self.device)

    def allreduce_best(self, key: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        self.out.copy_(key)
        dist.all_reduce(self.out, op=dist.ReduceOp.MAX)
        return self.out

    def check(self) -> None:
        return None


_instances: dict = {}


def get_peer_reduce(device) -> PeerReduce:
    """Process-wide instance per device; the first call is collective."""
    dev = torch.device(device)
    key = (dev.type, dev.index)
    if key not in _instances:
        import os

        _instances[key] = NcclReduce(dev) if os.environ.get("BB_PEER_REDUCE", "1") == "0" else PeerReduce(dev)
    return _instances[key]
