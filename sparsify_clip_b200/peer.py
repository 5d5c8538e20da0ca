"""SM-free all-gather of row shards between the GPUs of one node (exchange step 1 and 2).

NCCL's all-gather is a kernel: launched under a persistent sweep it takes SMs the sweep's clusters were sized for
(measured at 8 GPUs: 270 us instead of 217 us per sweep) and, issued before one, it is fully exposed (~180 us for
8 x 4 MB in the LL protocol).  Here every rank PUSHES its shard into every peer's gather buffer with the copy engines
(`scb_peer_push`: cudaMemcpyAsync on peer-mapped pointers, NVLink), then writes a 4-byte epoch flag per peer on the same
stream; the consumer's only kernel is a one-warp wait on its own flag words (`scb_wait_flags`).

Buffers are allocated by the library (`scb_peer_alloc`) and exchanged once per (group, role, shard size) through CUDA
IPC handles.  Each role ("I", "T", ...) owns TWO buffers used alternately: a rank can only be two gathers ahead of a peer
after having seen that peer's flag of the gather in between, and a peer pushes only after everything it enqueued before
(the sweeps that read the older buffer) has finished -- so the buffer being overwritten is no longer read anywhere.

Falls back to NCCL (the caller does) when the ranks are not all on one node, when CUDA IPC is unavailable, or while the
stream is being captured into a CUDA graph (cross-device copies are not capturable).
"""
import ctypes
import os
import socket

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check

_EPOCHS = 1 << 22          # table of epoch values on the device (16 MB): 4 M gathers per role before it is exhausted
_state = {"mode": os.environ.get("SCB_GATHER", "auto"), "objs": {}, "ok": {}, "table": {}}


def set_gather(mode):
    """'auto' (peer pushes when possible, else NCCL), 'peer' (same, but raise when impossible), 'nccl'."""
    if mode not in ("auto", "peer", "nccl"):
        raise ValueError("gather mode must be 'auto', 'peer' or 'nccl'")
    prev, _state["mode"] = _state["mode"], mode
    return prev


class _Raw:
    """CUDA array interface over a raw device pointer (zero-copy torch view of library-owned memory)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class _Handle:
    def __init__(self, pg, buf, epoch, keep):
        self.pg, self.buf, self.epoch, self.keep = pg, buf, epoch, keep

    def wait(self):
        pg = self.pg
        st = torch.cuda.current_stream(pg.device).cuda_stream
        check(pg.lib.scb_wait_flags(pg.base + pg.flag_off + 4 * self.buf * pg.world, pg.world, self.epoch, st), "wait_flags")
        self.keep = None


class PeerGather:
    """One role's double-buffered gather area on every rank of `group` (all ranks on this node)."""

    def __init__(self, group, shard_bytes, device):
        self.lib = _lib.load()
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.shard = int(shard_bytes)
        self.buf_bytes = ((self.shard * self.world + 255) // 256) * 256
        self.flag_off = 2 * self.buf_bytes
        total = self.flag_off + 4096
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            check(self.lib.scb_peer_alloc(total, ctypes.byref(ptr), handle), "peer_alloc")
        self.base = ptr.value
        mine = (socket.gethostname(), device.index, handle.raw)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if len({h for h, _, _ in everyone}) != 1:
            raise RuntimeError("peer gather needs all ranks on one node")
        self.ptrs = []
        with torch.cuda.device(device):
            for r, (_, _, hb) in enumerate(everyone):
                if r == self.rank:
                    self.ptrs.append(self.base)
                else:
                    p = ctypes.c_void_p()
                    check(self.lib.scb_peer_open(hb, ctypes.byref(p)), "peer_open")
                    self.ptrs.append(p.value)
        self.view = torch.as_tensor(_Raw(self.base, 2 * self.buf_bytes), device=device)
        self.stream = torch.cuda.Stream(device=device)
        key = device.index
        if key not in _state["table"]:
            _state["table"][key] = torch.arange(_EPOCHS, dtype=torch.int32, device=device)
        self.table = _state["table"][key]
        self.epoch = 0
        order = [(self.rank + k) % self.world for k in range(1, self.world)] + [self.rank]     # peers first, then myself
        VP = ctypes.c_void_p * self.world
        self._dst = [VP(*[self.ptrs[p] + b * self.buf_bytes + self.rank * self.shard for p in order]) for b in (0, 1)]
        self._flag = [VP(*[self.ptrs[p] + self.flag_off + 4 * (b * self.world + self.rank) for p in order]) for b in (0, 1)]
        dist.barrier(group=group)              # everybody has mapped everybody before the first push

    def gather(self, x):
        """x: contiguous tensor of `shard` bytes on self.device.  -> (flat uint8 view of the gathered buffer, handle)."""
        assert x.is_contiguous() and x.numel() * x.element_size() == self.shard
        self.epoch += 1
        if self.epoch >= _EPOCHS:
            raise RuntimeError("peer gather: epoch table exhausted")
        b = self.epoch & 1
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)           # x, and every earlier read of the buffer about to be overwritten elsewhere
        x.record_stream(self.stream)
        check(self.lib.scb_peer_push(x.data_ptr(), self.shard, self._dst[b], self._flag[b], self.world,
                                     self.table.data_ptr() + 4 * self.epoch, self.stream.cuda_stream), "peer_push")
        out = self.view[b * self.buf_bytes:b * self.buf_bytes + self.shard * self.world]
        return out, _Handle(self, b, self.epoch, x)


def available(group, device):
    """True when this process group can use peer pushes (decided once per group, identically on every rank)."""
    if _state["mode"] == "nccl" or group is None:
        return False
    key = (id(group), device.index)
    if key not in _state["ok"]:
        ok = 1
        try:
            if dist.get_backend(group) != "nccl" or not torch.cuda.is_available():
                ok = 0
            else:
                PeerGather(group, 1024, device)                # probe: IPC exchange + mapping works on every rank
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        _state["ok"][key] = bool(flag.item())
        if not _state["ok"][key] and _state["mode"] == "peer":
            raise RuntimeError("peer gather requested (set_gather('peer')) but CUDA IPC between the ranks is unavailable")
    return _state["ok"][key]


def all_gather_async(x, group, role):
    """All-gather the contiguous tensor x along dim 0 with peer pushes.  -> (gathered tensor, handle) or None when the
    peer path does not apply (the caller then uses NCCL)."""
    if not x.is_cuda or torch.cuda.is_current_stream_capturing() or not available(group, x.device):
        return None
    x = x.contiguous()
    nbytes = x.numel() * x.element_size()
    key = (id(group), x.device.index, role, nbytes)
    pg = _state["objs"].get(key)
    if pg is None:
        pg = _state["objs"][key] = PeerGather(group, nbytes, x.device)
    flat, h = pg.gather(x)
    out = flat.view(x.dtype).view((pg.world * x.shape[0],) + tuple(x.shape[1:]))
    return out, h
