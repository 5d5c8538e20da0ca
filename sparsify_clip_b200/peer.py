"""SM-free all-gather of row shards between the GPUs of one node (exchange steps 1 and 2).

NCCL's all-gather is a kernel: launched under a persistent sweep it takes SMs the sweep's clusters were sized for
(measured at 8 GPUs: 270 us instead of 217 us per sweep) and, issued before one, it is fully exposed (~180 us for
8 x 4 MB in the LL protocol).  Here every rank PUSHES its shard into every peer's gather buffer with the copy engines
(`scb_peer_push`: cudaMemcpyAsync on peer-mapped pointers, NVLink); the only kernels are one-warp flag kernels.

Protocol per role ("I", "T", "P", ...), one gather buffer per rank exchanged once through CUDA IPC:
    begin    (consumer's stream)  ++epoch; wait until every peer RELEASED the previous contents of its buffer
    push     (side stream)        my shard -> slot[rank] of every rank's buffer; then arrived[rank] := epoch on every rank
    wait     (consumer's stream)  until arrived[p] >= epoch for every p: the gathered buffer is complete
    release  (consumer's stream, after the last read)   done[rank] := epoch on every rank
The epoch lives in device memory and every flag kernel reads it there, so a step that contains a gather can be captured
into a CUDA graph and replayed (nothing about the epoch is baked into a launch).  A missing `release` (an exception
between gather and release) makes the peers' next `begin` wait until the library's watchdog ends the job.

Falls back to NCCL (the caller does) when the ranks are not all on one node or CUDA IPC is unavailable.
"""
import ctypes
import os
import socket

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check

_state = {"mode": os.environ.get("SCB_GATHER", "auto"), "objs": {}, "ok": {}, "open": [],
          # roles whose pushes are done by the SMs instead of the copy engines (scb_peer_push_sm): the gather a step starts
          # with, when no sweep occupies the SMs yet.  SCB_PEER_SM_ROLES="" turns it off, "I,T" widens it.
          "sm_roles": {r for r in os.environ.get("SCB_PEER_SM_ROLES", "I").split(",") if r}}
_SM_PUSH_MIN_BYTES = 1 << 18          # below this the copy engines' fixed cost is what counts either way


def set_sm_push_roles(roles):
    """Roles ('I', 'T', 'C', ...) pushed by an SM kernel; returns the previous set.  Takes effect for the next gather."""
    prev, _state["sm_roles"] = _state["sm_roles"], set(roles)
    return prev


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
    def __init__(self, pg, keep):
        self.pg, self.keep = pg, keep

    def wait(self):
        pg = self.pg
        st = torch.cuda.current_stream(pg.device).cuda_stream
        check(pg.lib.scb_peer_wait(pg.epoch.data_ptr(), pg.base + pg.arr_off, pg.world, st), "peer_wait")
        self.keep = None


class PeerGather:
    """One role's gather buffer on every rank of `group` (all ranks on this node)."""

    def __init__(self, group, shard_bytes, device, role=None):
        self.lib = _lib.load()
        self.group, self.device, self.role = group, device, role
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 32:
            raise RuntimeError("peer gather supports up to 32 ranks")
        self.shard = int(shard_bytes)
        self.buf_bytes = ((self.shard * self.world + 255) // 256) * 256
        self.arr_off, self.done_off = self.buf_bytes, self.buf_bytes + 128
        total = self.buf_bytes + 256
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            check(self.lib.scb_peer_alloc(total, ctypes.byref(ptr), handle), "peer_alloc")     # zero-filled
        self.base = ptr.value
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (socket.gethostname(), handle.raw), group=group)
        if len({h for h, _ in everyone}) != 1:
            raise RuntimeError("peer gather needs all ranks on one node")
        self.ptrs = []
        with torch.cuda.device(device):
            for r, (_, hb) in enumerate(everyone):
                if r == self.rank:
                    self.ptrs.append(self.base)
                else:
                    p = ctypes.c_void_p()
                    check(self.lib.scb_peer_open(hb, ctypes.byref(p)), "peer_open")
                    self.ptrs.append(p.value)
        self.view = torch.as_tensor(_Raw(self.base, self.buf_bytes), device=device)
        self.stream = torch.cuda.Stream(device=device)
        # Optional fan-out of a large shard's pushes over several streams (SCB_PEER_FAN=k extra streams).  Off by default:
        # measured at 8 GPUs it made the step slower (1.32 ms vs 1.26 ms with one stream per role), the extra fork / join
        # events cost more than the parallel copies save at 4 MB per peer.
        nfan = int(os.environ.get("SCB_PEER_FAN", "0"))
        self.fan = [torch.cuda.Stream(device=device) for _ in range(nfan)] if self.shard >= (1 << 20) and self.world > 2 else []
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)      # CTA counter of scb_peer_push_sm
        i64 = dict(dtype=torch.int64, device=device)
        self.arrived_words = torch.tensor([p + self.arr_off + 4 * self.rank for p in self.ptrs], **i64)
        self.done_words = torch.tensor([p + self.done_off + 4 * self.rank for p in self.ptrs], **i64)
        order = [(self.rank + k) % self.world for k in range(1, self.world)] + [self.rank]     # peers first, then myself
        dsts = [self.ptrs[p] + self.rank * self.shard for p in order]
        self._dst = (ctypes.c_void_p * self.world)(*dsts)
        nf = len(self.fan) + 1
        self._fan_dst = [(ctypes.c_void_p * len(dsts[k::nf]))(*dsts[k::nf]) for k in range(nf)]
        self.open_reads = False
        torch.cuda.synchronize(device)
        dist.barrier(group=group)              # everybody has mapped everybody (and zeroed its flags) before the first push

    def gather(self, x):
        """x: contiguous tensor of `shard` bytes on self.device.  -> (flat uint8 view of the gather buffer, handle)."""
        assert x.is_contiguous() and x.numel() * x.element_size() == self.shard
        if self.open_reads:
            raise RuntimeError("peer gather: the previous result of this role was never released (peer.release_all)")
        cur = torch.cuda.current_stream(self.device)
        check(self.lib.scb_peer_begin(self.epoch.data_ptr(), self.base + self.done_off, self.world, cur.cuda_stream), "peer_begin")
        self.stream.wait_stream(cur)           # x is ready, the epoch is advanced, the peers have released their buffers
        # (x is kept alive by the handle until wait() is enqueued: the wait kernel only passes once my own pushes -- the
        #  last thing to read x -- have set my arrived flag, and any reuse of x's memory is ordered after it on `cur`)
        if (self.role in _state["sm_roles"] and self.shard >= _SM_PUSH_MIN_BYTES and self.shard % 16 == 0
                and x.data_ptr() % 16 == 0):
            check(self.lib.scb_peer_push_sm(x.data_ptr(), self.shard, self._dst, self.world, self.epoch.data_ptr(),
                                            self.arrived_words.data_ptr(), self.world, self.counter.data_ptr(),
                                            self.stream.cuda_stream), "peer_push_sm")
        elif self.fan:
            for k, st in enumerate(self.fan):      # copies only (n destinations, no flag: world = 0 skips nothing but the
                st.wait_stream(self.stream)        # signal is sent once, below, after every fan stream has been joined)
                check(self.lib.scb_peer_copy(x.data_ptr(), self.shard, self._fan_dst[k + 1], len(self._fan_dst[k + 1]),
                                             st.cuda_stream), "peer_copy")
            check(self.lib.scb_peer_copy(x.data_ptr(), self.shard, self._fan_dst[0], len(self._fan_dst[0]),
                                         self.stream.cuda_stream), "peer_copy")
            for st in self.fan:
                self.stream.wait_stream(st)
            check(self.lib.scb_peer_push(x.data_ptr(), 0, self._dst, 0, self.epoch.data_ptr(),
                                         self.arrived_words.data_ptr(), self.world, self.stream.cuda_stream), "peer_push")
        else:
            check(self.lib.scb_peer_push(x.data_ptr(), self.shard, self._dst, self.world, self.epoch.data_ptr(),
                                         self.arrived_words.data_ptr(), self.world, self.stream.cuda_stream), "peer_push")
        self.open_reads = True
        if self not in _state["open"]:
            _state["open"].append(self)
        return self.view[:self.shard * self.world], _Handle(self, x)

    def release(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)           # joins the side stream (required inside a graph capture; cheap otherwise)
        check(self.lib.scb_peer_release(self.epoch.data_ptr(), self.done_words.data_ptr(), self.world, cur.cuda_stream),
              "peer_release")
        self.open_reads = False


def release_all():
    """After the last read of every gathered buffer of this step (stream-ordered): let the peers overwrite them.
    One launch for all roles of a device (the flag kernels are tiny, but at 8 GPUs a step is only ~1.2 ms)."""
    todo = [pg for pg in _state["open"] if pg.open_reads]
    _state["open"].clear()
    while todo:
        dev = todo[0].device
        batch = [pg for pg in todo if pg.device == dev and pg.world == todo[0].world][:8]
        todo = [pg for pg in todo if pg not in batch]
        cur = torch.cuda.current_stream(dev)
        for pg in batch:
            cur.wait_stream(pg.stream)         # joins the side streams (required inside a graph capture)
        n = len(batch)
        ep = (ctypes.c_void_p * n)(*[pg.epoch.data_ptr() for pg in batch])
        dw = (ctypes.c_void_p * n)(*[pg.done_words.data_ptr() for pg in batch])
        check(batch[0].lib.scb_peer_release_many(ep, dw, n, batch[0].world, cur.cuda_stream), "peer_release_many")
        for pg in batch:
            pg.open_reads = False


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
            elif torch.cuda.is_current_stream_capturing():
                ok = 0                                          # the one-time IPC exchange cannot happen inside a capture
            else:
                PeerGather(group, 1024, device)                 # probe: IPC exchange + mapping works on every rank
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
    peer path does not apply (the caller then uses NCCL).  The result lives in the role's buffer until release_all()."""
    if not x.is_cuda:
        return None
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing and (id(group), x.device.index) not in _state["ok"]:
        return None
    if not available(group, x.device):
        return None
    x = x.contiguous()
    nbytes = x.numel() * x.element_size()
    key = (id(group), x.device.index, role, nbytes)
    pg = _state["objs"].get(key)
    if pg is None:
        if capturing:
            return None                       # buffers must exist before a capture (run one eager step first)
        pg = _state["objs"][key] = PeerGather(group, nbytes, x.device, role)
    flat, h = pg.gather(x)
    out = flat.view(x.dtype).view((pg.world * x.shape[0],) + tuple(x.shape[1:]))
    return out, h
