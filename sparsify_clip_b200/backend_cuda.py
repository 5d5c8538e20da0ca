"""CUDA backend: torch tensors in, libscb200.so (C ABI, include/scb200.h) underneath.

PyTorch is only plumbing here: it owns the device buffers and the stream.  All arithmetic
on the hot path happens in the library's kernels.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import PATH_SIMT, PATH_TC, SCB_BF16, SCB_F16, SCB_F32, check

_DT = {torch.float32: SCB_F32, torch.bfloat16: SCB_BF16, torch.float16: SCB_F16}

# How fp32 inputs are computed: "exact" = fp32 CUDA-core kernels (parity gate: 1e-5 on gradients),
# "bf16" = round once to bf16 and use the tensor-core path.
_fp32_mode = "exact"
_force_path = None  # testing hook: PATH_SIMT / PATH_TC / None


# Every device buffer the backend hands to the library comes from these two hooks (PyTorch's caching allocator in
# production).  tests/ swap them for a guarded allocator that surrounds each buffer with canaries and verifies, after the
# kernels ran, that nothing wrote outside it.
_empty = torch.empty
_zeros = torch.zeros


def set_allocator(empty_fn=None, zeros_fn=None):
    """Test hook: replace the buffer allocators (None restores torch.empty / torch.zeros)."""
    global _empty, _zeros
    prev = (_empty, _zeros)
    _empty = empty_fn or torch.empty
    _zeros = zeros_fn or torch.zeros
    return prev


def set_fp32_mode(mode):
    global _fp32_mode
    if mode not in ("exact", "bf16"):
        raise ValueError("fp32 mode must be 'exact' or 'bf16'")
    prev, _fp32_mode = _fp32_mode, mode
    return prev


def force_path(path):
    global _force_path
    prev, _force_path = _force_path, path
    return prev


def choose_jparts(n_rb, nsplit, n_jb, n_sm=148, max_parts=16, overhead=1.0):
    """Split the column sweep so that (row blocks x column groups x parts) fills the SMs evenly.

    cost model: rounds of `n_sm` concurrent work items x (tiles per item + `overhead` tiles of prologue/drain;
    1 for the single-CTA kernels, 4 for the CTA-pair kernel).
    Host mirror of the planner inside the library (scb_pass_plan, csrc/api.cu), which is the one the
    backend uses; tests/test_host.py checks that the two agree.
    """
    best, best_cost = 1, None
    for jp in range(1, max(1, min(n_jb, max_parts)) + 1):
        rounds = math.ceil(n_rb * nsplit * jp / n_sm)
        cost = rounds * (math.ceil(n_jb / jp) + overhead)
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = jp, cost
    return best


def pair_span_plan(n_rb, n_jb, n_sm=148):
    """Host mirror of the CTA-pair kernel's work split (csrc/tc_pair.cu: scb_pair_span_plan): the linearised
    (row block, column tile) space is cut into equal contiguous spans, one per CTA pair.
    -> (pairs, tiles per pair, partial slots = largest number of segments a row block is cut into)."""
    total = n_rb * n_jb
    pairs = max(1, min(n_sm // 2, total))
    span = max(1, -(-total // pairs), (n_jb + 14) // 15)      # at most 16 partial slots per row block
    pairs = max(1, -(-total // span))
    pmax = max(((rb + 1) * n_jb - 1) // span - (rb * n_jb) // span + 1 for rb in range(n_rb)) if n_rb else 1
    return pairs, span, pmax


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# Host-side cost matters for small batches (a B=4096 step is ~0.45 ms of GPU time and ~50 launches): the public
# torch.cuda.device() / current_stream() helpers cost several microseconds per call, their C accessors a fraction of one.
_get_device = torch._C._cuda_getDevice
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


class _On:
    """`with _On(dev)` that does nothing when `dev` is already current (the usual case)."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index

    def __enter__(self):
        self.prev = _get_device()
        if self.idx is None:
            self.idx = self.prev
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


class CudaBackend:
    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        self.launches = 0          # kernels launched through this backend (bench.py reports it)
        self.pass_events = None    # when a list: (name, start_event, end_event) per B x B pass
        self._sm_count = {}

    def _count(self, n=1):
        self.launches += n

    class _Timed:
        """Optional CUDA-event bracket around one B x B pass (enabled by bench.py only)."""

        def __init__(self, be, name):
            self.be, self.name = be, name

        def __enter__(self):
            if self.be.pass_events is not None:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            return self

        def __exit__(self, *exc):
            if self.be.pass_events is not None:
                self.e1.record()
                self.be.pass_events.append((self.name, self.e0, self.e1))
            return False

    # ------------------------------------------------------------------ tensor plumbing
    def prep(self, x, cast_fp32=True):
        if not isinstance(x, torch.Tensor) or x.dim() != 2:
            raise ValueError("expected a 2-D [B, D] tensor")
        if not x.is_cuda:
            raise RuntimeError("sparsify_clip_b200 runs on CUDA tensors only (no CPU fallback)")
        x = x.detach()
        if x.dtype not in _DT:
            x = x.float()
        if x.dtype == torch.float32 and _fp32_mode == "bf16" and cast_fp32:
            x = x.to(torch.bfloat16)
        if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
            x = x.contiguous()
        return x

    def path_for(self, *mats):
        if _force_path is not None:
            return _force_path
        for m in mats:
            if m.dtype == torch.float32 or m.shape[1] % 8 or m.stride(0) % 8 or m.data_ptr() % 16:
                return PATH_SIMT
        # below one 128-column tile the 16-bit weight tile is not averaged over enough pairs to stay inside
        # the 1e-3 gradient gate (and a single CTA of fp32 FMAs is just as fast): exact path
        if min(m.shape[0] for m in mats) < 128:
            return PATH_SIMT
        return PATH_TC

    @staticmethod
    def _stream():
        """The caller's stream on the current device (every use sits inside `with _On(device)`)."""
        if _raw_stream is not None:
            return _raw_stream(_get_device())
        return torch.cuda.current_stream().cuda_stream

    def _n_sm(self, dev):
        idx = dev.index if dev.index is not None else _get_device()
        n = self._sm_count.get(idx)
        if n is None:
            n = self._sm_count[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
        return n

    def _plan(self, path, nA, nB, D, grad, dev):
        """(jparts, nsub) of one B x B pass, from the library's planner."""
        jp, nsub = ctypes.c_int(0), ctypes.c_int(0)
        check(self.lib.scb_pass_plan(path, nA, nB, D, int(bool(grad)), self._n_sm(dev), ctypes.byref(jp),
                                     ctypes.byref(nsub)), "pass_plan")
        return jp.value, nsub.value

    def sum(self, x, out=None):
        """Deterministic sum of all elements -> 0-dim fp32 (or written into `out`, a one-element fp32 view)."""
        x = x.contiguous().view(-1)
        if out is None:
            out = _empty((), dtype=torch.float32, device=x.device)
        scratch = _empty(1024, dtype=torch.float32, device=x.device)
        with _On(x.device):
            check(self.lib.scb_sum(_ptr(x), x.numel(), _ptr(scratch), _ptr(out), self._stream()), "sum")
        self._count(2)
        return out

    # ------------------------------------------------------------------ row-wise
    def row_sqnorm(self, x):
        out = _empty(x.shape[0], dtype=torch.float32, device=x.device)
        with _On(x.device):
            check(self.lib.scb_row_sqnorm(_ptr(x), x.shape[0], x.shape[1], x.stride(0), _DT[x.dtype], _ptr(out),
                                          self._stream()), "row_sqnorm")
        self._count()
        return out

    def row_dot(self, a, b):
        out = _empty(a.shape[0], dtype=torch.float32, device=a.device)
        with _On(a.device):
            check(self.lib.scb_row_dot(_ptr(a), _ptr(b), a.shape[0], a.shape[1], a.stride(0), b.stride(0), _DT[a.dtype],
                                       _ptr(out), self._stream()), "row_dot")
        self._count()
        return out

    def lalign_rows(self, x, y):
        out = _empty(x.shape[0], dtype=torch.float32, device=x.device)
        with _On(x.device):
            check(self.lib.scb_lalign_rows(_ptr(x), _ptr(y), x.shape[0], x.shape[1], x.stride(0), y.stride(0),
                                           _DT[x.dtype], _ptr(out), self._stream()), "lalign_rows")
        self._count()
        return out

    def lalign_bwd(self, x, y, host_scale, dev_scale, want_x=True, want_y=True):
        n, D = x.shape
        dX = _empty(n, D, dtype=torch.float32, device=x.device) if want_x else None
        dY = _empty(n, D, dtype=torch.float32, device=x.device) if want_y else None
        with _On(x.device):
            check(self.lib.scb_lalign_bwd(_ptr(x), _ptr(y), n, D, x.stride(0), y.stride(0), _DT[x.dtype], host_scale,
                                          _ptr(dev_scale), 0, _ptr(dX), _ptr(dY), self._stream()), "lalign_bwd")
        self._count()
        return dX, dY

    def centroid_fwd(self, a, b, out_dtype):
        n, D = a.shape
        C = _empty(n, D, dtype=out_dtype, device=a.device)
        inv = _empty(n, dtype=torch.float32, device=a.device)
        with _On(a.device):
            check(self.lib.scb_centroid_fwd(_ptr(a), _ptr(b), n, D, a.stride(0), b.stride(0), _DT[a.dtype], _ptr(C),
                                            _DT[out_dtype], _ptr(inv), self._stream()), "centroid_fwd")
        self._count()
        return C, inv

    def centroid_bwd(self, a, b, dC, inv, host_scale=1.0, dev_scale=None, both=True):
        """(dA, dB): the two are equal element for element; both=False writes one array and returns (dA, None)."""
        n, D = a.shape
        dA = _empty(n, D, dtype=torch.float32, device=a.device)
        dB = _empty(n, D, dtype=torch.float32, device=a.device) if both else None
        with _On(a.device):
            check(self.lib.scb_centroid_bwd(_ptr(a), _ptr(b), n, D, a.stride(0), b.stride(0), _DT[a.dtype], _ptr(dC),
                                            _ptr(inv), host_scale, _ptr(dev_scale), 0, _ptr(dA), _ptr(dB),
                                            self._stream()), "centroid_bwd")
        self._count()
        return dA, dB

    def normalize_fwd(self, x, out_dtype):
        n, D = x.shape
        Y = _empty(n, D, dtype=out_dtype, device=x.device)
        inv = _empty(n, dtype=torch.float32, device=x.device)
        with _On(x.device):
            check(self.lib.scb_normalize_fwd(_ptr(x), n, D, x.stride(0), _DT[x.dtype], _ptr(Y), _DT[out_dtype],
                                             _ptr(inv), self._stream()), "normalize_fwd")
        self._count()
        return Y, inv

    def normalize_bwd(self, x, dY, inv):
        n, D = x.shape
        dX = _empty(n, D, dtype=torch.float32, device=x.device)
        with _On(x.device):
            check(self.lib.scb_normalize_bwd(_ptr(x), n, D, x.stride(0), _DT[x.dtype], _ptr(dY), _ptr(inv), _ptr(dX),
                                             self._stream()), "normalize_bwd")
        self._count()
        return dX

    # ------------------------------------------------------------------ B x B passes
    def lse(self, A, Ball, scale, scale_dev=None):
        """Natural-log row LSE of scale * A @ Ball^T -> [nA] fp32.  (`scale_dev`, here and below: optional one-element
        device tensor multiplying `scale` on the device -- 1/tau of a device-resident temperature, no host sync.)"""
        nA, D = A.shape
        nB = Ball.shape[0]
        path = self.path_for(A, Ball)
        jp, nsub = self._plan(path, nA, nB, D, False, A.device)
        pm = _empty(jp * nsub, nA, dtype=torch.float32, device=A.device)
        pl = torch.empty_like(pm)
        out = _empty(nA, dtype=torch.float32, device=A.device)
        with _On(A.device):
            with self._Timed(self, "lse"):
                check(self.lib.scb_lse_pass(_ptr(A), nA, _ptr(Ball), nB, D, A.stride(0), Ball.stride(0), _DT[A.dtype],
                                            float(scale), jp, _ptr(pm), _ptr(pl), path, _ptr(scale_dev), self._stream()), "lse_pass")
            self._count(2)
            check(self.lib.scb_lse_combine(_ptr(pm), _ptr(pl), jp * nsub, nA, _ptr(out), self._stream()), "lse_combine")
        return out

    def lse_rows_cols(self, A, Bm, scale, scale_dev=None):
        """(row LSE [nA], column LSE [nB]) of scale * A @ Bm^T.  Tensor-core path: ONE sweep (scb_lse2_pass) gives
        both; the exact second sweep over Bm @ A^T is launched conditionally on a device-side norm bound (it returns
        at once for unit-norm rows at tau >= 0.033).  Other paths: two sweeps."""
        nA, D = A.shape
        nB = Bm.shape[0]
        path = self.path_for(A, Bm)
        if path != PATH_TC:
            return self.lse(A, Bm, scale, scale_dev), self.lse(Bm, A, scale, scale_dev)
        dev = A.device
        jp, nsub = self._plan(path, nA, nB, D, False, dev)
        jp2, nsub2 = self._plan(path, nB, nA, D, False, dev)
        n_strips = 4 * ((nA + 127) // 128)
        f32 = dict(dtype=torch.float32, device=dev)
        nsub = 4                      # the fused sweep runs 16 epilogue warps: four 32-column slices per tile
        pm, pl = _empty(jp * nsub, nA, **f32), _empty(jp * nsub, nA, **f32)
        cref, csum = _empty(n_strips, (nB + 31) // 32, **f32), _empty(n_strips, nB, **f32)
        r, c = _empty(nA, **f32), _empty(nB, **f32)
        flag = _empty(1, dtype=torch.int32, device=dev)
        pm2, pl2 = _empty(jp2 * nsub2, nB, **f32), _empty(jp2 * nsub2, nB, **f32)
        sqa, sqb = self.row_sqnorm(A), self.row_sqnorm(Bm)
        st = self._stream()
        with _On(dev):
            check(self.lib.scb_lse2_spread_flag(_ptr(sqa), nA, _ptr(sqb), nB, float(scale), _ptr(flag), _ptr(scale_dev), st),
                  "lse2_spread_flag")
            with self._Timed(self, "lse"):
                check(self.lib.scb_lse2_pass(_ptr(A), nA, _ptr(Bm), nB, D, A.stride(0), Bm.stride(0), _DT[A.dtype],
                                             float(scale), jp, _ptr(pm), _ptr(pl), _ptr(cref), _ptr(csum), _ptr(scale_dev), st),
                      "lse2_pass")
            check(self.lib.scb_lse_combine(_ptr(pm), _ptr(pl), jp * nsub, nA, _ptr(r), st), "lse_combine")
            check(self.lib.scb_colstat_combine(_ptr(cref), _ptr(csum), n_strips, nB, _ptr(c), st), "colstat_combine")
            with self._Timed(self, "lse"):
                check(self.lib.scb_lse_pass_cond(_ptr(Bm), nB, _ptr(A), nA, D, Bm.stride(0), A.stride(0), _DT[A.dtype],
                                                 float(scale), jp2, _ptr(pm2), _ptr(pl2), _ptr(flag), _ptr(scale_dev), st),
                      "lse_pass_cond")
            check(self.lib.scb_lse_combine_cond(_ptr(pm2), _ptr(pl2), jp2 * nsub2, nB, _ptr(c), _ptr(flag), st),
                  "lse_combine_cond")
        self._count(6)
        return r, c

    def lse_rows_colparts(self, A, Bm_all, Bm_rows, A_all, scale, scale_dev=None):
        """Row-sharded variant of lse_rows_cols.  A = my rows, Bm_all = all columns, Bm_rows = my rows of the column
        side, A_all = all rows.  Returns None off the tensor-core path, else
        (r [nA], M [nB], L [nB], c_exact_rows [nA], flag): column partial over MY rows = L * 2^M (log2 domain);
        c_exact_rows / flag = the conditional exact sweep (valid where flag != 0; flag is a global norm bound, the
        same on every rank)."""
        nA, D = A.shape
        nB = Bm_all.shape[0]
        path = self.path_for(A, Bm_all)
        if path != PATH_TC:
            return None
        dev = A.device
        jp, _ = self._plan(path, nA, nB, D, False, dev)
        jp2, nsub2 = self._plan(path, Bm_rows.shape[0], A_all.shape[0], D, False, dev)
        n_strips = 4 * ((nA + 127) // 128)
        f32 = dict(dtype=torch.float32, device=dev)
        nsub = 4
        pm, pl = _empty(jp * nsub, nA, **f32), _empty(jp * nsub, nA, **f32)
        cref, csum = _empty(n_strips, (nB + 31) // 32, **f32), _empty(n_strips, nB, **f32)
        r, M, L = _empty(nA, **f32), _empty(nB, **f32), _empty(nB, **f32)
        nR = Bm_rows.shape[0]
        c_exact = _zeros(nR, **f32)
        flag = _empty(1, dtype=torch.int32, device=dev)
        pm2, pl2 = _empty(jp2 * nsub2, nR, **f32), _empty(jp2 * nsub2, nR, **f32)
        sqa, sqb = self.row_sqnorm(A_all), self.row_sqnorm(Bm_all)
        st = self._stream()
        with _On(dev):
            check(self.lib.scb_lse2_spread_flag(_ptr(sqa), A_all.shape[0], _ptr(sqb), nB, float(scale), _ptr(flag),
                                                _ptr(scale_dev), st), "lse2_spread_flag")
            with self._Timed(self, "lse"):
                check(self.lib.scb_lse2_pass(_ptr(A), nA, _ptr(Bm_all), nB, D, A.stride(0), Bm_all.stride(0), _DT[A.dtype],
                                             float(scale), jp, _ptr(pm), _ptr(pl), _ptr(cref), _ptr(csum), _ptr(scale_dev), st),
                      "lse2_pass")
            check(self.lib.scb_lse_combine(_ptr(pm), _ptr(pl), jp * nsub, nA, _ptr(r), st), "lse_combine")
            check(self.lib.scb_colstat_partial(_ptr(cref), _ptr(csum), n_strips, nB, _ptr(M), _ptr(L), st), "colstat_partial")
            with self._Timed(self, "lse"):
                check(self.lib.scb_lse_pass_cond(_ptr(Bm_rows), nR, _ptr(A_all), A_all.shape[0], D, Bm_rows.stride(0),
                                                 A_all.stride(0), _DT[A.dtype], float(scale), jp2, _ptr(pm2), _ptr(pl2),
                                                 _ptr(flag), _ptr(scale_dev), st), "lse_pass_cond")
            check(self.lib.scb_lse_combine_cond(_ptr(pm2), _ptr(pl2), jp2 * nsub2, nR, _ptr(c_exact), _ptr(flag), st),
                  "lse_combine_cond")
        self._count(6)
        return r, M, L, c_exact, flag

    def lse2_fold_ranks(self, pack_all, n_loc, off_exact, off_ref, off_sum, flag):
        """Column LSE of all world * n_loc columns from the gathered per-rank rows of `pack_all` [world, width]
        (see scb_lse2_fold_ranks): one launch instead of ~9 element-wise ones."""
        ws, width = pack_all.shape
        assert pack_all.dtype == torch.float32 and pack_all.is_contiguous()
        out = _empty(ws * n_loc, dtype=torch.float32, device=pack_all.device)
        with _On(pack_all.device):
            check(self.lib.scb_lse2_fold_ranks(_ptr(pack_all), ws, width, n_loc, off_exact, off_ref, off_sum, _ptr(flag),
                                               _ptr(out), self._stream()), "lse2_fold_ranks")
        self._count()
        return out

    def anchor_grad(self, A, Ball, V_rows, scale, row_lse, col_lse_all, col_lse_rows, diag, diag_off,
                    host_scale, dev_scale, want_ws, scale_dev=None):
        """dA = s * [ sum_j (P_ij + Q_ij) Ball_j  (j != diagonal)  +  (P_ii + Q_ii - 2) V_i ]  (fp32 [nA, D]);
        ws = sum_ij (P+Q)_ij (A_i . Ball_j) as a 0-dim tensor (for d/dtau) when want_ws."""
        nA, D = A.shape
        nB = Ball.shape[0]
        path = self.path_for(A, Ball)
        jp, nsub = self._plan(path, nA, nB, D, True, A.device)
        out = _empty(jp, nA, D, dtype=torch.float32, device=A.device)
        ws = _empty(jp * nsub, nA, dtype=torch.float32, device=A.device) if want_ws else None
        dA = _empty(nA, D, dtype=torch.float32, device=A.device)
        with _On(A.device):
            with self._Timed(self, "anchor_grad"):
                check(self.lib.scb_anchor_grad_pass(_ptr(A), nA, _ptr(Ball), nB, D, A.stride(0), Ball.stride(0),
                                                    _DT[A.dtype], float(scale), _ptr(row_lse), _ptr(col_lse_all),
                                                    int(diag_off), jp, _ptr(out), _ptr(ws), path, _ptr(scale_dev),
                                                    self._stream()), "anchor_grad_pass")
            self._count(2)
            check(self.lib.scb_anchor_grad_finalize(_ptr(out), jp, nA, D, _ptr(V_rows), V_rows.stride(0), _DT[V_rows.dtype],
                                                    _ptr(row_lse), _ptr(col_lse_rows), _ptr(diag), float(scale),
                                                    float(host_scale), _ptr(dev_scale), 0, _ptr(dA), _ptr(scale_dev),
                                                    self._stream()), "anchor_grad_finalize")
        return dA, (self.sum(ws) if want_ws else None)

    def anchor_grad_pass(self, A, Ball, scale, row_lse, col_lse_all, diag_off, want_ws, scale_dev=None):
        """The recompute sweep alone (no finaliser): {'out': [jparts, nA, D] fp32 partials, 'jparts', 'ws': 0-dim or None}."""
        nA, D = A.shape
        nB = Ball.shape[0]
        path = self.path_for(A, Ball)
        jp, nsub = self._plan(path, nA, nB, D, True, A.device)
        out = _empty(jp, nA, D, dtype=torch.float32, device=A.device)
        ws = _empty(jp * nsub, nA, dtype=torch.float32, device=A.device) if want_ws else None
        with _On(A.device):
            with self._Timed(self, "anchor_grad"):
                check(self.lib.scb_anchor_grad_pass(_ptr(A), nA, _ptr(Ball), nB, D, A.stride(0), Ball.stride(0),
                                                    _DT[A.dtype], float(scale), _ptr(row_lse), _ptr(col_lse_all),
                                                    int(diag_off), jp, _ptr(out), _ptr(ws), path, _ptr(scale_dev),
                                                    self._stream()), "anchor_grad_pass")
        self._count()
        return {"out": out, "jparts": jp, "ws": self.sum(ws) if want_ws else None}

    @staticmethod
    def can_fuse_normalize(X, E):
        """Whether grad_combine can apply the backward of the pre-loss normalise itself (a row is reduced inside one
        thread block: D <= 2048 with 16-byte aligned rows, see scb_grad_combine)."""
        D = X.shape[1]
        if E.stride(1) != 1 or E.dtype not in _DT:
            return False
        vec = (D % 8 == 0 and X.stride(0) % 8 == 0 and E.stride(0) % 8 == 0 and X.data_ptr() % 16 == 0
               and E.data_ptr() % 16 == 0)
        return D <= 256 or (vec and D <= 2048)

    def grad_combine(self, X, Y, out_dtype, anchor=None, unif=None, l_coef=0.0, dev_scale=None, extra=None, e_coef=0.0,
                     unit=None):
        """dX = a_coef (sum_p out[p] + dcoef Y) + u_coef (rq X - sum_p U[p]) + l_coef (X - Y) + e_coef extra, one
        streaming pass.  anchor = dict(out, jparts, row_lse, col_lse_rows, diag, scale, coef);
        unif = dict(core, coef, dev_coef); extra = contiguous fp32 [n, D].
        unit = (E, inv): X is normalize(E) with inv = 1 / ||E_i||; the pass then returns the gradient w.r.t. E (the
        backward of sparsify_clip.py:772-773 fused in; check can_fuse_normalize first)."""
        if extra is not None:
            assert extra.dtype == torch.float32 and extra.is_contiguous() and extra.shape == X.shape
        n, D = X.shape
        dX = _empty(n, D, dtype=out_dtype, device=X.device)
        a, u = anchor or {}, unif or {}
        core = u.get("core") or {}
        with _On(X.device):
            check(self.lib.scb_grad_combine(
                _ptr(X), _ptr(Y), n, D, X.stride(0), Y.stride(0) if Y is not None else 0, _DT[X.dtype],
                _ptr(a.get("out")), int(a.get("jparts", 0)), _ptr(a.get("row_lse")), _ptr(a.get("col_lse_rows")),
                _ptr(a.get("diag")), float(a.get("scale", 0.0)), float(a.get("coef", 0.0)),
                _ptr(core.get("U")), int(core.get("jparts", 0)), _ptr(core.get("rq")), int(core.get("nparts", 0)),
                float(u.get("coef", 0.0)), _ptr(u.get("dev_coef")), float(l_coef), _ptr(extra), float(e_coef),
                _ptr(dev_scale), _ptr(dX), _DT[out_dtype], dX.stride(0), _ptr(a.get("scale_dev")),
                _ptr(unit[0]) if unit else None, unit[0].stride(0) if unit else 0, _DT[unit[0].dtype] if unit else 0,
                _ptr(unit[1]) if unit else None, self._stream()),
                "grad_combine")
        self._count()
        return dX

    def loss_assemble(self, parts, c_anchor, two_scale, c_align, w_img, w_txt, w_cen, pair_norm, scale_dev=None):
        """(loss 0-dim, inv_ssum [3]) from the partial sums (scb_loss_assemble)."""
        loss = _empty((), dtype=torch.float32, device=parts.device)
        inv = _empty(3, dtype=torch.float32, device=parts.device)
        with _On(parts.device):
            check(self.lib.scb_loss_assemble(_ptr(parts), float(c_anchor), float(two_scale), float(c_align), float(w_img),
                                             float(w_txt), float(w_cen), float(pair_norm), _ptr(loss), _ptr(inv),
                                             _ptr(scale_dev), self._stream()), "loss_assemble")
        self._count()
        return loss, inv

    def lunif_core(self, Xr, Xall, t, row_offset, need_grad, sqn_r=None, sqn_all=None, sum_out=None):
        """One sweep over the pairwise Gaussian potentials of the rows of Xr against Xall.
        Returns {'rs_sum': 0-dim sum_i sum_{j != i} w_ij, and, when need_grad, 'U', 'rq', ...}."""
        nR, D = Xr.shape
        nAll = Xall.shape[0]
        path = self.path_for(Xr, Xall)
        jp, nsub = self._plan(path, nR, nAll, D, need_grad, Xr.device)
        if sqn_all is None:
            sqn_all = self.row_sqnorm(Xall)
        if sqn_r is None:  # by contract Xr holds rows [row_offset, row_offset + nR) of Xall
            sqn_r = sqn_all[row_offset:row_offset + nR]
        rs = _empty(jp * nsub, nR, dtype=torch.float32, device=Xr.device)
        core = {"jparts": jp, "nparts": jp * nsub, "path": path}
        with _On(Xr.device):
            if need_grad:
                U = _empty(jp, nR, D, dtype=torch.float32, device=Xr.device)
                rq = _empty(jp * nsub, nR, dtype=torch.float32, device=Xr.device)
                with self._Timed(self, "lunif"):
                    check(self.lib.scb_lunif_pass(_ptr(Xr), nR, _ptr(Xall), nAll, D, Xr.stride(0), Xall.stride(0),
                                                  _DT[Xr.dtype], float(t), _ptr(sqn_r), _ptr(sqn_all), int(row_offset), jp,
                                                  _ptr(U), _ptr(rq), _ptr(rs), path, self._stream()), "lunif_pass")
                self._count()
                core.update(U=U, rq=rq)
            else:
                check(self.lib.scb_lunif_sum_pass(_ptr(Xr), nR, _ptr(Xall), nAll, D, Xr.stride(0), Xall.stride(0),
                                                  _DT[Xr.dtype], float(t), _ptr(sqn_r), _ptr(sqn_all), int(row_offset), jp,
                                                  _ptr(rs), path, self._stream()), "lunif_sum_pass")
                self._count()
        core["rs_sum"] = self.sum(rs, out=sum_out)
        return core

    def lunif_grad(self, core, Xr, host_scale, dev_scale):
        """dX = s * (rq_i x_i - U_i), s = host_scale * dev_scale (0-dim device tensor)."""
        nR, D = Xr.shape
        dX = _empty(nR, D, dtype=torch.float32, device=Xr.device)
        with _On(Xr.device):
            check(self.lib.scb_lunif_grad_finalize(_ptr(core["U"]), core["jparts"], _ptr(core["rq"]), core["nparts"], nR, D,
                                                   _ptr(Xr), Xr.stride(0), _DT[Xr.dtype], float(host_scale), _ptr(dev_scale),
                                                   0, _ptr(dX), self._stream()), "lunif_grad_finalize")
        self._count()
        return dX

    def sparsify_sum(self, Xr, Xall, row_offset):
        nR, D = Xr.shape
        nAll = Xall.shape[0]
        path = self.path_for(Xr, Xall)
        jp, nsub = self._plan(path, nR, nAll, D, False, Xr.device)
        rs = _empty(jp * nsub, nR, dtype=torch.float32, device=Xr.device)
        with _On(Xr.device):
            check(self.lib.scb_sparsify_sum_pass(_ptr(Xr), nR, _ptr(Xall), nAll, D, Xr.stride(0), Xall.stride(0),
                                                 _DT[Xr.dtype], int(row_offset), jp, _ptr(rs), path, self._stream()),
                  "sparsify_sum_pass")
        self._count()
        return self.sum(rs)


    # ------------------------------------------------------------------ cold variants / evaluation-side consumers
    def col_sum(self, X, Y=None, scale=1.0):
        """scale * sum_i (X[i] - Y[i]) -> [D] fp32 (Y optional)."""
        n, D = X.shape
        rows = max(1, min(256, (n + 127) // 128))
        scratch = _empty(rows, D, dtype=torch.float32, device=X.device)
        out = _empty(D, dtype=torch.float32, device=X.device)
        with _On(X.device):
            check(self.lib.scb_col_sum(_ptr(X), _ptr(Y), n, D, X.stride(0), Y.stride(0) if Y is not None else 0, _DT[X.dtype],
                                       float(scale), _ptr(scratch), rows, _ptr(out), self._stream()), "col_sum")
        self._count(2)
        return out

    def gram_dd(self, X, mu=None, scale=1.0):
        """scale * sum_i (x_i - mu)(x_i - mu)^T -> [D, D] fp32."""
        n, D = X.shape
        parts = max(1, min(64, (n + 511) // 512))
        scratch = _empty(parts, D, D, dtype=torch.float32, device=X.device)
        out = _empty(D, D, dtype=torch.float32, device=X.device)
        with _On(X.device):
            check(self.lib.scb_gram_dd(_ptr(X), n, D, X.stride(0), _DT[X.dtype], _ptr(mu), float(scale), _ptr(scratch), parts,
                                       _ptr(out), self._stream()), "gram_dd")
        self._count(2)
        return out

    def rows_times_dd(self, X, M):
        n, D = X.shape
        assert M.dtype == torch.float32 and M.is_contiguous() and M.shape == (D, D)
        out = _empty(n, D, dtype=torch.float32, device=X.device)
        with _On(X.device):
            check(self.lib.scb_rows_times_dd(_ptr(X), n, D, X.stride(0), _DT[X.dtype], _ptr(M), _ptr(out), self._stream()),
                  "rows_times_dd")
        self._count()
        return out

    def sum_parts(self, parts, scale=1.0):
        """[P, n] fp32 -> [n]: fixed-order sum over the partial slots."""
        P, n = parts.shape
        out = _empty(n, dtype=torch.float32, device=parts.device)
        with _On(parts.device):
            check(self.lib.scb_sum_parts(_ptr(parts), P, n, float(scale), _ptr(out), self._stream()), "sum_parts")
        self._count()
        return out

    def rank_count(self, S, gt, columns=False, line=None):
        """Position of S[l, gt] in the descending sort of line l (row, or column when `columns`) -> int32 [len(gt)]."""
        assert S.dim() == 2 and S.stride(1) == 1 and gt.dtype == torch.int64
        n_r, n_c = S.shape
        nq = gt.numel()
        out = _empty(nq, dtype=torch.int32, device=S.device)
        sl, se, ne = (1, S.stride(0), n_r) if columns else (S.stride(0), 1, n_c)
        with _On(S.device):
            check(self.lib.scb_rank_count(_ptr(S), nq, ne, sl, se, _DT[S.dtype], _ptr(line), _ptr(gt), _ptr(out),
                                          self._stream()), "rank_count")
        self._count()
        return out

    def rank_count_pass(self, A, Bm, gt_score, diag_off=0):
        """#{j != i + diag_off : A_i . Bm_j > gt_score[i]} per row of A, from the features -> fp32 [nA] (exact integers)."""
        nA, D = A.shape
        nB = Bm.shape[0]
        path = self.path_for(A, Bm)
        jp, nsub = self._plan(path, nA, nB, D, False, A.device)
        cnt = _empty(jp * nsub, nA, dtype=torch.float32, device=A.device)
        with _On(A.device):
            check(self.lib.scb_rank_count_pass(_ptr(A), nA, _ptr(Bm), nB, D, A.stride(0), Bm.stride(0), _DT[A.dtype],
                                               _ptr(gt_score), int(diag_off), jp, _ptr(cnt), path, self._stream()),
                  "rank_count_pass")
        self._count()
        return self.sum_parts(cnt)


_backend = None


def get_backend():
    """The one and only compute backend.  Raises if the CUDA library is unavailable."""
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def set_backend(b):
    """Test hook (tests/ inject an oracle-backed double to exercise the sharding logic on CPU)."""
    global _backend
    prev, _backend = _backend, b
    return prev
