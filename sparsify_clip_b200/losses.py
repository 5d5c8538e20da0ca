"""Drop-in replacements for the loss functions of noostale/sparsify-clip.

Same names, positional order and defaults as the reference (sparsify_clip.py:110-187,
:308-355, :487-505), so a copy of its training loop runs unchanged with
``from sparsify_clip_b200 import *``.  Each function is a ``torch.autograd.Function`` over
the C-ABI CUDA library (include/scb200.h): the B x B similarity / distance matrices are
produced tile by tile on the tensor cores and never reach memory.

Every function takes an optional keyword ``group`` (a ``torch.distributed`` process group,
or ``True`` for the default group).  With a group, the inputs are this rank's ROW SHARD of
the global batch (equal shard sizes); the returned loss is the full-batch loss (identical on
all ranks) and the gradient is this rank's slice of the full-batch gradient.  Exchange steps:
all-gather of the [B, D] operands and of the 2 B LSE values, all-reduce of scalar partials.
No gradient reduce-scatter is needed: each rank recomputes its own row block of S for dI and
its own column block for dT.
"""

import torch
import torch.distributed as dist

from .backend_cuda import get_backend

__all__ = [
    "contrastive_loss", "lunif_loss", "lalign_loss", "compute_centroids_only", "compute_centroids",
    "sparsify_loss", "random_alignment_loss", "contrastive_loss_roberta", "centroid_alignment_loss",
    "normalized_centroids", "l2_normalize", "operand_dtype", "centroid_operand_dtype", "fused_terms_loss",
]


# ----------------------------------------------------------------------------- distributed helpers
def _resolve_group(group):
    if group is None or group is False:
        return None
    if group is True:
        return dist.group.WORLD
    return group


def _world(group):
    return (dist.get_rank(group), dist.get_world_size(group)) if group is not None else (0, 1)


def _all_gather_rows(x, group):
    if group is None:
        return x
    ws = dist.get_world_size(group)
    x = x.contiguous()
    out = torch.empty((ws * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out


def _all_gather_rows_async(x, group, role=None):
    """-> (gathered tensor, work handle); the caller waits on the handle before the first use.  With a `role` the gather
    goes through the SM-free peer pushes of peer.py when the ranks share a node (the result then lives in a per-role
    double buffer: valid until the second next gather of the same role); otherwise through NCCL."""
    if role is not None:
        from . import peer
        got = peer.all_gather_async(x, group, role)
        if got is not None:
            return got
    ws = dist.get_world_size(group)
    x = x.contiguous()
    out = torch.empty((ws * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out, dist.all_gather_into_tensor(out, x, group=group, async_op=True)


class _SmallReduce:
    """Asynchronous sum over the ranks of a few floats: gathered with the SM-free peer pushes (role-keyed double buffer) and
    added locally when the ranks share a node -- no NCCL kernel sits on the SMs while the sweeps run -- else NCCL."""

    def __init__(self, x, group, role):
        self.x, self.h, self.gathered = x, None, None
        got = None
        if x.is_cuda:
            from . import peer
            got = peer.all_gather_async(x.contiguous().view(1, -1), group, role)
        if got is not None:
            self.gathered, self.h = got
        else:
            self.h = dist.all_reduce(x, group=group, async_op=True)

    def result(self):
        self.h.wait()
        return self.x if self.gathered is None else self.gathered.sum(0)


def _all_reduce_(x, group):
    if group is not None:
        dist.all_reduce(x, group=group)
    return x


def _common(a, b):
    """Bring two operands to one dtype (the wider of the two)."""
    if a.dtype != b.dtype:
        dt = torch.promote_types(a.dtype, b.dtype)
        a, b = a.to(dt), b.to(dt)
    return a, b


def _scale_of(tau_t, tau_f):
    """-> (host scale, device multiplier or None).  A float or a CPU tensor (the reference's placement,
    sparsify_clip.py:716-717) gives the host float 1/tau.  A CUDA tensor is NOT read on the host: the kernels get scale = 1
    and a one-element device tensor 1/tau (no synchronisation; a captured CUDA graph follows the parameter)."""
    if tau_t is not None and tau_t.is_cuda:
        return 1.0, (1.0 / tau_t.detach().to(torch.float32)).reshape(1)
    return 1.0 / (float(tau_t) if tau_t is not None else float(tau_f)), None


def _gout32(g):
    return g.detach().to(torch.float32).reshape(()).contiguous()


# ----------------------------------------------------------------------------- anchor (CLIP InfoNCE)
class _AnchorFn(torch.autograd.Function):
    """sparsify_clip.py:110-132.  forward: ONE sweep over S that yields the row and the column LSE (sharded: the row
    LSE of the local rows of S and of S^T) + the diagonal; backward: one recompute sweep per operand (flash-style),
    d/dtau from the dI sweep."""

    @staticmethod
    def forward(ctx, I, T, tau_t, tau_f, group):
        be = get_backend()
        I, T = _common(I, T)
        Ip, Tp = be.prep(I), be.prep(T)
        scale, sdev = _scale_of(tau_t, tau_f)
        rank, ws = _world(group)
        n = Ip.shape[0]
        I_all, T_all = _all_gather_rows(Ip, group), _all_gather_rows(Tp, group)
        if group is None:
            r, c = be.lse_rows_cols(Ip, Tp, scale, scale_dev=sdev)    # row and column LSE from one sweep over S
        else:
            r = be.lse(Ip, T_all, scale, scale_dev=sdev)              # row LSE of the local rows of S
            c = be.lse(Tp, I_all, scale, scale_dev=sdev)              # column LSE of the local columns of S
        diag = be.row_dot(Ip, Tp)
        sfac = scale if sdev is None else sdev[0]
        part = be.sum(r) + be.sum(c) - (2.0 * sfac) * be.sum(diag)
        _all_reduce_(part, group)
        B = n * ws
        ctx.group, ctx.scale, ctx.B, ctx.off = group, scale, B, rank * n
        ctx.sdev = sdev
        ctx.in_dtypes = (I.dtype, T.dtype)
        ctx.tau_meta = None if tau_t is None else (tau_t.dtype, tau_t.device, tau_t.shape)
        ctx.save_for_backward(Ip, Tp, I_all, T_all, r, c, diag)
        return part / (2.0 * B)

    @staticmethod
    def backward(ctx, gout):
        be = get_backend()
        Ip, Tp, I_all, T_all, r, c, diag = ctx.saved_tensors
        group, scale, B, off = ctx.group, ctx.scale, ctx.B, ctx.off
        g = _gout32(gout)
        need_I, need_T, need_tau = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        r_all, c_all = _all_gather_rows(r, group), _all_gather_rows(c, group)
        coef = scale / (2.0 * B)
        sdev = ctx.sdev
        dI = dT = dtau = None
        ws_sum = None
        if need_I or need_tau:
            dI, ws_sum = be.anchor_grad(Ip, T_all, Tp, scale, r, c_all, c, diag, off, coef, g, need_tau, scale_dev=sdev)
            dI = dI.to(ctx.in_dtypes[0]) if need_I else None
        if need_T:
            dT, _ = be.anchor_grad(Tp, I_all, Ip, scale, c, r_all, r, diag, off, coef, g, False, scale_dev=sdev)
            dT = dT.to(ctx.in_dtypes[1])
        if need_tau:
            # d/dtau = -(1/tau) sum_ij G_ij S_ij,  G = (P + Q - 2 I_B)/(2B),  S = scale * (a.b)
            part = ws_sum - 2.0 * be.sum(diag)
            _all_reduce_(part, group)
            dt, dev, shp = ctx.tau_meta
            s2 = scale * scale if sdev is None else sdev[0] * sdev[0]
            dtau = (part * g * (-s2 / (2.0 * B))).to(device=dev, dtype=dt).reshape(shp)
        return dI, dT, dtau, None, None


def contrastive_loss(image_embeds, text_embeds, temperature=0.07, *, group=None):
    """CLIP anchor loss; `temperature` divides the logits (sparsify_clip.py:119-120) and may be a
    float or a 0-dim tensor / nn.Parameter on CPU or CUDA (its gradient comes back on its device)."""
    group = _resolve_group(group)
    if isinstance(temperature, torch.Tensor):
        return _AnchorFn.apply(image_embeds, text_embeds, temperature, None, group)
    return _AnchorFn.apply(image_embeds, text_embeds, None, float(temperature), group)


# ----------------------------------------------------------------------------- L_unif
class _LunifFn(torch.autograd.Function):
    """sparsify_clip.py:159-164.  When a gradient is needed the forward is a SINGLE sweep that
    yields the loss and the unscaled gradient together (W.X on the tensor cores right behind the
    Gram tile); backward is one element-wise multiply by grad_output."""

    @staticmethod
    def forward(ctx, x, t, group, mma_dtype):
        be = get_backend()
        xp = be.prep(x, cast_fp32=mma_dtype is None)
        # `mma_dtype`: dtype of the tensor-core operand when x is a higher-precision intermediate (the
        # normalised centroids).  Distances are exact for the ROUNDED points; the x_i (sum_j w_ij) term of
        # the gradient uses the unrounded rows, so the large radial component cancels exactly downstream.
        xq = xp.to(mma_dtype) if (mma_dtype is not None and xp.dtype != mma_dtype) else xp
        rank, ws = _world(group)
        n = xp.shape[0]
        x_all = _all_gather_rows(xq, group)
        need = ctx.needs_input_grad[0]
        core = be.lunif_core(xq, x_all, float(t), rank * n, need)
        rs = _all_reduce_(core["rs_sum"], group)
        B = n * ws
        ssum = rs * 0.5
        loss = torch.log(ssum / (B * (B - 1) / 2.0))          # B == 1 -> log(0/0) = nan, as the reference
        if need:
            inv = (-2.0 * float(t)) / ssum
            gx = be.lunif_grad(core, xp, 1.0, inv.contiguous())
            ctx.save_for_backward(gx)
            ctx.in_dtype = x.dtype
        return loss

    @staticmethod
    def backward(ctx, gout):
        (gx,) = ctx.saved_tensors
        return (gx * _gout32(gout)).to(ctx.in_dtype), None, None, None


def lunif_loss(x, t=2, *, group=None, mma_dtype=None):
    return _LunifFn.apply(x, t, _resolve_group(group), mma_dtype)


# ----------------------------------------------------------------------------- fused composition
def _flatten(obj):
    """Nested tuples / dicts with tensor leaves -> (list of tensors, spec) for ctx.save_for_backward."""
    flat = []

    def walk(o):
        if isinstance(o, torch.Tensor):
            flat.append(o)
            return ("t", len(flat) - 1)
        if isinstance(o, dict):
            return ("d", {k: walk(v) for k, v in o.items()})
        if isinstance(o, (tuple, list)):
            return ("l", [walk(v) for v in o])
        return ("c", o)

    return flat, walk(obj)


def _unflatten(tensors, spec):
    kind, val = spec
    if kind == "t":
        return tensors[val]
    if kind == "d":
        return {k: _unflatten(tensors, v) for k, v in val.items()}
    if kind == "l":
        return tuple(_unflatten(tensors, v) for v in val)
    return val


class _FusedTermsFn(torch.autograd.Function):
    """w_a * contrastive_loss(I, T, tau) + w_l * lalign_loss(I, T) + w_i * lunif_loss(I) + w_t * lunif_loss(T)
    + w_c * lunif_loss(normalize((I + T) / 2))  (every composition of sparsify_clip.py:778-938) as ONE autograd node.

    Same B x B passes as the separate functions; what is fused is everything around them: the per-term gradient
    finalisers, the dtype casts and autograd's accumulation of the terms collapse into one streaming pass per
    operand (scb_grad_combine) that writes dI / dT directly in the input dtype.  The gradient is produced in
    forward (L_unif yields it in the same sweep as its value anyway); backward multiplies by grad_output."""

    @staticmethod
    def forward(ctx, I, T, tau_t, tau_f, w_a, w_l, w_i, w_t, w_c, t_unif, group, normalize=False):
        be = get_backend()
        I, T = _common(I, T)
        E_I = E_T = inv_I = inv_T = None
        if normalize:
            # the pre-loss normalise of the training loop (sparsify_clip.py:772-773, no eps) inside the node: the terms are
            # evaluated on e / ||e|| (written in the operand dtype from the un-rounded rows), and the combine pass of
            # backward pulls the gradient back to e itself (scb_grad_combine, unit_src / unit_inv)
            E_I, E_T = be.prep(I, cast_fp32=False), be.prep(T, cast_fp32=False)
            op_dt = be.prep(I[:1]).dtype
            (Ip, inv_I), (Tp, inv_T) = be.normalize_fwd(E_I, op_dt), be.normalize_fwd(E_T, op_dt)
        else:
            Ip, Tp = be.prep(I), be.prep(T)
        rank, ws = _world(group)
        n, D = Ip.shape
        B = n * ws
        off = rank * n
        dev = Ip.device
        need_I, need_T, need_tau = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need = need_I or need_T
        # ---- exchange step 1: the operands.  Every gather is asynchronous (NCCL's own stream) and is awaited right
        # before the first sweep that reads its result, so the later gathers run UNDER the earlier sweeps: the gradient
        # sweeps occupy clusters of 4 SMs (132 of a B200's 148 SMs), which leaves room for the small NCCL kernels.
        I_all, T_all = Ip, Tp
        hI = hT = hC = None
        # normalised centroids (sparsify_clip.py:353 + :804): fp32 master copy C, tensor-core operand Cq (see
        # centroid_operand_dtype: fp16 keeps the term inside the parity gates where bf16 does not)
        C = Cq = C_all = c_inv = None
        if w_c != 0.0:
            C, c_inv = be.centroid_fwd(Ip, Tp, torch.float32)
            cdt = torch.float16 if Ip.dtype in (torch.bfloat16, torch.float16) else Ip.dtype
            Cq = C.to(cdt) if C.dtype != cdt else C
            C_all = Cq
        if group is not None:
            if w_a != 0.0 or w_i != 0.0:
                I_all, hI = _all_gather_rows_async(Ip, group, "I")
            if w_a != 0.0 or w_t != 0.0:
                T_all, hT = _all_gather_rows_async(Tp, group, "T")
            if w_c != 0.0:
                C_all, hC = _all_gather_rows_async(Cq, group, "C")

        def _await(h):
            if h is not None:
                h.wait()
            return None

        # scalar partial sums of this rank, written in place by the kernels that produce them:
        #   0: sum r   1: sum c   2: sum diag   3: L_align   4: L_unif row sum (I)   |   5, 6: L_unif row sums (T, centroids)
        #   7: d/dtau.   Sharded: 0..4 travel in the packed gather (exchange step 2), 5..7 in the late all-reduce (step 3).
        NS, NP = 7, 5
        parts = torch.zeros(NS + 1, dtype=torch.float32, device=dev)
        if w_l != 0.0:
            be.sum(be.lalign_rows(Ip, Tp), out=parts[3:4])
        cores = {}

        def _unif(wu, Xp, X_all, needx, which, slot):
            if wu != 0.0:
                core = be.lunif_core(Xp, X_all, float(t_unif), off, needx, sum_out=parts[slot:slot + 1])
                cores[which] = (core, wu, needx, slot - 4)

        # L_unif(I) needs only the gathered I: it runs while T (and the centroids) are still in flight
        hI = _await(hI)
        _unif(w_i, Ip, I_all, need_I, "I", 4)
        an_I = an_T = None
        scale, sdev = 0.0, None
        r = c = None
        colparts = None
        if w_a != 0.0:
            hT = _await(hT)
            scale, sdev = _scale_of(tau_t, tau_f)
            if group is None:
                r, c = be.lse_rows_cols(Ip, Tp, scale, scale_dev=sdev)      # both from one sweep over S
            else:
                colparts = be.lse_rows_colparts(Ip, T_all, Tp, I_all, scale, scale_dev=sdev)
                if colparts is None:                        # not on the tensor-core path: two sweeps
                    r = be.lse(Ip, T_all, scale, scale_dev=sdev)
                    c = be.lse(Tp, I_all, scale, scale_dev=sdev)
                else:                                       # one sweep: my rows of S; the column sums are folded
                    r = colparts[0]                         # over the ranks after the gather below
            diag = be.row_dot(Ip, Tp)
            be.sum(r, out=parts[0:1])
            be.sum(diag, out=parts[2:3])
            if colparts is None:
                be.sum(c, out=parts[1:2])
        # ---- exchange step 2: ONE small gather carries r, c (or the column partials) and the scalar partials 0..4 of
        # every rank; it is issued here and awaited after the remaining L_unif sweeps, i.e. it runs under them.
        r_all, c_all = r, c
        gathered = False
        hP = pack_flat = None
        if group is not None and w_a != 0.0 and (need or need_tau or colparts is not None):
            if colparts is None:
                pieces = (r, c, parts[:NP])
            else:
                _, cM, cL, c_exact, flag = colparts
                pieces = (r, c_exact, parts[:NP], cM, cL)
            pack = torch.cat([x.to(torch.float32) for x in pieces])
            pack_flat, hP = _all_gather_rows_async(pack, group, "P")
            gathered = True
        hT = _await(hT)
        _unif(w_t, Tp, T_all, need_T, "T", 5)
        hC = _await(hC)
        _unif(w_c, Cq, C_all, need, "C", 6)
        local_sdiag = parts[2].clone() if (gathered and need_tau) else parts[2]
        if gathered:
            hP = _await(hP)
            pack_all = pack_flat.view(ws, -1)
            r_all = pack_all[:, :n].reshape(-1)
            parts = torch.cat((pack_all[:, 2 * n:2 * n + NP].sum(0), parts[NP:]))
            if colparts is None:
                c_all = pack_all[:, n:2 * n].reshape(-1)
            else:
                # fold the column partials of all ranks (or take the exact second sweep where the bound demanded it)
                c_all = be.lse2_fold_ranks(pack_all, n, n, 2 * n + NP, 2 * n + NP + B, flag)
                c = c_all[off:off + n]
                be.sum(c_all, out=parts[1:2])
        # ---- exchange step 3 (what step 2 could not carry: the L_unif sums produced under it, later d/dtau; everything
        # when there is no anchor term).  L_unif-only compositions reduce here, before the gradient coefficients are formed.
        hR = hD = None
        late_lo = NP if gathered else 0
        if group is not None and not (w_a != 0.0 and (need_I or need_tau or need_T)):
            if (not gathered) or w_t != 0.0 or w_c != 0.0:
                dist.all_reduce(parts[late_lo:NS], group=group)
        if w_a != 0.0 and (need or need_tau):
            coef = w_a * scale / (2.0 * B)
            if group is not None and (not gathered or w_t != 0.0 or w_c != 0.0):
                hR = _SmallReduce(parts[late_lo:NS].clone(), group, "R")      # reduced under the anchor-gradient sweeps
            if need_I or need_tau:
                p = be.anchor_grad_pass(Ip, T_all, scale, r, c_all, off, need_tau, scale_dev=sdev)
                an_I = dict(out=p["out"], jparts=p["jparts"], row_lse=r, col_lse_rows=c, diag=diag, scale=scale, coef=coef,
                            scale_dev=sdev)
                if need_tau:
                    parts[NS] = p["ws"] - 2.0 * local_sdiag
                    if group is not None:                    # reduced under the dT sweep
                        hD = _SmallReduce(parts[NS:].clone(), group, "D")
            if need_T:
                p = be.anchor_grad_pass(Tp, I_all, scale, c, r_all, off, False, scale_dev=sdev)
                an_T = dict(out=p["out"], jparts=p["jparts"], row_lse=c, col_lse_rows=r, diag=diag, scale=scale, coef=coef,
                            scale_dev=sdev)
            if hR is not None:
                parts[late_lo:NS] = hR.result()
            if hD is not None:
                parts[NS:] = hD.result()
        if group is not None and Ip.is_cuda:
            from . import peer
            peer.release_all()          # every sweep that reads a gathered buffer is enqueued: the peers may refill them
        # the additions of the ladder on the (now global) partial sums, and 1 / Ssum of every L_unif term, in one launch
        loss, inv_ssum = be.loss_assemble(parts, w_a / (2.0 * B), 2.0 * scale, w_l / B, w_i, w_t, w_c, B * (B - 1) / 2.0,
                                          scale_dev=sdev)
        cen = un_I = un_T = None
        for which, (core, wu, needx, k) in cores.items():
            if needx:
                u = dict(core=core, coef=wu * (-2.0 * float(t_unif)), dev_coef=inv_ssum[k:k + 1])
                if which == "I":
                    un_I = u
                elif which == "T":
                    un_T = u
                else:
                    # the centroid chain: L_unif gradient at the unrounded centroids, pulled back through
                    # normalize((I + T) / 2); the result is the same array for both operands
                    gC = be.lunif_grad(core, C, u["coef"], u["dev_coef"])
                    cen = be.centroid_bwd(Ip, Tp, gC, c_inv, both=False)[0]
        # d/dtau stays on the GPU here.  A temperature parameter that lives on the CPU (the reference's placement,
        # sparsify_clip.py:716-717) needs a device-to-host copy of this scalar, i.e. a host synchronisation: it happens in
        # backward AFTER the combine kernels are enqueued, so the device never waits for the host to come back from it
        # (done here, it cost ~0.7 ms of idle GPU per step between the last sweep and the combine passes).
        dtau = None
        ctx.tau_meta = None
        if need_tau and w_a != 0.0:
            ctx.tau_meta = (tau_t.dtype, tau_t.device, tau_t.shape)
            s2 = scale * scale if sdev is None else sdev[0] * sdev[0]
            dtau = parts[NS] * (s2 * (-w_a / (2.0 * B)))
        # The per-operand combine runs in backward with grad_output as its device-side scale: one pass writes the final
        # gradient in the input dtype (no separate multiply).  The sweep outputs stay alive until then.
        okdt = (torch.float32, torch.bfloat16, torch.float16)
        ctx.in_dtypes = (I.dtype, T.dtype)
        ctx.out_dtypes = (I.dtype if I.dtype in okdt else torch.float32, T.dtype if T.dtype in okdt else torch.float32)
        ctx.need = (need_I, need_T)
        # every tensor the backward reads goes through save_for_backward (autograd's version counters then catch an
        # in-place update of the embeddings between forward and backward); the dict structure is rebuilt from a spec
        flat, ctx.spec = _flatten((Ip, Tp, an_I, an_T, un_I, un_T, 2.0 * w_l / B, cen, dtau, E_I, E_T, inv_I, inv_T))
        ctx.save_for_backward(*flat)
        return loss

    @staticmethod
    def backward(ctx, gout):
        be = get_backend()
        g = _gout32(gout)
        Ip, Tp, an_I, an_T, un_I, un_T, lc, cen, dtau0, E_I, E_T, inv_I, inv_T = _unflatten(ctx.saved_tensors, ctx.spec)
        dI = dT = dtau = None

        def combine(X, Y, k, an, un, E, inv):
            if E is None:
                return be.grad_combine(X, Y, ctx.out_dtypes[k], anchor=an, unif=un, l_coef=lc, dev_scale=g, extra=cen,
                                       e_coef=1.0).to(ctx.in_dtypes[k])
            if be.can_fuse_normalize(X, E):          # one pass: every term, then the backward of e -> e / ||e||
                return be.grad_combine(X, Y, ctx.out_dtypes[k], anchor=an, unif=un, l_coef=lc, dev_scale=g, extra=cen,
                                       e_coef=1.0, unit=(E, inv)).to(ctx.in_dtypes[k])
            g32 = be.grad_combine(X, Y, torch.float32, anchor=an, unif=un, l_coef=lc, dev_scale=g, extra=cen, e_coef=1.0)
            return be.normalize_bwd(E, g32, inv).to(ctx.in_dtypes[k])

        if ctx.need[0]:
            dI = combine(Ip, Tp, 0, an_I, un_I, E_I, inv_I)
        if ctx.need[1]:
            dT = combine(Tp, Ip, 1, an_T, un_T, E_T, inv_T)
        if dtau0 is not None:
            dt, tdev, shp = ctx.tau_meta
            dtau = (dtau0 * g.to(device=dtau0.device, dtype=dtau0.dtype)).to(device=tdev, dtype=dt).reshape(shp)
        return dI, dT, dtau, None, None, None, None, None, None, None, None, None


def fused_terms_loss(image_embeds, text_embeds, temperature=0.07, w_anchor=1.0, w_align=0.0, w_unif_img=0.0,
                     w_unif_txt=0.0, t=2, *, w_unif_cen=0.0, group=None, normalize=False):
    """w_anchor * contrastive_loss + w_align * lalign_loss + w_unif_img * lunif_loss(I) + w_unif_txt * lunif_loss(T)
    + w_unif_cen * lunif_loss(normalized_centroids(I, T)), evaluated as one fused autograd node (see _FusedTermsFn).
    Zero weights skip their kernels.  normalize=True: the inputs are the encoders' un-normalised outputs; the node
    applies the training loop's pre-loss normalise (sparsify_clip.py:772-773) and its backward itself."""
    group = _resolve_group(group)
    tt = temperature if isinstance(temperature, torch.Tensor) else None
    tf = None if tt is not None else float(temperature)
    return _FusedTermsFn.apply(image_embeds, text_embeds, tt, tf, float(w_anchor), float(w_align), float(w_unif_img),
                               float(w_unif_txt), float(w_unif_cen), t, group, bool(normalize))


# ----------------------------------------------------------------------------- L_align
class _LalignFn(torch.autograd.Function):
    """sparsify_clip.py:186-187 with alpha = 2."""

    @staticmethod
    def forward(ctx, x, y, group):
        be = get_backend()
        x, y = _common(x, y)
        xp, yp = be.prep(x), be.prep(y)
        _, ws = _world(group)
        part = _all_reduce_(be.sum(be.lalign_rows(xp, yp)), group)
        ctx.B = xp.shape[0] * ws
        ctx.in_dtypes = (x.dtype, y.dtype)
        ctx.save_for_backward(xp, yp)
        return part / ctx.B

    @staticmethod
    def backward(ctx, gout):
        be = get_backend()
        xp, yp = ctx.saved_tensors
        dX, dY = be.lalign_bwd(xp, yp, 2.0 / ctx.B, _gout32(gout), ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return (None if dX is None else dX.to(ctx.in_dtypes[0]),
                None if dY is None else dY.to(ctx.in_dtypes[1]), None)


def lalign_loss(x, y, alpha=2, *, group=None):
    if alpha != 2:
        # the reference only ever calls alpha=2 (all YAMLs); other exponents stay in PyTorch
        if _resolve_group(group) is not None:
            raise NotImplementedError("lalign_loss(alpha != 2) has no sharded form; only alpha = 2 is on the kernel path")
        return (x - y).norm(dim=1).pow(alpha).mean()
    return _LalignFn.apply(x, y, _resolve_group(group))


# ----------------------------------------------------------------------------- centroids / normalise
class _CentroidFn(torch.autograd.Function):
    """normalize((a + b)/2, eps=1e-12): sparsify_clip.py:353 + F.normalize at :804, fused."""

    @staticmethod
    def forward(ctx, a, b):
        be = get_backend()
        a, b = _common(a, b)
        ap, bp = be.prep(a), be.prep(b)
        C, inv = be.centroid_fwd(ap, bp, torch.float32)     # fp32 master copy (see lunif_loss(mma_dtype=...))
        ctx.in_dtypes = (a.dtype, b.dtype)
        ctx.save_for_backward(ap, bp, inv)
        return C

    @staticmethod
    def backward(ctx, dC):
        be = get_backend()
        ap, bp, inv = ctx.saved_tensors
        dA, dB = be.centroid_bwd(ap, bp, dC.detach().to(torch.float32).contiguous(), inv)
        return dA.to(ctx.in_dtypes[0]), dB.to(ctx.in_dtypes[1])


def normalized_centroids(a, b):
    """F.normalize(compute_centroids_only(a, b), dim=-1) in one kernel (forward and backward).
    Returns fp32; pass ``mma_dtype=operand_dtype(a)`` to lunif_loss to keep it on the tensor-core path."""
    return _CentroidFn.apply(a, b)


def centroid_operand_dtype(x):
    """Tensor-core operand dtype for normalize((I+T)/2) given what I is computed in.  The centroid is an
    fp32 intermediate, not an input: when I/T run on the tensor cores it is handed over as fp16 (11-bit
    mantissa; components of a unit vector are far inside fp16's range), which keeps L_unif(centroids)
    within the 1e-5 / 1e-3 parity gates where a bf16 operand (8 bits) does not."""
    d = operand_dtype(x)
    return torch.float16 if d in (torch.bfloat16, torch.float16) else d


def operand_dtype(x):
    """dtype the B x B passes will use for `x` (bf16/fp16 stay; fp32 -> bf16 under set_fp32_mode('bf16'))."""
    return get_backend().prep(x[:1]).dtype


class _NormalizeFn(torch.autograd.Function):
    """e / e.norm(dim=-1, keepdim=True)  (no eps) -- sparsify_clip.py:772-773."""

    @staticmethod
    def forward(ctx, x):
        be = get_backend()
        xp = be.prep(x)
        Y, inv = be.normalize_fwd(xp, xp.dtype)
        ctx.in_dtype = x.dtype
        ctx.save_for_backward(xp, inv)
        return Y

    @staticmethod
    def backward(ctx, dY):
        be = get_backend()
        xp, inv = ctx.saved_tensors
        return be.normalize_bwd(xp, dY.detach().to(torch.float32).contiguous(), inv).to(ctx.in_dtype)


def l2_normalize(x):
    return _NormalizeFn.apply(x)


def compute_centroids_only(text_embeddings, visual_embeddings):
    """(a + b) / 2 -- sparsify_clip.py:334-355 (the call sites pass (image, text); symmetric)."""
    return (text_embeddings + visual_embeddings) / 2.0


def compute_centroids(text_embeddings, visual_embeddings):
    """All-pairs centroids [B1, B2, D] and their norms -- sparsify_clip.py:308-332 (never called by
    the reference; kept for the signature surface, O(B1*B2*D) memory by definition)."""
    centroids = (text_embeddings.unsqueeze(1) + visual_embeddings.unsqueeze(0)) / 2.0
    return torch.norm(centroids, dim=-1), centroids


# ----------------------------------------------------------------------------- cold variants
class _SparsifyFn(torch.autograd.Function):
    """mse(x x^T, 2 eye - 1) -- sparsify_clip.py:166-176.  Forward on the Gram-tile sweep; backward without a second
    B x B pass: with E = X X^T + 1 1^T - 2 I,  dX = (4 / B^2) E X = (4 / B^2) (X (X^T X) + 1 (sum_j x_j)^T - 2 X), i.e. one
    D x D second moment (scb_gram_dd) and one [B, D] x [D, D] product (scb_rows_times_dd) of the SAME prepared operand."""

    @staticmethod
    def forward(ctx, x):
        be = get_backend()
        xp = be.prep(x)
        B = xp.shape[0]
        ctx.save_for_backward(xp)
        ctx.in_dtype = x.dtype
        return be.sparsify_sum(xp, xp, 0) / float(B * B)

    @staticmethod
    def backward(ctx, gout):
        be = get_backend()
        (xp,) = ctx.saved_tensors
        B = xp.shape[0]
        XM = be.rows_times_dd(xp, be.gram_dd(xp))
        s = be.col_sum(xp)
        g = (XM + s[None, :] - 2.0 * xp.float()) * ((4.0 / (B * B)) * _gout32(gout))
        return g.to(ctx.in_dtype)


def sparsify_loss(x):
    return _SparsifyFn.apply(x)


def random_alignment_loss(x, y):
    """sparsify_clip.py:178-184: L_align against a random permutation of y."""
    idx = torch.randperm(y.size(0))          # CPU generator, like the reference (:181): seeds reproduce its stream
    return lalign_loss(x, y[idx.to(y.device)])


def contrastive_loss_roberta(image_embeds, text_embeds, roberta_similarity, temperature=0.07):
    """Soft-target variant, sparsify_clip.py:135-157.  Dead code in the reference (only referenced
    from a string literal); it needs a dense B x B target matrix as INPUT, so it stays in PyTorch."""
    logits = image_embeds @ text_embeds.t() / temperature
    li = torch.nn.functional.cross_entropy(logits, roberta_similarity)
    lt = torch.nn.functional.cross_entropy(logits.t(), roberta_similarity.t())
    return (li + lt) / 2


class _CentroidAlignFn(torch.autograd.Function):
    """|| mean(img) - mean(txt) ||_p -- sparsify_clip.py:487-505: one column-sum kernel over the rows (scb_col_sum); the
    p-norm of the D-vector and its gradient sign(d) |d|^(p-1) / |d|_p^(p-1) / B are O(D)."""

    @staticmethod
    def forward(ctx, a, b, p):
        be = get_backend()
        a, b = _common(a, b)
        ap, bp = be.prep(a), be.prep(b)
        n = ap.shape[0]
        d = be.col_sum(ap, bp, 1.0 / n)
        nrm = torch.linalg.vector_norm(d, ord=p)
        ctx.save_for_backward(d, nrm)
        ctx.meta = (p, n, a.shape, a.dtype, b.dtype)
        return nrm

    @staticmethod
    def backward(ctx, gout):
        d, nrm = ctx.saved_tensors
        p, n, shape, dta, dtb = ctx.meta
        if p == 2:
            gd = d / nrm
        else:
            gd = torch.sign(d) * d.abs().pow(p - 1) / nrm.pow(p - 1)
        row = (gd * (_gout32(gout) / n))[None, :].expand(shape)
        return row.to(dta), (-row).to(dtb), None


def centroid_alignment_loss(img_embeds, txt_embeds, p=2):
    """|| mean(img) - mean(txt) ||_p -- sparsify_clip.py:487-505 (unused by the YAMLs).  CUDA tensors run on the
    column-sum kernel; host tensors (the reference's own formula, for CPU-side checks) stay in PyTorch."""
    if img_embeds.is_cuda and img_embeds.dim() == 2:
        return _CentroidAlignFn.apply(img_embeds, txt_embeds, p)
    return torch.norm(img_embeds.mean(dim=0) - txt_embeds.mean(dim=0), p=p)
