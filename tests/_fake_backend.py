"""TEST DOUBLE for sparsify_clip_b200.backend_cuda.CudaBackend: the same interface implemented with
dense torch fp64 CPU math (oracle formulas).  Lets the CPU suite exercise the host logic of
losses.py -- sharding offsets, all-gathers, scalar assembly, autograd wiring -- under gloo.
Never imported by the product."""
import math

import torch


class FakeBackend:
    name = "fake-cpu"
    launches = 0
    pass_events = None

    def prep(self, x, cast_fp32=True):
        return x.detach().double().contiguous()

    def path_for(self, *mats):
        return 0

    def sum(self, x, out=None):
        v = x.double().sum()
        if out is not None:
            out.copy_(v.reshape(out.shape))
            return out
        return v

    def loss_assemble(self, parts, c_anchor, two_scale, c_align, w_img, w_txt, w_cen, pair_norm, scale_dev=None):
        p = parts.double()
        loss = torch.zeros((), dtype=torch.float64)
        if c_anchor != 0.0:
            loss = loss + c_anchor * (p[0] + p[1] - two_scale * p[2])
        if c_align != 0.0:
            loss = loss + c_align * p[3]
        inv = torch.zeros(3, dtype=torch.float64)
        for k, w in enumerate((w_img, w_txt, w_cen)):
            if w != 0.0:
                ss = 0.5 * p[4 + k]
                loss = loss + w * torch.log(ss / pair_norm)
                inv[k] = 1.0 / ss
        return loss, inv

    def row_sqnorm(self, x):
        return (x * x).sum(1)

    def row_dot(self, a, b):
        return (a * b).sum(1)

    def lalign_rows(self, x, y):
        return ((x - y) ** 2).sum(1)

    def lalign_bwd(self, x, y, host_scale, dev_scale, want_x=True, want_y=True):
        g = host_scale * dev_scale.double() * (x - y)
        return (g if want_x else None), (-g if want_y else None)

    def centroid_fwd(self, a, b, out_dtype):
        m = (a + b) / 2
        inv = 1.0 / m.norm(dim=1).clamp_min(1e-12)
        return m * inv[:, None], inv

    def centroid_bwd(self, a, b, dC, inv, host_scale=1.0, dev_scale=None, both=True):
        c = (a + b) / 2 * inv[:, None]
        dC = dC.double()
        dm = (dC - c * (c * dC).sum(1, keepdim=True)) * inv[:, None] * 0.5 * host_scale
        return dm, (dm.clone() if both else None)

    def normalize_fwd(self, x, out_dtype):
        inv = 1.0 / x.norm(dim=1)                                # no eps, as sparsify_clip.py:772-773
        return (x * inv[:, None]).to(out_dtype), inv

    def normalize_bwd(self, x, dY, inv):
        y = x * inv[:, None]
        dY = dY.to(x.dtype)
        return (dY - y * (y * dY).sum(1, keepdim=True)) * inv[:, None]

    def lse2_fold_ranks(self, pack_all, n_loc, off_exact, off_ref, off_sum, flag):
        B = pack_all.shape[0] * n_loc
        if int(flag.item()) != 0:
            return pack_all[:, off_exact:off_exact + n_loc].reshape(-1)
        Mr, Lr = pack_all[:, off_ref:off_ref + B], pack_all[:, off_sum:off_sum + B]
        Mx = Mr.max(0).values
        return (Mx + torch.log2((Lr * torch.exp2(Mr - Mx)).sum(0))) * math.log(2.0)

    def lse(self, A, Ball, scale, scale_dev=None):
        return torch.logsumexp(scale * A @ Ball.t(), dim=1)

    def lse_rows_cols(self, A, Bm, scale, scale_dev=None):
        S = scale * A @ Bm.t()
        return torch.logsumexp(S, dim=1), torch.logsumexp(S, dim=0)

    def lse_rows_colparts(self, A, Bm_all, Bm_rows, A_all, scale, scale_dev=None):
        S = (scale * A @ Bm_all.t()) / 0.6931471805599453        # log2 domain, as the kernels
        M = S.max(0).values
        L = torch.exp2(S - M).sum(0)
        r = torch.logsumexp(scale * A @ Bm_all.t(), dim=1)
        return r, M, L, torch.zeros(Bm_rows.shape[0], dtype=A.dtype), torch.zeros(1, dtype=torch.int32)

    def anchor_grad(self, A, Ball, V_rows, scale, row_lse, col_lse_all, col_lse_rows, diag, diag_off, host_scale,
                    dev_scale, want_ws, scale_dev=None):
        G0 = A @ Ball.t()
        S = scale * G0
        W = torch.exp(S - row_lse[:, None]) + torch.exp(S - col_lse_all[None, :])
        ws = (W * G0).sum() if want_ws else None
        idx = torch.arange(A.shape[0])
        Wd = W.clone()
        Wd[idx, idx + diag_off] = 0
        sii = scale * diag
        dcoef = torch.exp(sii - row_lse) + torch.exp(sii - col_lse_rows) - 2
        dA = host_scale * dev_scale.double() * (Wd @ Ball + dcoef[:, None] * V_rows)
        return dA, ws

    def anchor_grad_pass(self, A, Ball, scale, row_lse, col_lse_all, diag_off, want_ws, scale_dev=None):
        G0 = A @ Ball.t()
        S = scale * G0
        W = torch.exp(S - row_lse[:, None]) + torch.exp(S - col_lse_all[None, :])
        ws = (W * G0).sum() if want_ws else None
        idx = torch.arange(A.shape[0])
        W[idx, idx + diag_off] = 0
        return {"out": (W @ Ball)[None], "jparts": 1, "ws": ws}

    @staticmethod
    def can_fuse_normalize(X, E):
        return X.shape[1] % 8 == 0          # exercise both routes of the fused node's backward

    def grad_combine(self, X, Y, out_dtype, anchor=None, unif=None, l_coef=0.0, dev_scale=None, extra=None, e_coef=0.0,
                     unit=None):
        g = l_coef * (X - Y) if l_coef != 0.0 else torch.zeros_like(X)
        if extra is not None:
            g = g + e_coef * extra
        if anchor:
            sii = anchor["scale"] * anchor["diag"]
            dcoef = torch.exp(sii - anchor["row_lse"]) + torch.exp(sii - anchor["col_lse_rows"]) - 2
            g = g + anchor["coef"] * (anchor["out"].sum(0) + dcoef[:, None] * Y)
        if unif:
            core = unif["core"]
            U = core["U"] if core["U"].dim() == 2 else core["U"].sum(0)
            rq = core["rq"] if core["rq"].dim() == 1 else core["rq"].sum(0)
            uc = unif["coef"] * (unif["dev_coef"].double() if unif.get("dev_coef") is not None else 1.0)
            g = g + uc * (rq[:, None] * X - U)
        if dev_scale is not None:
            g = g * dev_scale.double()
        if unit is not None:                 # dE = (g - e (e . g)) / ||E||
            E, inv = unit
            e = E * inv[:, None]
            g = (g - e * (e * g).sum(1, keepdim=True)) * inv[:, None]
        return g.to(out_dtype)

    def lunif_core(self, Xr, Xall, t, row_offset, need_grad, sqn_r=None, sqn_all=None, sum_out=None):
        n = (Xall * Xall).sum(1)
        nr = (Xr * Xr).sum(1)
        d2 = (nr[:, None] + n[None, :] - 2 * Xr @ Xall.t()).clamp_min(0)
        W = torch.exp(-t * d2)
        idx = torch.arange(Xr.shape[0])
        W[idx, idx + row_offset] = 0
        core = {"rs_sum": W.sum()}
        if sum_out is not None:
            sum_out.copy_(core["rs_sum"].reshape(sum_out.shape))
        if need_grad:
            core.update(U=W @ Xall, rq=W.sum(1))
        return core

    def lunif_grad(self, core, Xr, host_scale, dev_scale):
        return host_scale * dev_scale.double() * (core["rq"][:, None] * Xr - core["U"])

    def sparsify_sum(self, Xr, Xall, row_offset):
        E = Xr @ Xall.t() + 1.0
        idx = torch.arange(Xr.shape[0])
        E[idx, idx + row_offset] -= 2.0
        return (E * E).sum()
