"""TEST INFRASTRUCTURE ONLY -- numpy fp64 restatement of the reference loss path.

Every function states the reference lines it follows (paths relative to
/root/reference).  Values AND analytic gradients are given in closed form so the
oracle also works at sizes where autograd over B x B tensors is too large.
Pinned against the reference itself by tests/test_oracle.py (via
oracle/ref_loader.py in this container) and against tests/golden/*.npz anywhere.
"""
import numpy as np

F64 = np.float64


def _f64(a):
    return np.asarray(a, dtype=F64)


def _lse(s, axis):
    m = np.max(s, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(s - m), axis=axis, keepdims=True))).squeeze(axis)


# ----------------------------------------------------------------------------
# anchor / CLIP InfoNCE -- sparsify_clip.py:110-132
#   logits = I @ T.t() / temperature                        (:119-120)
#   loss = (CE(logits, arange) + CE(logits.t(), arange))/2  (:124-132)
# ----------------------------------------------------------------------------
def contrastive_loss(image_embeds, text_embeds, temperature=0.07, need_grad=True):
    I, T = _f64(image_embeds), _f64(text_embeds)
    tau = float(temperature)
    B = I.shape[0]
    S = (I @ T.T) / tau
    r = _lse(S, 1)            # row log-sum-exp  (image -> text)
    c = _lse(S, 0)            # column log-sum-exp (text -> image)
    d = np.diagonal(S)
    loss = float((np.sum(r - d) + np.sum(c - d)) / (2.0 * B))
    if not need_grad:
        return loss
    G = (np.exp(S - r[:, None]) + np.exp(S - c[None, :])) / (2.0 * B)
    G[np.arange(B), np.arange(B)] -= 1.0 / B
    dI = (G @ T) / tau
    dT = (G.T @ I) / tau
    dtau = float(-np.sum(G * S) / tau)
    return loss, dI, dT, dtau


# ----------------------------------------------------------------------------
# L_unif -- sparsify_clip.py:159-164
#   torch.pdist(x, p=2).pow(2).mul(-t).exp().mean().log()  (strict i<j pairs)
# ----------------------------------------------------------------------------
def lunif_loss(x, t=2, need_grad=True):
    X = _f64(x)
    B = X.shape[0]
    if B < 2:  # mean over an empty pdist vector -> nan, like the reference
        return (float("nan"), np.full_like(X, np.nan)) if need_grad else float("nan")
    n = np.sum(X * X, axis=1)
    d2 = np.maximum(n[:, None] + n[None, :] - 2.0 * (X @ X.T), 0.0)
    W = np.exp(-float(t) * d2)
    W[np.arange(B), np.arange(B)] = 0.0
    ssum = 0.5 * np.sum(W)
    loss = float(np.log(ssum / (B * (B - 1) / 2.0)))
    if not need_grad:
        return loss
    rs = np.sum(W, axis=1)
    dX = (-2.0 * float(t) / ssum) * (rs[:, None] * X - W @ X)
    return loss, dX


# ----------------------------------------------------------------------------
# L_align -- sparsify_clip.py:186-187     (x - y).norm(dim=1).pow(alpha).mean()
# ----------------------------------------------------------------------------
def lalign_loss(x, y, alpha=2, need_grad=True):
    X, Y = _f64(x), _f64(y)
    B = X.shape[0]
    diff = X - Y
    nrm = np.sqrt(np.sum(diff * diff, axis=1))
    loss = float(np.mean(nrm ** alpha))
    if not need_grad:
        return loss
    if alpha == 2:
        dX = 2.0 * diff / B
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            coef = np.where(nrm > 0, alpha * nrm ** (alpha - 2), 0.0)
        dX = coef[:, None] * diff / B
    return loss, dX, -dX


# ----------------------------------------------------------------------------
# centroids -- sparsify_clip.py:334-355 ((a+b)/2) followed by
# F.normalize(centroids, dim=-1) at the call sites (:803-805 ...), eps = 1e-12.
# ----------------------------------------------------------------------------
def compute_centroids_only(text_embeddings, visual_embeddings):
    return (_f64(text_embeddings) + _f64(visual_embeddings)) / 2.0


def normalize_eps(m, eps=1e-12):
    m = _f64(m)
    nrm = np.maximum(np.sqrt(np.sum(m * m, axis=-1, keepdims=True)), eps)
    return m / nrm


def normalized_centroids(a, b):
    return normalize_eps(compute_centroids_only(a, b))


def normalized_centroids_backward(a, b, dc):
    """Given d loss / d c for c = normalize((a+b)/2): returns (da, db)."""
    m = compute_centroids_only(a, b)
    nrm = np.maximum(np.sqrt(np.sum(m * m, axis=-1, keepdims=True)), 1e-12)
    c = m / nrm
    dm = (_f64(dc) - c * np.sum(c * _f64(dc), axis=-1, keepdims=True)) / nrm
    return dm / 2.0, dm / 2.0


# pre-loss normalise -- sparsify_clip.py:772-773   e / e.norm(dim=-1, keepdim=True)  (no eps)
def l2_normalize(e):
    e = _f64(e)
    return e / np.sqrt(np.sum(e * e, axis=-1, keepdims=True))


def l2_normalize_backward(e, dxhat):
    e = _f64(e)
    nrm = np.sqrt(np.sum(e * e, axis=-1, keepdims=True))
    xh = e / nrm
    return (_f64(dxhat) - xh * np.sum(xh * _f64(dxhat), axis=-1, keepdims=True)) / nrm


# ----------------------------------------------------------------------------
# cold variants (signature surface) -- sparsify_clip.py:166-176, :135-157, :487-505, :308-332
# ----------------------------------------------------------------------------
def sparsify_loss(x, need_grad=True):
    X = _f64(x)
    B = X.shape[0]
    E = X @ X.T - (2.0 * np.eye(B) - 1.0)
    loss = float(np.mean(E * E))
    if not need_grad:
        return loss
    dX = (2.0 / (B * B)) * (E + E.T) @ X
    return loss, dX


def contrastive_loss_roberta(image_embeds, text_embeds, roberta_similarity, temperature=0.07):
    # soft-target cross entropy: -mean_i sum_j p_ij * log_softmax(S)_ij   (:149-154)
    I, T, R = _f64(image_embeds), _f64(text_embeds), _f64(roberta_similarity)
    S = (I @ T.T) / float(temperature)
    ls_r = S - _lse(S, 1)[:, None]
    ls_c = S.T - _lse(S, 0)[:, None]
    li2t = -np.mean(np.sum(R * ls_r, axis=1))
    lt2i = -np.mean(np.sum(R.T * ls_c, axis=1))
    return float((li2t + lt2i) / 2.0)


def centroid_alignment_loss(img_embeds, txt_embeds, p=2):
    d = np.mean(_f64(img_embeds), axis=0) - np.mean(_f64(txt_embeds), axis=0)
    return float(np.sum(np.abs(d) ** p) ** (1.0 / p))


def compute_centroids(text_embeddings, visual_embeddings):
    c = (_f64(text_embeddings)[:, None, :] + _f64(visual_embeddings)[None, :, :]) / 2.0
    return np.sqrt(np.sum(c * c, axis=-1)), c


# ----------------------------------------------------------------------------
# schedules -- sparsify_clip.py:41-51 (get_beta), :54-64 (get_alpha)
# ----------------------------------------------------------------------------
def get_beta(current_step, total_steps, warmup_epoch=20, decay_epoch=50):
    e = total_steps / 100
    if current_step < warmup_epoch * e:
        return 1.0
    if current_step < (warmup_epoch + decay_epoch) * e:
        return 1.0 - float(current_step - warmup_epoch * e) / float(max(1, decay_epoch * e))
    return 0.0


def get_alpha(current_step, total_steps, warmup_epoch=20, increment_epoch=50):
    e = total_steps / 100
    if current_step < warmup_epoch * e:
        return 1.0
    if current_step < (warmup_epoch + increment_epoch) * e:
        return 1.0 + float(current_step - warmup_epoch * e) / float(max(1, increment_epoch * e))
    return 2.0


# ----------------------------------------------------------------------------
# loss-composition ladder -- sparsify_clip.py:775-938 (as coded, including the
# unreachable "EXP 8" branch: both exp-7 and exp-8 YAML strings hit :813 first)
# Returns (loss, dI, dT, dtau) for the embeddings as given (already normalised).
# ----------------------------------------------------------------------------
LUNIF_WARMUP_TYPES = (
    "only_lunif_n_then_anchor+lalign+lunif(text)+lunif(img)",
    "only_lunif_n_then_anchor+lalign+lunif(centroids)",
    "only_lunif_n_then_anchor+lalign+BETA*lunif(centroids)",
    "only_lunif_n_then_anchor+ALPHA*lalign+BETA*(lunif(text)+lunif(img))",
    "only_lunif_n_then_anchor+ALPHA*lalign+BETA*lunif(centroids)",
)


def ladder_terms(config, epoch, current_batch, t_total):
    """(w_anchor, w_align, w_unif_img, w_unif_txt, w_unif_centroid) as coded."""
    lt = config["loss_type"]
    if lt in LUNIF_WARMUP_TYPES and epoch < config["only_lunif_epochs"]:
        return 0.0, 0.0, 0.5, 0.5, 0.0
    if lt == "anchor":
        return 1.0, 0.0, 0.0, 0.0, 0.0
    if lt == "only_lunif_n_then_anchor+lalign+lunif(text)+lunif(img)":
        return 1.0, 1.0, 0.5, 0.5, 0.0
    if lt == "only_lunif_n_then_anchor+lalign+lunif(centroids)":
        return 1.0, 1.0, 0.0, 0.0, 1.0
    if lt == "only_lunif_n_then_anchor+lalign+BETA*lunif(centroids)":  # :813 wins over :833
        beta = get_beta(current_batch, t_total, config["beta_warmup_epoch"], config["beta_decay_epoch"])
        return 1.0, 1.0, 0.5 * beta, 0.5 * beta, 0.0
    if lt == "only_lunif_n_then_anchor+ALPHA*lalign+BETA*(lunif(text)+lunif(img))":
        beta = get_beta(current_batch, t_total, config["beta_warmup_epoch"], config["beta_decay_epoch"])
        alpha = get_alpha(current_batch, t_total, config["alpha_warmup_epoch"], config["alpha_increment_epoch"])
        return 1.0, alpha, 0.5 * beta, 0.5 * beta, 0.0
    if lt == "only_lunif_n_then_anchor+ALPHA*lalign+BETA*lunif(centroids)":
        beta = get_beta(current_batch, t_total, config["beta_warmup_epoch"], config["beta_decay_epoch"])
        alpha = get_alpha(current_batch, t_total, config["alpha_warmup_epoch"], config["alpha_increment_epoch"])
        return 1.0, alpha, 0.0, 0.0, beta
    if lt == "ANCHOR(IMAGE,TEXT)+LALIGN(IMAGE,TEXT)+LUNIF(CENTROIDS)":
        return 1.0, 1.0, 0.0, 0.0, 1.0
    if lt == "ANCHOR(IMAGE,TEXT)+LALIGN(IMAGE,TEXT)":
        return 1.0, 1.0, 0.0, 0.0, 0.0
    if lt == "ANCHOR(IMAGE,TEXT)+LUNIF(CENTROIDS)":
        return 1.0, 0.0, 0.0, 0.0, 1.0
    raise KeyError(f"loss_type {lt!r} matches no branch of the reference ladder")


def weighted_loss(I, T, tau, w_anchor, w_align, w_ui, w_ut, w_uc, t=2):
    """loss = w_anchor*A + w_align*Al + w_ui*U(I) + w_ut*U(T) + w_uc*U(normalize((I+T)/2))."""
    I, T = _f64(I), _f64(T)
    loss, dI, dT, dtau = 0.0, np.zeros_like(I), np.zeros_like(T), 0.0
    terms = {}
    if w_anchor != 0.0:
        a, gi, gt, gtau = contrastive_loss(I, T, tau)
        terms["anchor"] = a
        loss += w_anchor * a; dI += w_anchor * gi; dT += w_anchor * gt; dtau += w_anchor * gtau
    if w_align != 0.0:
        a, gi, gt = lalign_loss(I, T)
        terms["lalign"] = a
        loss += w_align * a; dI += w_align * gi; dT += w_align * gt
    if w_ui != 0.0:
        a, g = lunif_loss(I, t)
        terms["lunif_img"] = a
        loss += w_ui * a; dI += w_ui * g
    if w_ut != 0.0:
        a, g = lunif_loss(T, t)
        terms["lunif_txt"] = a
        loss += w_ut * a; dT += w_ut * g
    if w_uc != 0.0:
        c = normalized_centroids(I, T)
        a, g = lunif_loss(c, t)
        terms["lunif_centroids"] = a
        da, db = normalized_centroids_backward(I, T, g)
        loss += w_uc * a; dI += w_uc * da; dT += w_uc * db
    return loss, dI, dT, dtau, terms


def compose_loss(config, I, T, tau, epoch=0, current_batch=1, t_total=100):
    w = ladder_terms(config, epoch, current_batch, t_total)
    return weighted_loss(I, T, tau, *w)


# ----------------------------------------------------------------------------
# evaluation-side consumers -- sparsify_clip.py:357-528 (numpy restatements; ranks through a full descending sort,
# like the reference)
# ----------------------------------------------------------------------------
def retrieval_ranks(score_matrix, ids, ids_txt):
    """(forward ranks [N_text], backward ranks [N_image]) as compute_metric_ret computes them (:374-378, :396-400)."""
    S = _f64(score_matrix)
    order_r = np.argsort(-S, axis=1, kind="stable")
    fwd = []
    for i in range(len(ids_txt)):
        gt = ids.index(ids_txt[i])
        fwd.append(int(np.where(order_r[i] == gt)[0][0]))
    order_c = np.argsort(-S, axis=0, kind="stable").T
    bwd = []
    for i in range(len(ids)):
        gts = [k for k, t in enumerate(ids_txt) if t == ids[i]]
        bwd.append(min(int(np.where(order_c[i] == g)[0][0]) for g in gts))
    return np.array(fwd), np.array(bwd)


def recall_log(rank, prefix):
    n = len(rank)
    r1, r5, r10 = (rank < 1).sum() / n, (rank < 5).sum() / n, (rank < 10).sum() / n
    return {f"{prefix}_r1": round(r1 * 100, 4), f"{prefix}_r5": round(r5 * 100, 4), f"{prefix}_r10": round(r10 * 100, 4),
            f"{prefix}_ravg": round((r1 + r5 + r10) / 3 * 100, 4)}


def compute_gap(f1, f2):                                   # :418-436
    return float(np.linalg.norm(np.mean(_f64(f1), axis=0) - np.mean(_f64(f2), axis=0)))


def mean_angular_value(f):                                 # :438-457
    X = _f64(f)
    G = X @ X.T
    n = X.shape[0]
    return float((G.sum() - np.trace(G)) / (n * (n - 1)))


def mean_true_pair_cosine(f1, f2):                         # :508-528
    return float(np.mean(np.sum(_f64(f1) * _f64(f2), axis=1)))


def w2_uniformity(f1, f2, eps=1e-8):                       # :459-485 / uniformity.py:101-128 (returns -W2)
    x = np.concatenate([_f64(f1), _f64(f2)], axis=0)
    n, dim = x.shape
    mu = x.mean(axis=0)
    xc = x - mu
    cov = xc.T @ xc / n
    lam = np.linalg.eigvalsh(cov)
    tr_sqrt = np.sqrt(np.clip(lam + eps, 0, None)).sum()
    return -float(np.sqrt(np.sum(mu * mu) + 1.0 + np.trace(cov) - 2.0 / np.sqrt(dim) * tr_sqrt))
