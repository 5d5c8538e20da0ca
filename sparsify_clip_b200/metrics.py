"""Evaluation-side consumers of the similarity matrix, with the names and signatures of the reference
(sparsify_clip.py: compute_metric_ret :357-416, compute_gap :418-436, compute_mean_angular_value_of_a_modality :438-457,
uniformity :459-485, mean_distance_of_true_pairs :508-528), on the kernels of libscb200.so.

The reference materialises the N x N similarity (:628), sorts every row and column and walks Python lists; here
  * retrieval_ranks / retrieval_metrics work from the FEATURES: the similarity is produced tile by tile by the forward
    sweep (tensor cores for 16-bit features, fp32 FMAs otherwise) and only "how many scores beat the true pair's" leaves
    the SM -- no N x N matrix, no sort;
  * compute_metric_ret keeps the reference's interface (a given score matrix and id lists) and counts instead of sorting;
  * the mean off-diagonal cosine is (|sum_i x_i|^2 - sum_i |x_i|^2) / (N (N - 1)): O(N D), no N x N matrix;
  * the W2 uniformity needs only the eigenVALUES of the D x D covariance (tr Sigma^(1/2) = sum sqrt(lambda)): covariance
    on the GPU (scb_gram_dd), torch.linalg.eigvalsh (cuSOLVER syevd) instead of the reference's NumPy eig round trip.
All functions take CUDA tensors and return Python floats / dicts like the reference's.
"""
import math

import torch

from .backend_cuda import get_backend

__all__ = ["compute_metric_ret", "compute_gap", "compute_mean_angular_value_of_a_modality", "uniformity",
           "mean_distance_of_true_pairs", "retrieval_ranks", "retrieval_metrics"]


def _recall_log(rank, n, prefix):
    r1 = (rank < 1).sum().item() / n
    r5 = (rank < 5).sum().item() / n
    r10 = (rank < 10).sum().item() / n
    return {f"{prefix}_r1": round(r1 * 100, 4), f"{prefix}_r5": round(r5 * 100, 4), f"{prefix}_r10": round(r10 * 100, 4),
            f"{prefix}_ravg": round((r1 + r5 + r10) / 3 * 100, 4)}


def compute_metric_ret(score_matrix, ids, ids_txt, direction="forward"):
    """sparsify_clip.py:357-416.  score_matrix [N_text, N_image]; 'forward' = text-to-vision (rows), otherwise
    vision-to-text (columns, best rank over all texts of the image).  The rank of a ground-truth entry is the number
    of entries of its row / column that beat it (its position in the reference's descending sort)."""
    assert score_matrix.shape == (len(ids_txt), len(ids)), \
        f"Score matrix shape {score_matrix.shape} does not match (len(ids_txt), len(ids))"
    be = get_backend()
    S = be.prep(score_matrix)
    dev = S.device
    if direction == "forward":
        first = {}
        for k, v in enumerate(ids):
            first.setdefault(v, k)                                  # ids.index(...)
        gt = torch.tensor([first[t] for t in ids_txt], dtype=torch.int64, device=dev)
        rank = be.rank_count(S, gt, columns=False)
        return _recall_log(rank, len(ids_txt), "forward")
    by_id = {}
    for k, v in enumerate(ids_txt):
        by_id.setdefault(v, []).append(k)
    line, gt = [], []
    for i, v in enumerate(ids):
        for k in by_id[v]:
            line.append(i)
            gt.append(k)
    line_t = torch.tensor(line, dtype=torch.int64, device=dev)
    r = be.rank_count(S, torch.tensor(gt, dtype=torch.int64, device=dev), columns=True, line=line_t)
    rank = torch.full((len(ids),), 2 ** 31 - 1, dtype=torch.int32, device=dev).scatter_reduce(0, line_t, r, reduce="amin")
    return _recall_log(rank, len(ids), "backward")


def retrieval_ranks(text_feats, image_feats):
    """(rank of image i for text i [N], rank of text i for image i [N]) for paired features, straight from the features:
    two forward sweeps of the tile kernel, nothing of size N x N."""
    be = get_backend()
    t, v = be.prep(text_feats), be.prep(image_feats)
    if t.dtype != v.dtype:
        t, v = t.float(), v.float()
    diag = be.row_dot(t, v)
    fwd = be.rank_count_pass(t, v, diag)
    bwd = be.rank_count_pass(v, t, diag)
    return fwd.to(torch.int32), bwd.to(torch.int32)


def retrieval_metrics(text_feats, image_feats):
    """R@1/5/10 both ways for paired features (what evaluate_model logs, sparsify_clip.py:641-673) without the matrix."""
    fwd, bwd = retrieval_ranks(text_feats, image_feats)
    n = fwd.numel()
    out = _recall_log(fwd, n, "forward")
    out.update(_recall_log(bwd, n, "backward"))
    return out


def compute_gap(feat_modality1, feat_modality2):
    """sparsify_clip.py:418-436: Euclidean distance between the two modality centroids."""
    be = get_backend()
    a, b = be.prep(feat_modality1), be.prep(feat_modality2)
    if a.dtype != b.dtype:
        a, b = a.float(), b.float()
    return torch.linalg.vector_norm(be.col_sum(a, b, 1.0 / a.shape[0])).item()


def compute_mean_angular_value_of_a_modality(feat_modality):
    """sparsify_clip.py:438-457: mean of the off-diagonal entries of X X^T."""
    be = get_backend()
    x = be.prep(feat_modality)
    n = x.shape[0]
    s = be.col_sum(x)
    total = (s.double() * s.double()).sum() - be.sum(be.row_sqnorm(x)).double()
    return (total / (n * (n - 1))).item()


def mean_distance_of_true_pairs(features_modality1, features_modality2):
    """sparsify_clip.py:508-528: mean cosine of the true pairs = mean of the diagonal of X Y^T."""
    be = get_backend()
    a, b = be.prep(features_modality1), be.prep(features_modality2)
    if a.dtype != b.dtype:
        a, b = a.float(), b.float()
    return (be.sum(be.row_dot(a, b)) / a.shape[0]).item()


def w2_from_moments(mu, cov, dim, eps=1e-8):
    """sqrt(|mu|^2 + 1 + tr(cov) - (2 / sqrt(dim)) tr(cov^(1/2))), cov^(1/2) through the clipped spectrum + eps."""
    lam = torch.linalg.eigvalsh(cov.double())
    tr_sqrt = torch.sqrt((lam + eps).clamp_min(0)).sum()
    val = (mu.double() ** 2).sum() + 1.0 + torch.trace(cov.double()) - (2.0 / math.sqrt(dim)) * tr_sqrt
    return math.sqrt(val.item())


def uniformity(features_modality1, features_modality2):
    """sparsify_clip.py:459-485 (= uniformity.py:101 numpy_uniformity): -W2 of the two modalities' joint cloud."""
    be = get_backend()
    x = be.prep(torch.cat([features_modality1, features_modality2], dim=0))
    n, dim = x.shape
    mu = be.col_sum(x, None, 1.0 / n)
    cov = be.gram_dd(x, mu, 1.0 / n)
    return -w2_from_moments(mu, cov, dim)
