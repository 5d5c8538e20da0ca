"""sparsify_clip_b200 -- B200-native (sm_100a) contrastive-loss hot path of noostale/sparsify-clip.

``from sparsify_clip_b200 import *`` provides the reference's loss functions under their own
names (contrastive_loss, lunif_loss, lalign_loss, compute_centroids_only, ...), the
loss_type ladder (compose_loss) and the alpha/beta schedules.  Compute happens in
libscb200.so (hand-written CUDA: TMA + tcgen05/TMEM tile kernels); importing this package
fails if that library cannot be loaded.
"""
from . import _lib
from .backend_cuda import choose_jparts, force_path, get_backend, pair_span_plan, set_fp32_mode
from .ladder import LOSS_TYPES, compose_loss, get_alpha, get_beta, ladder_weights, set_fused, weighted_loss
from .losses import (centroid_alignment_loss, compute_centroids, compute_centroids_only, contrastive_loss,
                     contrastive_loss_roberta, fused_terms_loss, l2_normalize, lalign_loss, lunif_loss, normalized_centroids, operand_dtype, centroid_operand_dtype,
                     random_alignment_loss, sparsify_loss)

from . import metrics  # noqa: E402  (evaluation-side consumers: sparsify_clip.py:357-528)

_lib.load()   # fail loudly at import time: there is no CPU fallback

__all__ = [
    "contrastive_loss", "lunif_loss", "lalign_loss", "compute_centroids_only", "compute_centroids",
    "sparsify_loss", "random_alignment_loss", "contrastive_loss_roberta", "centroid_alignment_loss",
    "normalized_centroids", "l2_normalize", "get_beta", "get_alpha", "compose_loss", "weighted_loss",
    "ladder_weights", "LOSS_TYPES", "set_fp32_mode", "operand_dtype", "centroid_operand_dtype", "fused_terms_loss",
    "set_fused",
]
