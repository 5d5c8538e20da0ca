"""The loss_type composition ladder and the alpha/beta schedules of the reference training loop.

Reference: sparsify_clip.py:775-938 (ladder, inline in train_model) and :41-64 (get_beta /
get_alpha).  Behaviour is reproduced AS CODED:
  * every "only_lunif_n_then_*" type returns (lunif(img) + lunif(txt))/2 while
    epoch < config["only_lunif_epochs"]                                   (:783-786 ...)
  * the string "only_lunif_n_then_anchor+lalign+BETA*lunif(centroids)" is tested twice
    (:813 and :833); the first branch wins, so experiment 7 AND experiment 8 both compute
    anchor + lalign + beta*(lunif(img)+lunif(txt))/2.  The intended-but-unreachable
    centroid variant is available as loss_type "...BETA*lunif(centroids)[intended]".
  * an unknown loss_type is an error (the reference dies at loss.item() one line later).
"""
from .losses import (contrastive_loss, lalign_loss, lunif_loss, normalized_centroids, centroid_operand_dtype,
                     fused_terms_loss, l2_normalize)

__all__ = ["get_beta", "get_alpha", "ladder_weights", "compose_loss", "weighted_loss", "set_fused", "LOSS_TYPES"]


def get_beta(current_step, total_steps, warmup_epoch=20, decay_epoch=50):
    """Weight on L_unif: 1 until warm-up ends, then linear to 0 (sparsify_clip.py:41-51)."""
    per_epoch = total_steps / 100          # the reference hard-codes 100 epochs
    start, length = warmup_epoch * per_epoch, decay_epoch * per_epoch
    if current_step < start:
        return 1.0
    if current_step < start + length:
        return 1.0 - float(current_step - start) / float(max(1, length))
    return 0.0


def get_alpha(current_step, total_steps, warmup_epoch=20, increment_epoch=50):
    """Weight on L_align: 1 until warm-up ends, then linear to 2 (sparsify_clip.py:54-64)."""
    per_epoch = total_steps / 100
    start, length = warmup_epoch * per_epoch, increment_epoch * per_epoch
    if current_step < start:
        return 1.0
    if current_step < start + length:
        return 1.0 + float(current_step - start) / float(max(1, length))
    return 2.0


# loss_type -> (has lunif-only warm-up, align weight, unif target, unif weight); "A"/"B" = alpha/beta schedule
_WARM = "only_lunif_n_then_"
LOSS_TYPES = {
    "anchor":                                                       (False, 0.0, None, 0.0),
    _WARM + "anchor+lalign+lunif(text)+lunif(img)":                 (True, 1.0, "modalities", 1.0),
    _WARM + "anchor+lalign+lunif(centroids)":                       (True, 1.0, "centroids", 1.0),
    _WARM + "anchor+lalign+BETA*lunif(centroids)":                  (True, 1.0, "modalities", "B"),   # as coded (:813)
    _WARM + "anchor+lalign+BETA*lunif(centroids)[intended]":        (True, 1.0, "centroids", "B"),    # :833, unreachable
    _WARM + "anchor+ALPHA*lalign+BETA*(lunif(text)+lunif(img))":    (True, "A", "modalities", "B"),
    _WARM + "anchor+ALPHA*lalign+BETA*lunif(centroids)":            (True, "A", "centroids", "B"),
    "ANCHOR(IMAGE,TEXT)+LALIGN(IMAGE,TEXT)+LUNIF(CENTROIDS)":       (False, 1.0, "centroids", 1.0),
    "ANCHOR(IMAGE,TEXT)+LALIGN(IMAGE,TEXT)":                        (False, 1.0, None, 0.0),
    "ANCHOR(IMAGE,TEXT)+LUNIF(CENTROIDS)":                          (False, 0.0, "centroids", 1.0),
}


def ladder_weights(config, epoch, current_batch, t_total):
    """-> dict(anchor, align, unif_img, unif_txt, unif_cen, alpha, beta) for this step."""
    lt = config["loss_type"]
    if lt not in LOSS_TYPES:
        raise KeyError(f"loss_type {lt!r} matches no branch of the ladder (sparsify_clip.py:775-938)")
    warm, w_align, target, w_unif = LOSS_TYPES[lt]
    out = dict(anchor=1.0, align=0.0, unif_img=0.0, unif_txt=0.0, unif_cen=0.0, alpha=0.0, beta=0.0)
    if warm and epoch < config["only_lunif_epochs"]:
        out.update(anchor=0.0, unif_img=0.5, unif_txt=0.5)
        return out
    if w_unif == "B":
        w_unif = out["beta"] = get_beta(current_batch, t_total, config["beta_warmup_epoch"], config["beta_decay_epoch"])
    if w_align == "A":
        w_align = out["alpha"] = get_alpha(current_batch, t_total, config["alpha_warmup_epoch"],
                                           config["alpha_increment_epoch"])
    out["align"] = float(w_align)
    if target == "modalities":
        out["unif_img"] = out["unif_txt"] = 0.5 * float(w_unif)
    elif target == "centroids":
        out["unif_cen"] = float(w_unif)
    return out


_fuse = True


def set_fused(flag):
    """Evaluate the compositions as one fused autograd node (default) or term by term."""
    global _fuse
    prev, _fuse = _fuse, bool(flag)
    return prev


def weighted_loss(image_embeds, text_embeds, temperature, w, *, group=None, normalize=False):
    """sum of the selected terms; a zero weight skips the kernel (its gradient is exactly 0).
    normalize=True: the embeddings are the encoders' raw outputs and the pre-loss normalise of the training loop
    (sparsify_clip.py:772-773) is part of this call -- inside the fused node, where its backward rides on the gradient
    combine pass; as a separate l2_normalize in front of the modular path."""
    if _fuse and any(w[k] != 0.0 for k in ("anchor", "align", "unif_img", "unif_txt", "unif_cen")):
        return fused_terms_loss(image_embeds, text_embeds, temperature, w["anchor"], w["align"], w["unif_img"],
                                w["unif_txt"], w_unif_cen=w["unif_cen"], group=group, normalize=normalize)
    if normalize:
        image_embeds, text_embeds = l2_normalize(image_embeds), l2_normalize(text_embeds)
    loss = None

    def add(acc, wt, term):
        term = term if wt == 1.0 else wt * term
        return term if acc is None else acc + term

    if w["anchor"] != 0.0:
        loss = add(loss, w["anchor"], contrastive_loss(image_embeds, text_embeds, temperature, group=group))
    if w["align"] != 0.0:
        loss = add(loss, w["align"], lalign_loss(image_embeds, text_embeds, group=group))
    if w["unif_img"] != 0.0:
        loss = add(loss, w["unif_img"], lunif_loss(image_embeds, group=group))
    if w["unif_txt"] != 0.0:
        loss = add(loss, w["unif_txt"], lunif_loss(text_embeds, group=group))
    if w["unif_cen"] != 0.0:
        loss = add(loss, w["unif_cen"], lunif_loss(normalized_centroids(image_embeds, text_embeds), group=group,
                                                   mma_dtype=centroid_operand_dtype(image_embeds)))
    if loss is None:
        loss = image_embeds.sum() * 0.0
    return loss


def compose_loss(config, image_embeds, text_embeds, temperature, epoch=0, current_batch=1, t_total=100, *, group=None,
                 normalize=False):
    """The per-batch loss of the reference training loop for config["loss_type"] (normalize=True: from the encoders'
    raw outputs, sparsify_clip.py:768-773 included)."""
    return weighted_loss(image_embeds, text_embeds, temperature,
                         ladder_weights(config, epoch, current_batch, t_total), group=group, normalize=normalize)
