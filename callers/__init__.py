"""Callers of the loss path: the training step of sparsify_clip.py:685-965 around
`sparsify_clip_b200.compose_loss`.  Plain PyTorch; nothing here is on the hot path of the library itself."""
