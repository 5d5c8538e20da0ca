"""TEST INFRASTRUCTURE ONLY -- mint tests/golden_eval/*.npz from the UNMODIFIED reference (this container only).

The reference's evaluation-side consumers (sparsify_clip.py:357-528: compute_metric_ret, compute_gap,
compute_mean_angular_value_of_a_modality, uniformity, mean_distance_of_true_pairs) run here on CPU on seeded inputs of the
reference's own evaluation size (num_test_samples: 512, experiments_configs/*.yaml); inputs and outputs are stored so that
the GPU box -- where /root/reference does not exist -- can check the kernels against them.
    python -m oracle.make_golden_eval
"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_eval")


def main():
    ref, uni = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for name, (N, D, seed, noise) in {"eval_n512_d64": (512, 64, 42, 2.5), "eval_n300_d128": (300, 128, 7, 4.0)}.items():
        g = torch.Generator().manual_seed(seed)
        img = torch.randn(N, D, generator=g) + 0.4                     # a common offset: the "cone" of real embeddings
        txt = img + noise * torch.randn(N, D, generator=g) - 0.2
        img = img / img.norm(dim=-1, keepdim=True)                     # sparsify_clip.py:624-625
        txt = txt / txt.norm(dim=-1, keepdim=True)
        S = torch.matmul(txt, img.t())                                 # [N_text, N_image], :628
        ids = list(range(N))
        with contextlib.redirect_stdout(io.StringIO()):
            fwd = ref.compute_metric_ret(S, ids, ids, direction="forward")
            bwd = ref.compute_metric_ret(S, ids, ids, direction="backward")
            out = dict(gap=ref.compute_gap(img, txt), ang_img=ref.compute_mean_angular_value_of_a_modality(img),
                       ang_txt=ref.compute_mean_angular_value_of_a_modality(txt), unif=ref.uniformity(img, txt),
                       cos_true=ref.mean_distance_of_true_pairs(img, txt),
                       u1=float(uni.torch_uniformity1(img)), u2=float(uni.torch_uniformity(img, txt)),
                       u_eq=float(uni.torch_uniformity_equivalent(img)))
        keys_f = ["forward_r1", "forward_r5", "forward_r10", "forward_ravg"]
        keys_b = ["backward_r1", "backward_r5", "backward_r10", "backward_ravg"]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), img=img.numpy(), txt=txt.numpy(),
                            forward=np.array([fwd[k] for k in keys_f]), backward=np.array([bwd[k] for k in keys_b]),
                            **{k: np.float64(v) for k, v in out.items()})
        print(name, fwd, bwd, out)


if __name__ == "__main__":
    main()
