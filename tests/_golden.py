"""Loader for tests/golden/*.npz (minted by oracle/make_golden.py from the reference)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _bits_to_f32(a):
    return (a.astype(np.uint32) << 16).view(np.float32)


def load(name):
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    if bool(z["inputs_are_bf16_bits"]):
        z["I"] = _bits_to_f32(z["I"])
        z["T"] = _bits_to_f32(z["T"])
    z["name"] = name
    return z


def grad_check(z, key, got, rtol, row_factor=1.0):
    """Compare a full gradient `got` [B,D] against golden `f64_<key>` (full or sampled rows)."""
    got = np.asarray(got, dtype=np.float64)
    if "f64_" + key in z:
        ref = np.asarray(z["f64_" + key], dtype=np.float64)
        err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
        assert err <= rtol, f"{z['name']}:{key} rel fro err {err:.3e} > {rtol}"
        return err
    rows = z["sample_rows"]
    ref_rows = z["f64_" + key + "_rows"]
    fro = float(z["f64_" + key + "_fro"])
    scale = fro / np.sqrt(got.shape[0])          # typical row norm
    err_rows = np.linalg.norm(got[rows] - ref_rows, axis=1) / max(scale, 1e-30)
    assert err_rows.max() <= rtol * row_factor, f"{z['name']}:{key} sampled-row err {err_rows.max():.3e} > {rtol * row_factor}"
    err_fro = abs(np.linalg.norm(got) - fro) / max(fro, 1e-30)
    assert err_fro <= rtol, f"{z['name']}:{key} fro-norm err {err_fro:.3e} > {rtol}"
    return max(err_rows.max(), err_fro)
