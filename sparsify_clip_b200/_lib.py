"""ctypes binding of libscb200.so (C ABI declared in include/scb200.h).

The library is the product: if it cannot be loaded this module raises -- there is no CPU
or PyTorch fallback anywhere in the package.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libscb200" + os.environ.get("SCB_LIB_SUFFIX", "") + ".so")   # suffix: tuning variants only
HEADER = os.path.join(os.path.dirname(HERE), "include", "scb200.h")

SCB_F32, SCB_BF16, SCB_F16 = 0, 1, 2
PATH_SIMT, PATH_TC = 0, 1

_vp, _i64, _i32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float

# name -> argtypes, in the order of include/scb200.h
SIGNATURES = {
    "scb_version": [],
    "scb_pass_plan": [_i32, _i64, _i64, _i32, _i32, _i32, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)],
    "scb_set_tc_flags": [_i32],
    "scb_grad_kernel_kind": [_i64, _i32, _i32, ctypes.POINTER(ctypes.c_int)],
    "scb_quad_plan": [_i64, _i64, _i32, _i32, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64),
                      ctypes.POINTER(ctypes.c_int)],
    "scb_row_sqnorm": [_vp, _i64, _i32, _i64, _i32, _vp, _vp],
    "scb_row_dot": [_vp, _vp, _i64, _i32, _i64, _i64, _i32, _vp, _vp],
    "scb_lalign_rows": [_vp, _vp, _i64, _i32, _i64, _i64, _i32, _vp, _vp],
    "scb_lalign_bwd": [_vp, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _vp, _i32, _vp, _vp, _vp],
    "scb_centroid_fwd": [_vp, _vp, _i64, _i32, _i64, _i64, _i32, _vp, _i32, _vp, _vp],
    "scb_centroid_bwd": [_vp, _vp, _i64, _i32, _i64, _i64, _i32, _vp, _vp, _f32, _vp, _i32, _vp, _vp, _vp],
    "scb_normalize_fwd": [_vp, _i64, _i32, _i64, _i32, _vp, _i32, _vp, _vp],
    "scb_normalize_bwd": [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _vp],
    "scb_sum": [_vp, _i64, _vp, _vp, _vp],
    "scb_lse_pass": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _i32, _vp, _vp, _i32, _vp, _vp],
    "scb_lse_combine": [_vp, _vp, _i32, _i64, _vp, _vp],
    "scb_lse2_pass": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "scb_colstat_combine": [_vp, _vp, _i32, _i64, _vp, _vp],
    "scb_colstat_partial": [_vp, _vp, _i32, _i64, _vp, _vp, _vp],
    "scb_lse2_spread_flag": [_vp, _i64, _vp, _i64, _f32, _vp, _vp, _vp],
    "scb_lse_pass_cond": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp],
    "scb_lse_combine_cond": [_vp, _vp, _i32, _i64, _vp, _vp, _vp],
    "scb_anchor_grad_pass": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp],
    "scb_anchor_grad_finalize": [_vp, _i32, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _f32, _f32, _vp, _i32, _vp, _vp, _vp],
    "scb_lunif_pass": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp],
    "scb_lunif_sum_pass": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _vp, _vp, _i64, _i32, _vp, _i32, _vp],
    "scb_lunif_grad_finalize": [_vp, _i32, _vp, _i32, _i64, _i32, _vp, _i64, _i32, _f32, _vp, _i32, _vp, _vp],
    "scb_debug_pair_trace": [_vp],
    "scb_loss_assemble": [_vp, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp],
    "scb_lse2_fold_ranks": [_vp, _i32, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp],
    "scb_grad_combine": [_vp, _vp, _i64, _i32, _i64, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _f32, _f32, _vp, _i32, _vp, _i32,
                         _f32, _vp, _f32, _vp, _f32, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _i32, _vp, _vp],
    "scb_sparsify_sum_pass": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _i64, _i32, _vp, _i32, _vp],
    "scb_col_sum": [_vp, _vp, _i64, _i32, _i64, _i64, _i32, _f32, _vp, _i32, _vp, _vp],
    "scb_gram_dd": [_vp, _i64, _i32, _i64, _i32, _vp, _f32, _vp, _i32, _vp, _vp],
    "scb_rows_times_dd": [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp],
    "scb_sum_parts": [_vp, _i32, _i64, _f32, _vp, _vp],
    "scb_rank_count": [_vp, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp],
    "scb_peer_alloc": [_i64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p],
    "scb_peer_open": [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)],
    "scb_peer_close": [_vp, _i32],
    "scb_peer_begin": [_vp, _vp, _i32, _vp],
    "scb_peer_push": [_vp, _i64, ctypes.POINTER(ctypes.c_void_p), _i32, _vp, _vp, _i32, _vp],
    "scb_peer_push_sm": [_vp, _i64, ctypes.POINTER(ctypes.c_void_p), _i32, _vp, _vp, _i32, _vp, _vp],
    "scb_peer_copy": [_vp, _i64, ctypes.POINTER(ctypes.c_void_p), _i32, _vp],
    "scb_peer_wait": [_vp, _vp, _i32, _vp],
    "scb_peer_release": [_vp, _vp, _i32, _vp],
    "scb_peer_release_many": [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), _i32, _i32, _vp],
    "scb_rank_count_pass": [_vp, _i64, _vp, _i64, _i32, _i64, _i64, _i32, _vp, _i64, _i32, _vp, _i32, _vp],
}

_lib = None


def header_symbols():
    """Every function name include/scb200.h declares (used by the CPU symbol test)."""
    with open(HEADER) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", src)))


def header_abi_version():
    with open(HEADER) as f:
        m = re.search(r"#define\s+SCB_ABI_VERSION\s+(\d+)", f.read())
    if not m:
        raise RuntimeError(f"{HEADER}: SCB_ABI_VERSION not found")
    return int(m.group(1))


def load(build_if_missing=True):
    """Load (building first if the sources are newer and nvcc is present) and type the library.

    The build is serialised across processes by a file lock (every rank of a torchrun job imports this module at the same
    moment).  A failed compile is an error, never a silent fall-back to a stale binary; only a box WITHOUT nvcc may use
    a prebuilt library, and in every case the library's ABI revision must match include/scb200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        if _build.have_nvcc():
            _build.build_locked()
        elif not os.path.exists(LIB_PATH):
            raise RuntimeError("libscb200.so is missing and there is no nvcc to build it (there is no CPU fallback)")
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m sparsify_clip_b200.build` "
                           "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.scb_version.argtypes = []
    lib.scb_version.restype = ctypes.c_int
    have, want = lib.scb_version(), header_abi_version()
    if have != want:
        raise RuntimeError(f"{LIB_PATH} was built for ABI revision {have}, include/scb200.h declares {want}: rebuild it "
                           "(`python -m sparsify_clip_b200.build --force`)")
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.scb_last_error.argtypes = []
    lib.scb_last_error.restype = ctypes.c_char_p
    if os.environ.get("SCB_TC_FLAGS"):          # tuning/debug knob, see scb_set_tc_flags in include/scb200.h
        lib.scb_set_tc_flags(int(os.environ["SCB_TC_FLAGS"]))
    _lib = lib
    return lib


def check(rc, what=""):
    if rc == 0:
        return
    msg = load().scb_last_error().decode(errors="replace")
    if rc < 0:
        raise ValueError(f"scb200 {what}: {msg} (code {rc})")
    raise RuntimeError(f"scb200 {what}: CUDA error {rc}: {msg}")
