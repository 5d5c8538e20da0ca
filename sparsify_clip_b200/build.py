"""Build libscb200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SCB_LIB_SUFFIX / SCB_EXTRA_FLAGS (tuning experiments): build a variant library next to the shipped one
LIB = os.path.join(HERE, "libscb200" + os.environ.get("SCB_LIB_SUFFIX", "") + ".so")
SOURCES = ["rowwise.cu", "simt_pass.cu", "tc_pass.cu", "tc_pair.cu", "tc_quad.cu", "metrics.cu", "api.cu"]
HEADERS = ["common.cuh", "ptx.cuh", os.path.join("..", "..", "include", "scb200.h")]
# no --use_fast_math: the SIMT path is the exact fp32 path (expf/exp2f/division must stay IEEE-accurate)
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Wno-deprecated-gpu-targets"]


# SCB_DEV=1 in the environment: development build (verbose, short watchdog; tracer and experiment knobs of the pair kernel)
FLAGS += os.environ.get("SCB_EXTRA_FLAGS", "").split()
if os.environ.get("SCB_TRACE"):          # timeline tracer only (tools/pair_trace.py, tools/quad_trace.py), release speed otherwise
    FLAGS += ["-DSCB_PAIR_TRACE"]
if os.environ.get("SCB_DEV"):
    FLAGS += ["-DSCB_TC_WATCHDOG_VERBOSE", "-DSCB_TC_WATCHDOG_NS=3000000000ull", "-DSCB_PAIR_TRACE", "-DSCB_PAIR_EXPERIMENTS"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def have_nvcc():
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile and link; every intermediate file carries this process id, the final rename is atomic."""
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    tag = f".{os.getpid()}"
    flags = FLAGS + (["-Xptxas", "-v"] if verbose else [])
    procs, objs = [], []
    for src in SOURCES:            # one nvcc per translation unit, in parallel
        obj = os.path.join(objdir, src.replace(".cu", tag + ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    try:
        for src, p in procs:
            out, _ = p.communicate()
            if verbose or p.returncode:
                sys.stderr.write(f"--- nvcc {src}\n{out}\n")
            failed |= p.returncode != 0
        if failed:
            raise RuntimeError("nvcc failed (see the compiler output above)")
        tmp = LIB + tag + ".tmp"
        subprocess.check_call([_nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", tmp, *objs,
                               "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
        os.replace(tmp, LIB)
    finally:
        for obj in objs:
            if os.path.exists(obj):
                os.remove(obj)
    return LIB


def build_locked(force=False, verbose=False):
    """build() under an exclusive inter-process lock: the ranks of a multi-process job all import the package at once;
    one compiles, the others wait and then find the library up to date."""
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            return build(force=force, verbose=verbose)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build_locked(force="--force" in sys.argv, verbose="-v" in sys.argv))
