"""Build libscb200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscb200.so")
SOURCES = ["rowwise.cu", "simt_pass.cu", "tc_pass.cu", "tc_pair.cu", "api.cu"]
HEADERS = ["common.cuh", "ptx.cuh", os.path.join("..", "..", "include", "scb200.h")]
# no --use_fast_math: the SIMT path is the exact fp32 path (expf/exp2f/division must stay IEEE-accurate)
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Wno-deprecated-gpu-targets"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = FLAGS + (["-Xptxas", "-v"] if verbose else [])
    procs, objs = [], []
    for src in SOURCES:            # one nvcc per translation unit, in parallel
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([_nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB + ".tmp", *objs,
                           "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
