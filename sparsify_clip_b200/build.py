"""Build libscb200.so in-tree with nvcc for sm_100a (no torch headers: the library is plain C ABI)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscb200.so")
SOURCES = ["rowwise.cu", "simt_pass.cu", "tc_pass.cu", "api.cu"]
HEADERS = ["common.cuh", "ptx.cuh", os.path.join("..", "..", "include", "scb200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
             "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3"]
    # --use_fast_math would turn expf/exp2f/division into approximations in the exact SIMT path:
    flags.remove("--use_fast_math")
    procs = []
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        procs.append((s, subprocess.Popen([_nvcc(), *flags, "-c", os.path.join(CSRC, s), "-o", o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    fail = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(f"--- nvcc {s}\n{out}\n")
        fail |= p.returncode != 0
    if fail:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB + ".tmp", *objs, "-lcuda" if False else "-lcudart_static",
                           "-lpthread", "-ldl", "-lrt"])
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
