"""
Builds libsmcb200.so (hand-written sm_100a kernels + the C ABI of include/smcb200.h)
in-tree with nvcc.  `python -m smcpy_b200.build` or `__graft_entry__.build()`.
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsmcb200.so")
SOURCES = ["api.cu", "reweight.cu", "resample.cu", "moments.cu", "metropolis.cu", "usermodel.cu", "xchg.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "smcb200.h"))
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            r = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
            log = os.path.join(objdir, src + ".log")
            with open(log, "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
