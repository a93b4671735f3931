"""Build libdtlr_b200.so (the C-ABI CUDA library of include/dtlr_b200.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the snapshot.
    python -m dtlr_b200.build [--force] [--verbose]
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdtlr_b200.so")            # 16-bit type = bf16
LIB_F16 = os.path.join(HERE, "libdtlr_b200_f16.so")    # the same sources with fp16 as the 16-bit type (-DDTLR_BUILD_F16)
OBJ = os.path.join(CSRC, "build")
FLAVORS = {"bf16": (LIB, OBJ, []), "f16": (LIB_F16, os.path.join(CSRC, "build", "f16"), ["-DDTLR_BUILD_F16"])}

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177",
]


def _nvcc():
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found; cannot build libdtlr_b200.so")
    return n


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, flavors=("bf16", "f16")):
    """compile every .cu for sm_100a once per 16-bit flavour (all nvcc processes run concurrently) and link the two libraries"""
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    procs, plan = [], []
    for fl in flavors:
        lib, objdir, defs = FLAVORS[fl]
        os.makedirs(objdir, exist_ok=True)
        objs, dirty = [], False
        for src in sources():
            obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            if force or _stale(obj, [src] + headers):
                dirty = True
                cmd = [_nvcc()] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
                procs.append((fl + ":" + os.path.basename(src), subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        plan.append((lib, objs, dirty))
    failed = False
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (name, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    for lib, objs, dirty in plan:
        if force or dirty or _stale(lib, objs):
            # -Bsymbolic: calls between the library's own extern "C" entry points bind inside the library (two flavours coexist)
            subprocess.check_call([_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-Bsymbolic"])
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
