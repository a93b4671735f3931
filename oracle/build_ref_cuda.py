"""ORACLE / BASELINE -- test infrastructure, NOT product code.

Compiles the reference's OWN CUDA op (models/dino/ops/src: vision.cpp + cuda/ms_deform_attn_cuda.cu, the pybind module
`MultiScaleDeformableAttention` of models/dino/ops/setup.py) for sm_100a, as a second, GPU-side oracle and the "recompiled
reference kernel" baseline for the deformable-attention core.  The reference's setup.py refuses to build without a
visible GPU (ops/setup.py:38-49), so this recipe calls nvcc / g++ directly with torch's header and library paths.

No reference source enters the repository: the sources are read where they lie under /root/reference; the one file that no
longer compiles against torch 2.x (`AT_DISPATCH_FLOATING_TYPES(value.type(), ...)` at cuda/ms_deform_attn_cuda.cu:64,134 --
DeprecatedTypeProperties is no longer convertible to ScalarType) is patched in a scratch copy under /tmp (`.type()` ->
`.scalar_type()` on those two dispatch lines only).  Output: oracle/_ref/MultiScaleDeformableAttention.so (git-ignored, travels
to the GPU box with the snapshot).  Only tests/ and tests/perf_msda_vs_ref_cuda.py load it.

    python oracle/build_ref_cuda.py          (build container: needs /root/reference, nvcc, torch headers; ~2 min)
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DTLR_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "models", "dino", "ops", "src")
OUT = os.path.join(HERE, "_ref", "MultiScaleDeformableAttention.so")


def build(force=False):
    if os.path.exists(OUT) and not force:
        return OUT
    if not os.path.isdir(SRC):
        raise RuntimeError("reference sources not found under %s" % SRC)
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="dtlr_ref_cuda_")
    try:
        shutil.copytree(SRC, os.path.join(tmp, "src"))
        cu = os.path.join(tmp, "src", "cuda", "ms_deform_attn_cuda.cu")
        text = open(cu).read()
        patched, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES\(value\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()", text)
        assert n == 2, "expected the two dispatch lines of ms_deform_attn_cuda.cu:64,134 (found %d)" % n
        open(cu, "w").write(patched)
        inc = ["-I" + os.path.join(tmp, "src")] + ["-I" + p for p in ce.include_paths("cuda")] + ["-I" + sysconfig.get_paths()["include"]]
        defs = ["-DWITH_CUDA", "-DTORCH_EXTENSION_NAME=MultiScaleDeformableAttention", "-DTORCH_API_INCLUDE_EXTENSION_H",
                "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        objs = []
        for src, flags in ((os.path.join(tmp, "src", "vision.cpp"), []),
                           (os.path.join(tmp, "src", "cpu", "ms_deform_attn_cpu.cpp"), []),
                           (cu, ["-gencode", "arch=compute_100a,code=sm_100a", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                 "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"])):
            obj = os.path.join(tmp, os.path.basename(src) + ".o")
            cmd = [nvcc, "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-w"] + flags + defs + inc + ["-c", src, "-o", obj]
            subprocess.check_call(cmd)
            objs.append(obj)
        libdirs = ce.library_paths("cuda")
        link = [nvcc, "-shared", "-o", OUT] + objs + ["-L" + d for d in libdirs] + \
               ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python"] + \
               sum((["-Xlinker", "-rpath", "-Xlinker", d] for d in libdirs), [])
        subprocess.check_call(link)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


def load():
    """import the compiled reference module (torch must already be imported); None when it was not built"""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttention", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
