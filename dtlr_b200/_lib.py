"""ctypes binding of libdtlr_b200.so (include/dtlr_b200.h).  Fails loudly: there is no fallback path."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdtlr_b200.so")
# the same sources built twice (dtlr_b200/build.py): 16-bit operand type bf16 | fp16.  fp32 / fp64 kernels exist in both.
LIB_PATHS = {"bf16": LIB_PATH, "f16": os.path.join(_HERE, "libdtlr_b200_f16.so")}
_libs = {}
FLAVOR = "bf16"

F32, BF16, F64, F16 = 0, 1, 2, 3
SPLIT16 = 4          # output-only code of dtlr_gemm / dtlr_conv2d_nhwc: the fp32 result as the 16-bit [hi | hi | lo] split operand
_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64, torch.float16: F16}
_FLAVOR_OF = {torch.bfloat16: "bf16", torch.float16: "f16"}


class DtlrError(RuntimeError):
    pass


def set_flavor(dtype_or_name):
    """select the library that serves 16-bit tensors of this dtype (torch.bfloat16 / torch.float16 or "bf16" / "f16"); fp32 / fp64
    dtypes leave the selection alone.  The engine calls this at the start of every forward; dtype_code() does it for direct op calls."""
    global FLAVOR
    name = _FLAVOR_OF.get(dtype_or_name, dtype_or_name if isinstance(dtype_or_name, str) else None)
    if name is not None:
        if name not in LIB_PATHS:
            raise DtlrError("unknown 16-bit flavour %r" % (name,))
        FLAVOR = name
    return FLAVOR


def lib():
    handle = _libs.get(FLAVOR)
    if handle is None:
        path = LIB_PATHS[FLAVOR]
        if not os.path.exists(path):
            raise DtlrError(
                "%s is not built (%s). Run `python -m dtlr_b200.build` (needs nvcc); "
                "dtlr_b200 has no CPU or PyTorch fallback." % (os.path.basename(path), path))
        handle = ctypes.CDLL(path)
        handle.dtlr_last_error.restype = ctypes.c_char_p
        handle.dtlr_version.restype = ctypes.c_int
        if os.environ.get("DTLR_DEBUG_FLAGS"):      # tuning / A-B switches of include/dtlr_b200.h (dtlr_debug_flags)
            handle.dtlr_debug_flags(int(os.environ["DTLR_DEBUG_FLAGS"]))
        _libs[FLAVOR] = handle
    return handle


# number of libdtlr_b200 kernel launches issued through the C ABI (bench.py reports it as "gpu_launches")
LAUNCHES = 0
# optional profiler hook: callable(name) -> context manager, set by bench.py to time one kernel family with CUDA events
TIMER = None
# algorithmic bytes (A + W + output [+ residual]) of the dtlr_gemm launches issued while TIMER is set (bench.py: HBM view of the family)
GEMM_BYTES = 0


def check(rc, what):
    global LAUNCHES
    if rc != 0:
        raise DtlrError("%s failed (status %d): %s" % (what, rc, lib().dtlr_last_error().decode()))
    LAUNCHES += 1


def dtype_code(t):
    try:
        code = _DT[t.dtype]
    except KeyError:
        raise DtlrError("unsupported dtype %s" % t.dtype)
    if code in (BF16, F16):
        set_flavor(t.dtype)        # a 16-bit tensor decides which of the two libraries serves the call
    return code


def ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None and t.numel() > 0 else (t.data_ptr() if t is not None else 0))


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DtlrError("dtlr_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                            % t.device.type)


def i64_host(seq):
    arr = (ctypes.c_int64 * len(seq))(*[int(v) for v in seq])
    return arr
