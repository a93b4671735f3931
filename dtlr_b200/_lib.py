"""ctypes binding of libdtlr_b200.so (include/dtlr_b200.h).  Fails loudly: there is no fallback path."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdtlr_b200.so")
_lib = None

F32, BF16, F64 = 0, 1, 2
_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64}


class DtlrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DtlrError(
                "libdtlr_b200.so is not built (%s). Run `python -m dtlr_b200.build` (needs nvcc); "
                "dtlr_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dtlr_last_error.restype = ctypes.c_char_p
        _lib.dtlr_version.restype = ctypes.c_int
        if os.environ.get("DTLR_DEBUG_FLAGS"):      # tuning / A-B switches of include/dtlr_b200.h (dtlr_debug_flags)
            _lib.dtlr_debug_flags(int(os.environ["DTLR_DEBUG_FLAGS"]))
    return _lib


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
        return _DT[t.dtype]
    except KeyError:
        raise DtlrError("unsupported dtype %s" % t.dtype)


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
