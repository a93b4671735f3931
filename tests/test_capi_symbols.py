"""CPU: the C-ABI library builds/loads and exports every symbol include/dtlr_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dtlr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dtlr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dtlr_b200 import build
    build.build_library()
    syms = declared_symbols()
    assert "dtlr_msda_forward" in syms and "dtlr_msda_backward" in syms and "dtlr_ctc_loss" in syms and "dtlr_postprocess" in syms
    for path in (build.LIB, build.LIB_F16):          # the two 16-bit flavours of the same sources (bf16 | fp16 operands)
        lib = ctypes.CDLL(path)
        for s in syms:
            assert hasattr(lib, s), "%s: missing export %s" % (os.path.basename(path), s)
        lib.dtlr_version.restype = ctypes.c_int
        assert lib.dtlr_version() >= 0x000100
        assert lib.dtlr_built_for_sm() == 100


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    from dtlr_b200 import msda, _lib
    v = torch.zeros(1, 30, 2, 32)
    shp = torch.tensor([[6, 4], [3, 2]])
    lsi = torch.tensor([0, 24])
    with pytest.raises(_lib.DtlrError):
        msda.ms_deform_attn_forward(v, shp, lsi, torch.zeros(1, 2, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2), 64)
