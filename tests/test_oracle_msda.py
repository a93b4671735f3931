"""CPU: the C restatement of the MSDA core (oracle/msda_ref.c) against the committed known-answer vectors that
tests/golden/make_golden.py produced with the reference's own ms_deform_attn_core_pytorch on the fixture of
reference models/dino/ops/test.py:21-60 (N=1,M=2,Lq=2,L=2,P=2, shapes (6,4),(3,2), seed 3)."""
import os

import numpy as np
import pytest
import torch

from oracle import msda as omsda


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "msda_kat.npz"))


@pytest.mark.parametrize("D", [2, 30, 32, 64, 71])
def test_forward_double_matches_reference(kat, D):
    t = lambda k: torch.from_numpy(kat["D%d_%s" % (D, k)])
    out = omsda.msda_forward(t("value"), kat["shapes"], kat["lsi"], t("loc"), t("w"))
    assert torch.allclose(out, t("out"))            # reference test.py:40 (default allclose tolerances, fp64)


@pytest.mark.parametrize("D", [2, 32])
def test_forward_float_matches_reference(kat, D):
    t = lambda k: torch.from_numpy(kat["D%d_%s" % (D, k)])
    out = omsda.msda_forward(t("value").float(), kat["shapes"], kat["lsi"], t("loc").float(), t("w").float())
    assert torch.allclose(out.double(), t("out"), rtol=1e-2, atol=1e-3)   # reference test.py:56


@pytest.mark.parametrize("D", [2, 30, 32, 64, 71])
def test_backward_matches_reference_autograd(kat, D):
    t = lambda k: torch.from_numpy(kat["D%d_%s" % (D, k)])
    gv, gl, gw = omsda.msda_backward(t("value"), kat["shapes"], kat["lsi"], t("loc"), t("w"), t("gout"))
    assert torch.allclose(gv, t("gvalue"), rtol=1e-9, atol=1e-14)
    assert torch.allclose(gl, t("gloc"), rtol=1e-9, atol=1e-14)
    assert torch.allclose(gw, t("gw"), rtol=1e-9, atol=1e-14)


def msda_case_A():
    """config-A sized call with out-of-range points; inputs regenerated from the seed used by make_golden.py."""
    g = torch.Generator().manual_seed(11)
    value = torch.randn(2, 912, 8, 32, generator=g)
    loc = torch.rand(2, 37, 8, 4, 4, 2, generator=g) * 1.3 - 0.15
    w = torch.softmax(torch.randn(2, 37, 8, 16, generator=g), -1).view(2, 37, 8, 4, 4)
    return value, loc, w


def test_forward_config_A_shape(kat):
    value, loc, w = msda_case_A()
    out = omsda.msda_forward(value, kat["A_shapes"], kat["A_lsi"], loc, w)
    assert torch.allclose(out, torch.from_numpy(kat["A_out"]), rtol=1e-4, atol=1e-5)


def test_empty_queries(kat):
    value, loc, w = msda_case_A()
    out = omsda.msda_forward(value, kat["A_shapes"], kat["A_lsi"], loc[:, :0], w[:, :0])
    assert out.shape == (2, 0, 256)
