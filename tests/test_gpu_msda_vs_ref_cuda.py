"""GPU: our deformable-attention kernels against the reference's OWN CUDA kernels (models/dino/ops/src/cuda, compiled unmodified
except for the two torch-2.x dispatch lines by oracle/build_ref_cuda.py into oracle/_ref/) on the same device tensors.
Skipped when that module was not built (it is built in the development container, where /root/reference exists, and travels to the
GPU box with the snapshot).  This is the second, GPU-side oracle of SURVEY 8c; the CPU oracle tests are test_gpu_msda.py."""
import pytest
import torch

from oracle import build_ref_cuda

pytestmark = pytest.mark.gpu

A_SHAPES = [(5, 128), (3, 64), (2, 32), (1, 16)]


@pytest.fixture(scope="module")
def ref():
    try:
        mod = build_ref_cuda.load()
    except Exception as e:      # e.g. built against another torch ABI: a baseline that cannot load is skipped, not failed
        pytest.skip("oracle/_ref/MultiScaleDeformableAttention.so does not load here: %s" % e)
    if mod is None:
        pytest.skip("oracle/_ref/MultiScaleDeformableAttention.so not built (python oracle/build_ref_cuda.py)")
    return mod


def _case(B, Lq, shapes, P, seed, D=32, M=8):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shp = torch.tensor(shapes, dtype=torch.long, device="cuda")
    lsi = torch.cat((shp.new_zeros((1,)), shp.prod(1).cumsum(0)[:-1]))
    S, L = int(shp.prod(1).sum()), len(shapes)
    value = torch.randn(B, S, M, D, device="cuda", generator=g)
    loc = torch.rand(B, Lq, M, L, P, 2, device="cuda", generator=g) * 1.2 - 0.1
    w = torch.softmax(torch.randn(B, Lq, M, L * P, device="cuda", generator=g), -1).view(B, Lq, M, L, P)
    go = torch.randn(B, Lq, M * D, device="cuda", generator=g)
    return value, shp, lsi, loc, w, go


@pytest.mark.parametrize("B,Lq,shapes,P,D", [(2, 900, A_SHAPES, 4, 32), (4, 100, [(12, 167), (6, 84), (3, 42), (2, 21)], 4, 32),
                                             (2, 57, [(6, 4), (3, 2)], 2, 30), (64, 912, A_SHAPES, 4, 32)])
def test_forward_matches_reference_cuda_kernel(ref, B, Lq, shapes, P, D):
    from dtlr_b200 import msda
    value, shp, lsi, loc, w, _ = _case(B, Lq, shapes, P, 11, D=D, M=8 if D == 32 else 2)
    want = ref.ms_deform_attn_forward(value, shp, lsi, loc, w, 64 if B % 64 == 0 else B)
    got = msda.ms_deform_attn_forward(value, shp, lsi, loc, w, 64)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,Lq,shapes,P,D", [(2, 900, A_SHAPES, 4, 32), (32, 900, A_SHAPES, 4, 32), (2, 57, [(6, 4), (3, 2)], 2, 30)])
def test_backward_matches_reference_cuda_kernel(ref, B, Lq, shapes, P, D):
    """grad_value is a sum of fp32 atomics in both kernels (order differs) -> tolerance, not bit equality"""
    from dtlr_b200 import msda
    value, shp, lsi, loc, w, go = _case(B, Lq, shapes, P, 12, D=D, M=8 if D == 32 else 2)
    want = ref.ms_deform_attn_backward(value, shp, lsi, loc, w, go, B)
    got = msda.ms_deform_attn_backward(value, shp, lsi, loc, w, go, 64)
    for g_, w_, atol in zip(got, want, (2e-4, 2e-3, 2e-4)):
        assert g_.shape == w_.shape
        assert torch.allclose(g_, w_, rtol=1e-3, atol=atol)
