"""The operator boundary as a DROP-IN: after `dtlr_b200.msda.install_as_reference_extension()` the
reference's own, unmodified caller code -- ops/functions/ms_deform_attn_func.py:18 `import MultiScaleDeformableAttention as MSDA`,
`MSDeformAttnFunction` (:21-38) and the `MSDeformAttn` module (ops/modules/ms_deform_attn.py:78-126) -- runs on dtlr_b200's kernels.
The reference files are staged (unmodified, git-ignored) by oracle/stage_ref_ops.py; tests skip when they were not staged."""
import sys

import pytest
import torch

from oracle import stage_ref_ops


@pytest.fixture(scope="module")
def ref_ops():
    from dtlr_b200 import msda
    msda.install_as_reference_extension()
    assert sys.modules["MultiScaleDeformableAttention"] is msda
    try:
        stage_ref_ops.stage()
    except RuntimeError:
        pass
    mods = stage_ref_ops.load()
    if mods is None:
        pytest.skip("oracle/_ref/ops not staged (python oracle/stage_ref_ops.py in the build container)")
    return mods


def test_reference_function_binds_to_our_module_and_fails_loudly_on_cpu(ref_ops):
    """CPU: the reference file imports OUR module under the reference's name, calls it with the pybind signature, and the product
    refuses CPU tensors instead of falling back"""
    from dtlr_b200 import _lib, msda
    modules, func = ref_ops
    assert func.MSDA is msda
    value = torch.rand(1, 30, 2, 4)
    shapes = torch.tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    loc = torch.rand(1, 2, 2, 2, 2, 2)
    w = torch.rand(1, 2, 2, 2, 2)
    with pytest.raises(_lib.DtlrError):
        func.MSDeformAttnFunction.apply(value, shapes, lsi, loc, w, 2)


@pytest.mark.gpu
def test_reference_module_forward_backward_through_our_kernels(ref_ops):
    """GPU: the reference's `MSDeformAttn` nn.Module (its own Linear layers, softmax and location arithmetic) with the core op bound
    to dtlr_msda_forward/backward, against the SAME module evaluated with the reference's `ms_deform_attn_core_pytorch`
    (ops/functions/ms_deform_attn_func.py:41-61) -- output and every gradient, encoder (2-coordinate) and decoder (4-coordinate)
    reference points, at the config-A level shapes."""
    modules, func = ref_ops
    torch.manual_seed(0)
    shapes_l = [(5, 128), (3, 64), (2, 32), (1, 16)]
    shapes = torch.tensor(shapes_l, dtype=torch.long, device="cuda")
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    for Lq, refdim in ((S, 2), (300, 4)):
        attn = modules.MSDeformAttn(d_model=256, n_levels=4, n_heads=8, n_points=4).cuda()
        with torch.no_grad():
            attn.sampling_offsets.weight.normal_(0, 0.02)
            attn.attention_weights.weight.normal_(0, 0.05)
        query = torch.randn(2, Lq, 256, device="cuda", requires_grad=True)
        src = torch.randn(2, S, 256, device="cuda", requires_grad=True)
        ref_pts = torch.rand(2, Lq, 4, refdim, device="cuda") * (0.8 if refdim == 2 else 0.4) + 0.1
        pad = torch.zeros(2, S, dtype=torch.bool, device="cuda")
        pad[1, -40:] = True
        out = attn(query, ref_pts, src, shapes, lsi, pad)
        g = torch.randn_like(out)
        grads = torch.autograd.grad(out, [query, src] + list(attn.parameters()), g)

        class _TorchCore:      # the reference's pure-torch core in the place of the autograd function, for the expected values
            @staticmethod
            def apply(value, shp, ls, loc, w, step):
                return func.ms_deform_attn_core_pytorch(value, shp.tolist(), loc, w)

        orig = modules.ms_deform_attn.MSDeformAttnFunction
        modules.ms_deform_attn.MSDeformAttnFunction = _TorchCore
        try:
            want = attn(query, ref_pts, src, shapes, lsi, pad)
            want_grads = torch.autograd.grad(want, [query, src] + list(attn.parameters()), g)
        finally:
            modules.ms_deform_attn.MSDeformAttnFunction = orig
        assert torch.allclose(out, want, rtol=1e-4, atol=1e-5), (out - want).abs().max()
        for a, b in zip(grads, want_grads):
            assert torch.allclose(a, b, rtol=1e-3, atol=1e-4 * float(b.abs().max())), (a - b).abs().max()
