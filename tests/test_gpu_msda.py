"""GPU: the sm_100a MSDA kernels through the C ABI against (a) the committed reference known-answer vectors
(reference ops/test.py fixture) and (b) the C oracle on seeded inputs, including out-of-range points, empty query
sets, a slab too large for shared memory and the full BASELINE config-2 size."""
import os

import numpy as np
import pytest
import torch

from oracle import msda as omsda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "msda_kat.npz"))


def _levels(shapes):
    shp = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shp.new_zeros((1,)), shp.prod(1).cumsum(0)[:-1]))
    return shp, lsi


def _rand_case(B, shapes, M, D, Lq, P, seed, dtype=torch.float32, oob=True):
    g = torch.Generator().manual_seed(seed)
    shp, lsi = _levels(shapes)
    S = int(shp.prod(1).sum())
    L = len(shapes)
    value = torch.randn(B, S, M, D, generator=g).to(dtype)
    loc = torch.rand(B, Lq, M, L, P, 2, generator=g)
    if oob:
        loc = loc * 1.3 - 0.15
    w = torch.softmax(torch.randn(B, Lq, M, L * P, generator=g), -1).view(B, Lq, M, L, P)
    return value, shp, lsi, loc.to(dtype), w.to(dtype)


@pytest.mark.parametrize("D", [2, 30, 32, 64, 71])
def test_forward_double_reference_kat(kat, D):
    from dtlr_b200 import msda
    t = lambda k: torch.from_numpy(kat["D%d_%s" % (D, k)]).cuda()
    shp, lsi = torch.from_numpy(kat["shapes"]).cuda(), torch.from_numpy(kat["lsi"]).cuda()
    out = msda.ms_deform_attn_forward(t("value"), shp, lsi, t("loc"), t("w"), 2)
    assert torch.allclose(out.cpu(), torch.from_numpy(kat["D%d_out" % D]))          # reference test.py:40


@pytest.mark.parametrize("D", [2, 32])
def test_forward_float_reference_kat(kat, D):
    from dtlr_b200 import msda
    t = lambda k: torch.from_numpy(kat["D%d_%s" % (D, k)]).float().cuda()
    shp, lsi = torch.from_numpy(kat["shapes"]).cuda(), torch.from_numpy(kat["lsi"]).cuda()
    out = msda.ms_deform_attn_forward(t("value"), shp, lsi, t("loc"), t("w"), 2)
    ref = torch.from_numpy(kat["D%d_out" % D])
    assert torch.allclose(out.cpu().double(), ref, rtol=1e-2, atol=1e-3)               # reference test.py:56
    assert torch.allclose(out.cpu().double(), ref, rtol=1e-5, atol=1e-8)               # and much tighter


@pytest.mark.parametrize("D", [2, 30, 32, 64, 71])
def test_backward_double_reference_kat(kat, D):
    from dtlr_b200 import msda
    t = lambda k: torch.from_numpy(kat["D%d_%s" % (D, k)]).cuda()
    shp, lsi = torch.from_numpy(kat["shapes"]).cuda(), torch.from_numpy(kat["lsi"]).cuda()
    gv, gl, ga = msda.ms_deform_attn_backward(t("value"), shp, lsi, t("loc"), t("w"), t("gout"), 2)
    assert torch.allclose(gv.cpu(), torch.from_numpy(kat["D%d_gvalue" % D]), rtol=1e-9, atol=1e-13)
    assert torch.allclose(gl.cpu(), torch.from_numpy(kat["D%d_gloc" % D]), rtol=1e-9, atol=1e-13)
    assert torch.allclose(ga.cpu(), torch.from_numpy(kat["D%d_gw" % D]), rtol=1e-9, atol=1e-13)


def test_autograd_function_gradcheck():
    """reference test.py:63-78 (gradcheck in fp64 through MSDeformAttnFunction)."""
    from dtlr_b200 import msda
    value, shp, lsi, loc, w = _rand_case(1, [(6, 4), (3, 2)], 2, 32, 2, 2, 5, torch.float64, oob=False)
    value = (value * 0.01).cuda().requires_grad_()
    loc = loc.cuda().requires_grad_()
    w = w.cuda().requires_grad_()
    assert torch.autograd.gradcheck(msda.MSDeformAttnFunction.apply, (value, shp.cuda(), lsi.cuda(), loc, w, 2))


A_SHAPES = [(5, 128), (3, 64), (2, 32), (1, 16)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Lq", [1, 37, 900])
def test_fast_path_vs_oracle(dtype, Lq):
    from dtlr_b200 import msda
    value, shp, lsi, loc, w = _rand_case(2, A_SHAPES, 8, 32, Lq, 4, 100 + Lq)
    vq = value.to(dtype)
    ref = omsda.msda_forward(vq.float(), shp, lsi, loc, w)           # oracle on the same (rounded) values
    out = msda.ms_deform_attn_forward(vq.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), 64).float().cpu()
    if dtype == torch.float32:
        assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)
    else:   # bf16 output rounding only (fp32 accumulation inside)
        assert torch.allclose(out, ref, rtol=1e-2, atol=1e-2)


def test_fast_path_committed_vector(kat):
    from dtlr_b200 import msda
    g = torch.Generator().manual_seed(11)
    value = torch.randn(2, 912, 8, 32, generator=g)
    loc = torch.rand(2, 37, 8, 4, 4, 2, generator=g) * 1.3 - 0.15
    w = torch.softmax(torch.randn(2, 37, 8, 16, generator=g), -1).view(2, 37, 8, 4, 4)
    shp, lsi = torch.from_numpy(kat["A_shapes"]), torch.from_numpy(kat["A_lsi"])
    out = msda.ms_deform_attn_forward(value.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), 64).cpu()
    assert torch.allclose(out, torch.from_numpy(kat["A_out"]), rtol=1e-4, atol=1e-5)


def test_points_on_borders_and_far_outside():
    from dtlr_b200 import msda
    value, shp, lsi, loc, w = _rand_case(1, A_SHAPES, 8, 32, 64, 4, 7)
    # exact borders, half-pixel positions, far outside, negative
    special = torch.tensor([0.0, 1.0, -1e-7, 1.0 + 1e-7, 0.5 / 128, 1 - 0.5 / 128, -5.0, 7.0, 1.5 / 5, 0.999999])
    idx = torch.randint(0, len(special), loc.shape, generator=torch.Generator().manual_seed(1))
    loc = special[idx]
    ref = omsda.msda_forward(value, shp, lsi, loc, w)
    out = msda.ms_deform_attn_forward(value.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), 64).cpu()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)


def test_empty_query_set_and_batch():
    from dtlr_b200 import msda
    value, shp, lsi, loc, w = _rand_case(2, A_SHAPES, 8, 32, 4, 4, 3)
    out = msda.ms_deform_attn_forward(value.cuda(), shp.cuda(), lsi.cuda(), loc[:, :0].contiguous().cuda(),
                                      w[:, :0].contiguous().cuda(), 64)
    assert out.shape == (2, 0, 256)


def test_large_map_gathers_from_global():
    """real IAM resolution (~94x1333 -> S=2676+): the fp32 slab (342 KB) does not fit shared memory."""
    from dtlr_b200 import msda
    shapes = [(12, 167), (6, 84), (3, 42), (2, 21)]
    for dtype in (torch.float32, torch.bfloat16):
        value, shp, lsi, loc, w = _rand_case(1, shapes, 8, 32, 200, 4, 9)
        vq = value.to(dtype)
        ref = omsda.msda_forward(vq.float(), shp, lsi, loc, w)
        out = msda.ms_deform_attn_forward(vq.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), 64).float().cpu()
        tol = 1e-4 if dtype == torch.float32 else 1e-2
        assert torch.allclose(out, ref, rtol=tol, atol=tol * 0.1 if dtype == torch.float32 else 1e-2)


def test_generic_path_other_head_width():
    from dtlr_b200 import msda
    value, shp, lsi, loc, w = _rand_case(2, [(6, 4), (3, 2)], 2, 30, 5, 2, 13)
    ref = omsda.msda_forward(value, shp, lsi, loc, w)
    out = msda.ms_deform_attn_forward(value.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), 64).cpu()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)


def test_backward_fp32_vs_oracle_config_size():
    from dtlr_b200 import msda
    value, shp, lsi, loc, w = _rand_case(2, A_SHAPES, 8, 32, 50, 4, 21)
    go = torch.randn(2, 50, 256, generator=torch.Generator().manual_seed(2))
    rv, rl, ra = omsda.msda_backward(value.double(), shp, lsi, loc.double(), w.double(), go.double())
    gv, gl, ga = msda.ms_deform_attn_backward(value.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), go.cuda(), 64)
    assert torch.allclose(gv.cpu().double(), rv, rtol=1e-3, atol=1e-4)
    assert torch.allclose(gl.cpu().double(), rl, rtol=1e-3, atol=1e-3)
    assert torch.allclose(ga.cpu().double(), ra, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("shapes,P", [(A_SHAPES, 4), ([(5, 128), (3, 64)], 8), ([(7, 33)], 16),
                                      ([(4, 9), (3, 7), (2, 5), (2, 3), (1, 4), (1, 3), (1, 2), (1, 1)], 2)])
def test_backward_fast_path_vs_oracle_and_generic_kernel(shapes, P):
    """D = 32 fp32 with L*P = 16 runs msda_bwd_d32_kernel (vector reductions, transposed shuffle reduction): against the fp64
    C oracle, and against the generic one-warp-per-item kernel (debug flag 8192) on the same inputs, with points on the map
    borders / outside and a ragged warp count."""
    from dtlr_b200 import _lib, msda
    B, Lq, M = 3, 37, 8
    value, shp, lsi, loc, w = _rand_case(B, shapes, M, 32, Lq, P, 31 + P)
    special = torch.tensor([0.0, 1.0, -1e-7, 1.0 + 1e-7, 0.5 / 128, 1 - 0.5 / 128, -5.0, 7.0, 0.3, 0.999999])
    g = torch.Generator().manual_seed(5)
    pick = torch.rand(loc.shape, generator=g) < 0.25
    loc = torch.where(pick, special[torch.randint(0, len(special), loc.shape, generator=g)], loc)
    go = torch.randn(B, Lq, M * 32, generator=g)
    rv, rl, ra = omsda.msda_backward(value.double(), shp, lsi, loc.double(), w.double(), go.double())
    args = (value.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), go.cuda(), 64)
    gv, gl, ga = msda.ms_deform_attn_backward(*args)
    _lib.lib().dtlr_debug_flags(8192)
    try:
        gv0, gl0, ga0 = msda.ms_deform_attn_backward(*args)
    finally:
        _lib.lib().dtlr_debug_flags(0)
    for new, old, ref, atol in ((gv, gv0, rv, 1e-4), (gl, gl0, rl, 1e-3), (ga, ga0, ra, 1e-4)):
        assert torch.isfinite(new).all()
        assert torch.allclose(new.cpu().double(), ref, rtol=1e-3, atol=atol)
        assert torch.allclose(new, old, rtol=1e-3, atol=atol)


def test_backward_full_size_matches_generic_kernel_and_is_linear_in_grad_out():
    """BASELINE config-5 shape per GPU (B = 32, 900 queries): fast vs generic kernel, and linearity in grad_out
    (size-independent property: bwd(a*g1 + g2) == a*bwd(g1) + bwd(g2))."""
    from dtlr_b200 import _lib, msda
    B, Lq = 32, 900
    value, shp, lsi, loc, w = _rand_case(B, A_SHAPES, 8, 32, Lq, 4, 77)
    g = torch.Generator().manual_seed(6)
    g1, g2 = torch.randn(B, Lq, 256, generator=g).cuda(), torch.randn(B, Lq, 256, generator=g).cuda()
    a = (value.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda())
    r1 = msda.ms_deform_attn_backward(*a, g1, 64)
    r2 = msda.ms_deform_attn_backward(*a, g2, 64)
    r12 = msda.ms_deform_attn_backward(*a, 0.5 * g1 + g2, 64)
    for x1, x2, x12 in zip(r1, r2, r12):
        assert torch.allclose(x12, 0.5 * x1 + x2, rtol=1e-3, atol=2e-3)
    _lib.lib().dtlr_debug_flags(8192)
    try:
        r1g = msda.ms_deform_attn_backward(*a, g1, 64)
    finally:
        _lib.lib().dtlr_debug_flags(0)
    for x, y in zip(r1, r1g):
        assert torch.allclose(x, y, rtol=1e-3, atol=2e-3)


def test_full_size_linearity_config2():
    """BASELINE config 2 size (B=64, S=912, Lq=912): size-independent property -- the op is linear in value and in the
    attention weights: f(a*v1 + v2, w) == a*f(v1,w) + f(v2,w)."""
    from dtlr_b200 import msda
    value, shp, lsi, loc, w = _rand_case(64, A_SHAPES, 8, 32, 912, 4, 33)
    v1 = value.cuda()
    v2 = torch.randn_like(v1)
    shp, lsi, loc, w = shp.cuda(), lsi.cuda(), loc.cuda(), w.cuda()
    f = lambda v: msda.ms_deform_attn_forward(v, shp, lsi, loc, w, 64)
    lhs = f(2.5 * v1 + v2)
    rhs = 2.5 * f(v1) + f(v2)
    assert torch.allclose(lhs, rhs, rtol=1e-4, atol=1e-4)
    # and a slice of it against the oracle
    ref = omsda.msda_forward(value[:1], shp.cpu(), lsi.cpu(), loc[:1].cpu(), w[:1].cpu())
    assert torch.allclose(f(v1)[:1].cpu(), ref, rtol=1e-4, atol=1e-5)


def test_rejects_cpu_tensors():
    from dtlr_b200 import msda, _lib
    value, shp, lsi, loc, w = _rand_case(1, [(6, 4), (3, 2)], 2, 32, 2, 2, 1)
    with pytest.raises(_lib.DtlrError):
        msda.ms_deform_attn_forward(value, shp, lsi, loc, w, 64)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ref_dim", [2, 4])
def test_fused_prologue_matches_prep_plus_core(dtype, ref_dim):
    """dtlr_msda_forward_fused == dtlr_msda_prep + dtlr_msda_forward (softmax + sampling locations folded in),
    and both follow reference ms_deform_attn.py:98-108 (checked against plain torch here)."""
    from dtlr_b200 import msda, ops, _lib
    B, Lq, M, L, P = 3, 77, 8, 4, 4
    g = torch.Generator().manual_seed(5 + ref_dim)
    shp, lsi = _levels(A_SHAPES)
    S = 912
    value = torch.randn(B, S, M, 32, generator=g).to(dtype).cuda()
    proj = torch.randn(B * Lq, 384, generator=g).cuda()
    proj[:, :256] *= 2.0
    ref = torch.rand(B * Lq, ref_dim, generator=g).cuda()
    if ref_dim == 4:
        ref[:, 2:] *= 0.3
    vr = (0.5 + 0.5 * torch.rand(B, L, 2, generator=g)).cuda()
    sh, ls, n = msda._host_levels(shp.cuda(), lsi.cuda())
    fused = msda.msda_forward_fused(value, sh, ls, n, proj, ref, vr, Lq, P).float()
    loc, attn = ops.msda_prep(proj, ref, vr, sh, n, B, Lq, M, P)
    two = msda.msda_forward_raw(value, sh, ls, n, loc, attn).float()
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert torch.allclose(fused, two, rtol=tol, atol=tol)
    if dtype == torch.bfloat16:      # throughput mode feeds the projection rows in bf16 as well
        pb = proj.bfloat16()
        fused_b = msda.msda_forward_fused(value, sh, ls, n, pb, ref, vr, Lq, P).float()
        loc_b, attn_b = ops.msda_prep(pb.float(), ref, vr, sh, n, B, Lq, M, P)
        two_b = msda.msda_forward_raw(value, sh, ls, n, loc_b, attn_b).float()
        assert torch.allclose(fused_b, two_b, rtol=tol, atol=tol)
    # value given as a column block of a wider matrix (row pitch 3*256), as the engine's batched value projection does
    wide = torch.randn(B * S, 3 * M * 32, device="cuda").to(dtype)
    wide[:, 256:512] = value.view(B * S, 256)
    blk = wide[:, 256:512].unflatten(0, (B, S)).unflatten(2, (M, 32))
    assert torch.equal(msda.msda_forward_fused(blk, sh, ls, n, proj, ref, vr, Lq, P).float(), fused)
    # plain torch statement of the prologue
    off = proj[:, :256].view(B, Lq, M, L, P, 2)
    aw = torch.softmax(proj[:, 256:].view(B, Lq, M, L * P), -1).view(B, Lq, M, L, P)
    r = ref.view(B, Lq, 1, ref_dim) * torch.cat([vr, vr], -1)[:, None, :, :ref_dim]
    if ref_dim == 2:
        norm = torch.stack([shp[:, 1], shp[:, 0]], -1).float().cuda()
        loc_t = r[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc_t = r[:, :, None, :, None, :2] + off / P * r[:, :, None, :, None, 2:] * 0.5
    assert torch.allclose(loc, loc_t, rtol=1e-5, atol=1e-6) and torch.allclose(attn, aw, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_tensor_core_gather_kernel_matches_simt_kernel_full_size(ref_dim):
    """bf16 throughput mode at BASELINE config-2 size: msda_fwd_mma_kernel (ldmatrix + mma.sync gather, the default) against
    msda_fwd_d32_kernel (SIMT, selected with dtlr_debug_flags(16)) on the same inputs, ragged valid ratios included.
    The two differ only by the bf16 rounding of the 64 corner weights (fp32 accumulation in both)."""
    from dtlr_b200 import msda, _lib
    B, Lq, M, L, P, S = 64, 900 if ref_dim == 4 else 912, 8, 4, 4, 912
    g = torch.Generator(device="cuda").manual_seed(17 + ref_dim)
    shp, lsi = _levels(A_SHAPES)
    sh, ls, n = msda._host_levels(shp.cuda(), lsi.cuda())
    value = torch.randn(B, S, M, 32, device="cuda", generator=g).bfloat16()
    proj = torch.randn(B * Lq, 384, device="cuda", generator=g)
    proj[:, :256] *= 3.0                                    # offsets large enough to leave the map at every level
    proj = proj.bfloat16()
    ref = torch.rand(B * Lq, ref_dim, device="cuda", generator=g)
    if ref_dim == 4:
        ref[:, 2:] *= 0.3
    vr = 0.5 + 0.5 * torch.rand(B, L, 2, device="cuda", generator=g)
    new = msda.msda_forward_fused(value, sh, ls, n, proj, ref, vr, Lq, P).float()
    _lib.lib().dtlr_debug_flags(16)
    try:
        old = msda.msda_forward_fused(value, sh, ls, n, proj, ref, vr, Lq, P).float()
    finally:
        _lib.lib().dtlr_debug_flags(0)
    assert torch.isfinite(new).all()
    assert torch.allclose(new, old, rtol=2e-2, atol=2e-2)
    assert (new - old).abs().mean().item() < 2e-3
