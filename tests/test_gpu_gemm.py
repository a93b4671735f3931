"""GPU: the tcgen05 bf16 GEMM and the fp32 SIMT GEMM (dtlr_gemm) against torch fp32/fp64 matmul of the same operands
(floating-point kernel -> plain torch reference, tolerance stated per case)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K) -- SURVEY.md appendix C shapes (per image x small batch) + ragged tails
    (912 * 2, 256, 256), (900 * 2, 2048, 256), (912 * 2, 256, 2048), (900, 384, 256), (900, 768, 256),
    (900, 256, 512), (912, 166, 256), (900, 4, 256), (130, 64, 64), (1, 256, 256), (257, 200, 72), (640, 256, 512),
]


def _mk(M, N, K, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    return a, w, bias


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_bf16_tcgen05(M, N, K, out_dtype):
    from dtlr_b200 import ops
    if K % 8:
        pytest.skip("bf16 rows must be 16-byte aligned")
    a, w, bias = _mk(M, N, K, torch.bfloat16, M + N + K)
    ref = a.double() @ w.double().T + bias.double()
    out = ops.gemm(a, w, bias, out_dtype=out_dtype)
    torch.cuda.synchronize()
    tol = 2e-2 if out_dtype == torch.bfloat16 else 1e-4      # bf16 output rounding vs fp32 accumulate only
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, err
    # relu + residual epilogue
    res = torch.randn(M, N, device="cuda").to(out_dtype)
    out2 = ops.gemm(a, w, bias, residual=res, relu=True, out_dtype=out_dtype)
    ref2 = torch.relu(ref) + res.double()
    err2 = (out2.double() - ref2).abs().max().item() / ref2.abs().max().item()
    assert err2 < tol, err2


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_fp32_simt(M, N, K):
    from dtlr_b200 import ops
    a, w, bias = _mk(M, N, K, torch.float32, M * 3 + N + K)
    ref = a.double() @ w.double().T + bias.double()
    out = ops.gemm(a, w, bias)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err
    res = torch.randn(M, N, device="cuda")
    out2 = ops.gemm(a, w, None, residual=res, relu=True)
    ref2 = torch.relu(a.double() @ w.double().T) + res.double()
    assert (out2.double() - ref2).abs().max().item() / ref2.abs().max().item() < 1e-5


def test_strided_operands_and_output():
    """activations are often column slices of a wider buffer (fused QKV etc.)"""
    from dtlr_b200 import ops
    big = torch.randn(300, 1024, device="cuda").bfloat16()
    a = big[:, 256:512]
    w = (torch.randn(128, 256, device="cuda") / 16).bfloat16()
    outbuf = torch.zeros(300, 512, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, None, out=outbuf[:, 128:256])
    ref = a.float() @ w.float().T
    assert (outbuf[:, 128:256].float() - ref).abs().max() / ref.abs().max() < 2e-2
    assert outbuf[:, :128].abs().sum() == 0 and outbuf[:, 256:].abs().sum() == 0


@pytest.mark.parametrize("M,K", [(2700, 256), (912 * 3, 2048), (130, 256), (58368, 256)])
@pytest.mark.parametrize("with_res,with_add2", [(True, True), (True, False), (False, False)])
def test_gemm_with_fused_layernorm(M, K, with_res, with_add2):
    """dtlr_gemm_ln: Linear -> (+residual) -> LayerNorm(256) [-> y + add2] in one tcgen05 kernel vs torch fp32 on the same
    bf16 operands (the LN input is the bf16-rounded sum, exactly like the un-fused GEMM + LN kernels)."""
    import torch.nn.functional as F
    from dtlr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(256, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(256, device="cuda", generator=g)
    res = torch.randn(M, 256, device="cuda", generator=g).bfloat16() if with_res else None
    add2 = torch.randn(M, 256, device="cuda", generator=g).bfloat16() if with_add2 else None
    gamma = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    beta = 0.1 * torch.randn(256, device="cuda", generator=g)
    got = ops.gemm_ln(a, w, bias, res, gamma, beta, add2)
    y = got[0] if with_add2 else got
    x = (a.float() @ w.float().T + bias).bfloat16().float()
    if with_res:
        x = (x + res.float()).bfloat16().float()
    ref = F.layer_norm(x, (256,), gamma, beta, 1e-5)
    assert (y.float() - ref).abs().max().item() < 5e-2
    assert (y.float() - ref).abs().mean().item() < 4e-3
    if with_add2:
        ref2 = ref.bfloat16().float() + add2.float()
        assert (got[1].float() - ref2).abs().max().item() < 8e-2
    # and the composed un-fused path gives the same thing
    un = ops.add_layernorm(ops.gemm(a, w, bias, residual=res), None, gamma, beta)
    assert (y.float() - un.float()).abs().max().item() < 5e-2


WS_SHAPES = [  # (M, N, K): large enough that dtlr_gemm picks the weight-stationary kernel (K <= 256, >= 2 row tiles per CTA)
    (58368, 256, 256), (57600 + 77, 384, 256), (20000, 2048, 256), (57600, 512, 256), (163840, 256, 64), (80000, 64, 256),
    (50000, 128, 192), (45000, 256, 152), (60000, 1536, 256), (40960, 512, 128),
]


@pytest.mark.parametrize("M,N,K", WS_SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_weight_stationary_kernel(M, N, K, out_dtype):
    """gemm_ws_tcgen05_kernel (resident weight slice, A-only ring) vs fp64 matmul of the same bf16 operands, all epilogues
    (bias, ReLU before / after the residual add; TMA-store epilogue, TMA-loaded residual), ragged M and K tails; and
    identical to the tile kernel (dtlr_debug_flags(32) disables the weight-stationary path): same MMA order."""
    from dtlr_b200 import ops, _lib
    a, w, bias = _mk(M, N, K, torch.bfloat16, M + N + K)
    res = torch.randn(M, N, device="cuda").to(out_dtype)
    tol = 2e-2 if out_dtype == torch.bfloat16 else 1e-4
    ref = a.double() @ w.double().T + bias.double()
    cases = [dict(relu=0, residual=None), dict(relu=1, residual=res), dict(relu=2, residual=res)]
    outs = [ops.gemm(a, w, bias, out_dtype=out_dtype, **c) for c in cases]
    _lib.lib().dtlr_debug_flags(32)
    try:
        olds = [ops.gemm(a, w, bias, out_dtype=out_dtype, **c) for c in cases]
    finally:
        _lib.lib().dtlr_debug_flags(0)
    refs = [ref, torch.relu(ref) + res.double(), torch.relu(ref + res.double())]
    for c, o, old, r in zip(cases, outs, olds, refs):
        err = (o.double() - r).abs().max().item() / r.abs().max().item()
        assert err < tol, err
        if c["residual"] is None or out_dtype == torch.float32:
            assert torch.equal(o, old)
        else:   # the tile kernel rounds to bf16 before the residual add, this one adds in fp32 and rounds once
            assert (o.float() - old.float()).abs().max().item() <= 2 ** -7 * r.abs().max().item()


def _ffn_operands(M, hid):
    g = torch.Generator(device="cuda").manual_seed(M + hid)
    x = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(hid, 256, device="cuda", generator=g) / 16).bfloat16()
    b1 = 0.5 * torch.randn(hid, device="cuda", generator=g)
    w2 = (torch.randn(256, hid, device="cuda", generator=g) / hid ** 0.5).bfloat16()
    b2 = 0.5 * torch.randn(256, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    beta = 0.1 * torch.randn(256, device="cuda", generator=g)
    return x, w1, b1, w2, b2, gamma, beta


FFN_SHAPES = [(128, 2048), (300, 256), (5000, 1024), (58368, 2048), (57600 + 13, 2048), (37000, 2048), (50688, 2048), (40000, 1024),
              (148 * 256, 2048), (148 * 128, 2048), (19000, 256), (29184, 2048)]
# dtlr_debug_flags: 0 default plan (stream-K on CTA pairs from 74 pair tiles upwards), 1073741824 single-CTA kernels (stream-K with
# more than one round of tiles), 536870912 full rounds + split tail + tail kernel, 4096 CTA pairs that only multicast the weights
FFN_MODES = {"default": 0, "single": 1073741824, "split-tail": 1073741824 | 536870912, "multicast": 4096}


@pytest.mark.parametrize("M,hid", FFN_SHAPES)
@pytest.mark.parametrize("mode", list(FFN_MODES))
def test_fused_ffn_block(M, hid, mode):
    """dtlr_ffn_ln_ws: LN(x + W2 relu(W1 x + b1) + b2) in one tcgen05 kernel (hidden activation only in TMEM) vs torch fp32 on the
    same 16-bit operands with the hidden activation rounded to 16 bits (as every path does), vs the un-fused kernels (linear1 GEMM +
    linear2/LayerNorm GEMM), and -- every plan of the library -- vs the plain persistent kernel: rows of (pair) tiles that are not
    shared between two neighbouring CTAs (pairs) must be BIT-EQUAL (same chunk order), shared ones agree to fp32 re-association;
    repeated calls on the same workspace are bit-equal (ready flags handed back, fixed summation order)."""
    import torch.nn.functional as F
    from dtlr_b200 import ops, _lib
    x, w1, b1, w2, b2, gamma, beta = _ffn_operands(M, hid)
    lib = _lib.lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    tiles, nj = (M + 127) // 128, hid // 128
    saved, ops.FFN_FUSED = ops.FFN_FUSED, True
    try:
        lib.dtlr_debug_flags(262144)
        assert lib.dtlr_ffn_plan(M, hid) == 0
        y0 = ops.ffn_ln(x, w1, b1, w2, b2, gamma, beta)
        lib.dtlr_debug_flags(FFN_MODES[mode])
        plan = lib.dtlr_ffn_plan(M, hid)
        y = ops.ffn_ln(x, w1, b1, w2, b2, gamma, beta)
        for _ in range(3):
            assert torch.equal(ops.ffn_ln(x, w1, b1, w2, b2, gamma, beta), y)
    finally:
        ops.FFN_FUSED = saved
        lib.dtlr_debug_flags(0)
    if mode == "default":
        assert plan == (3 if (tiles + 1) // 2 >= sms // 2 else 0)
    elif mode == "single":
        assert plan == (2 if tiles > sms and tiles % sms else 0)
    elif mode == "multicast":
        assert plan == 0
    h = torch.relu(x.float() @ w1.float().T + b1).bfloat16().float()
    pre = (h @ w2.float().T + b2 + x.float()).bfloat16().float()
    ref = F.layer_norm(pre, (256,), gamma, beta, 1e-5)
    assert torch.isfinite(y).all()
    assert (y.float() - ref).abs().max().item() < 5e-2
    assert (y.float() - ref).abs().mean().item() < 4e-3
    un = ops.linear_ln(ops.gemm(x, w1, b1, relu=1), w2, b2, x, gamma, beta)
    assert (y.float() - un.float()).abs().max().item() < 5e-2
    # which rows may differ from the plain kernel
    keep = torch.ones(((tiles + 1) // 2) * 256, dtype=torch.bool, device="cuda")
    if plan == 3:
        groups, rows, n_t = sms // 2, 256, (tiles + 1) // 2
    elif plan == 2:
        groups, rows, n_t = sms, 128, tiles
    else:
        groups, rows, n_t = 0, 128, tiles
    if plan == 1:
        keep[(tiles // sms) * sms * 128:] = False
    elif groups:
        units = n_t * nj
        for c in range(1, groups):
            if (units * c // groups) % nj:
                t = (units * c // groups) // nj
                keep[t * rows:(t + 1) * rows] = False
    keep = keep[:M]
    assert torch.equal(y[keep], y0[keep])
    if (~keep).any():
        assert (y[~keep].float() - y0[~keep].float()).abs().max().item() < 5e-2
        assert (y[~keep].float() - y0[~keep].float()).abs().mean().item() < 2e-3


@pytest.mark.parametrize("M,N,K,out_dtype", [(57600 * 2, 166, 256, torch.float32), (58368, 166, 256, torch.float32),
                                              (60000, 100, 128, torch.bfloat16), (50000, 200, 256, torch.float32)])
def test_weight_stationary_ragged_slice_pitched_output(M, N, K, out_dtype):
    """class heads: N = 166 is no multiple of anything -- one ragged weight slice (rows beyond N zero-filled by TMA), output rows
    pitched to 16 bytes (166 -> 168 floats); the TMA store clips at 16-byte granularity, so the pad columns (padding by
    construction: pitch == N rounded up) receive zeros, nothing beyond the row is touched."""
    from dtlr_b200 import ops, _lib
    a, w, bias = _mk(M, N, K, torch.bfloat16, M + N + K)
    epc = 4 if out_dtype == torch.float32 else 8
    buf = torch.full((M, (N + epc - 1) // epc * epc), 7.0, device="cuda", dtype=out_dtype)
    out = ops.gemm(a, w, bias, out_dtype=out_dtype, out=buf[:, :N])
    ref = a.double() @ w.double().T + bias.double()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < (1e-4 if out_dtype == torch.float32 else 2e-2), err
    assert ((buf[:, N:] == 7.0) | (buf[:, N:] == 0.0)).all()
    _lib.lib().dtlr_debug_flags(32)
    try:
        old = ops.gemm(a, w, bias, out_dtype=out_dtype)
    finally:
        _lib.lib().dtlr_debug_flags(0)
    assert torch.equal(out, old)


@pytest.mark.parametrize("M,with_ref", [(57600, True), (128 * 148 + 77, True), (300, False), (58368, False)])
def test_fused_mlp_head_box_refine(M, with_ref):
    """dtlr_mlp_head (the FFN kernel's HEAD variant): 256 -> 256 -> 256 -> 4 MLP (reference models/dino/utils.py:110-122) + box refinement
    (deformable_transformer.py:734-738, util/misc.py:575-579) against torch fp32 on the same 16-bit operands (first hidden layer
    rounded to 16 bits as the kernel keeps it in TMEM, second one in fp32)."""
    from dtlr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M)
    x = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(256, 256, device="cuda", generator=g) / 16).bfloat16()
    w2 = (torch.randn(256, 256, device="cuda", generator=g) / 16).bfloat16()
    b1 = torch.randn(256, device="cuda", generator=g) * 0.1
    b2 = torch.randn(256, device="cuda", generator=g) * 0.1
    w3 = torch.randn(4, 256, device="cuda", generator=g) * 0.05
    b3 = torch.randn(4, device="cuda", generator=g) * 0.1
    ref = torch.rand(M, 4, device="cuda", generator=g) if with_ref else None
    if with_ref:
        ref[::7] = 0.0            # clamped by eps = 1e-3 on both sides
        ref[3::11] = 1.0
    got = ops.mlp_head(x, (w1, b1), (w2, b2), w3, b3, ref)
    h1 = torch.relu(x.float() @ w1.float().T + b1).bfloat16().float()
    h2 = torch.relu(h1 @ w2.float().T + b2)
    want = h2 @ w3.T + b3
    if with_ref:
        r = ref.clamp(0, 1)
        want = (want + torch.log(r.clamp(min=1e-3) / (1 - r).clamp(min=1e-3))).sigmoid()
    assert got.shape == (M, 4) and got.dtype == torch.float32
    assert torch.allclose(got, want, rtol=2e-3, atol=2e-3), (got - want).abs().max()
