"""GPU: the fused inference engine (every compute step a libdtlr_b200 kernel through the C ABI) against the committed
reference vectors.  fp32 parity mode: 1e-3 relative-to-max on logits/boxes (north star) and bit-identical argmax
character sequences; bf16 throughput mode: reported agreement with documented looser bounds."""
import numpy as np
import pytest
import torch

from dtlr_b200 import dino, synth
from gpu_common import build_model, fixture, near_tie_mask, rel

pytestmark = pytest.mark.gpu
TOL = 1e-3
MODE_TOL = 5e-2     # documented bound of the 16-bit-operand throughput mode with the ranking forced; measured ~2e-2


def run_engine(model, x, force=None, dtype=torch.float32, split=False):
    model.eval()
    model.use_engine = True
    model.compute_dtype = dtype
    model.split_precision = split
    st = {}
    model.transformer.debug_force_topk = force
    from dtlr_b200.misc import nested_tensor_from_tensor_list
    try:
        with torch.no_grad():
            out = model.engine().forward(nested_tensor_from_tensor_list(x), stages=st)
    finally:
        model.transformer.debug_force_topk = None
        model.split_precision = False
    return out, st


def check(fx, out, st, forced, tol=TOL):
    assert rel(st["feats"][0][0][:, ::16, :, ::8], fx["feat_c3_s"]) < tol
    assert rel(st["feats"][2][0][:, ::32], fx["feat_c5_s"]) < tol
    assert rel(st["memory"][:, ::8, ::4], fx["memory_s"]) < tol
    assert rel(st["topk_scores"], fx["topk_scores"]) < tol
    if not forced:
        mism = st["topk_idx"].cpu().numpy() != fx["topk_idx"]
        assert near_tie_mask(fx)[mism].all()
    assert rel(st["hs"][0][:, ::8, ::4], fx["hs0_s"]) < tol
    assert rel(st["refs"][1], fx["ref1"]) < tol
    assert rel(out["pred_logits"], fx["pred_logits"]) < tol
    assert rel(out["pred_boxes"], fx["pred_boxes"]) < tol
    assert rel(out["aux_outputs"][4]["pred_logits"][:, ::8, ::4], fx["aux4_logits_s"]) < tol
    assert rel(out["aux_outputs"][0]["pred_boxes"], fx["aux0_boxes"]) < tol
    assert rel(out["interm_outputs"]["pred_logits"][:, ::8, ::4], fx["interm_logits_s"]) < tol
    assert rel(out["interm_outputs"]["pred_boxes"], fx["interm_boxes"]) < tol
    assert rel(out["interm_outputs_for_matching_pre"]["pred_boxes"], fx["init_box_proposal"]) < tol


def assert_identical_sequences_where_decidable(new, fx):
    """argmax character frames must be bit-identical to the reference's, except at frames whose decision the reference
    itself only takes by a margin below fp32 round-off of the path (random synthetic weights produce such ties; on
    trained weights there are none): two x-sorted neighbours with |cx_i - cx_j| < 2e-6, or a blank-vs-character /
    character-vs-character top-2 margin < 2e-6.  Everything else -- in practice every frame -- must match exactly."""
    from oracle import dino_ref
    ref_logits, ref_boxes = torch.from_numpy(fx["pred_logits"]), torch.from_numpy(fx["pred_boxes"])
    ref_new, idx = dino_ref.ctc_view(ref_logits, ref_boxes)
    cx = torch.gather(ref_boxes[:, :, 0], 1, idx)
    gap = torch.minimum(torch.diff(cx, dim=1, prepend=cx[:, :1] - 1), torch.diff(cx, dim=1, append=cx[:, -1:] + 1))
    top2 = ref_new.topk(2, dim=-1)[0]
    undecidable = (gap < 2e-6) | ((top2[..., 0] - top2[..., 1]) < 2e-6)
    mism = new.argmax(-1).cpu() != torch.from_numpy(fx["ctc_argmax"]).long()
    print("frames %d, mismatching %d, undecidable in the reference %d" % (mism.numel(), int(mism.sum()), int(undecidable.sum())))
    assert not (mism & ~undecidable).any()
    assert mism.float().mean() < 0.01


def test_fp32_config1_single_line_100_queries():
    fx = fixture("dino_P_b1")
    model, _, _ = build_model(100)
    out, st = run_engine(model, synth.synth_images(1, 40, 704, seed=1).cuda())
    assert (st["topk_idx"].cpu().numpy() == fx["topk_idx"]).all()
    check(fx, out, st, forced=False)


def test_fp32_ragged_batch():
    fx = fixture("dino_R_b3")
    model, _, _ = build_model(300)
    imgs = [t.cuda() for t in synth.synth_images(3, 40, 1024, seed=2, widths=fx["widths"].tolist())]
    out, st = run_engine(model, imgs, force=torch.from_numpy(fx["topk_idx"]).long())
    check(fx, out, st, forced=True)


def test_fp32_config2_shape_and_identical_character_sequences():
    fx = fixture("dino_A_b2")
    model, crit, _ = build_model(900)
    x = synth.synth_images(2, 40, 1024, seed=0).cuda()
    out, st = run_engine(model, x, force=torch.from_numpy(fx["topk_idx"]).long())
    check(fx, out, st, forced=True)
    new = dino.ctc_view(out["pred_logits"], out["pred_boxes"])
    assert_identical_sequences_where_decidable(new, fx)
    # model(...) in eval/no_grad dispatches to the engine and returns every key the reference returns
    model.transformer.debug_force_topk = torch.from_numpy(fx["topk_idx"]).long()
    with torch.no_grad():
        out2 = model(x)
    model.transformer.debug_force_topk = None
    assert set(out2) == {"pred_logits", "pred_boxes", "aux_outputs", "interm_outputs", "interm_outputs_for_matching_pre", "dn_meta"}
    assert len(out2["aux_outputs"]) == 5 and torch.equal(out2["pred_logits"], out["pred_logits"])


def _edit_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


@pytest.mark.parametrize("dt,bound,min_agree", [(torch.float16, 1e-2, 0.97), (torch.bfloat16, 5e-2, 0.90)])
def test_throughput_mode_agreement_with_reference_fixture(dt, bound, min_agree):
    """16-bit operands / activations, fp32 accumulation, against the reference-generated fixture.  Not 1e-3 modes: with
    the reference ranking forced the documented bounds are, relative-to-max on logits / boxes, 1e-2 for fp16 (measured ~3e-3) and
    5e-2 for bf16 (measured on B200: 2.2e-2 / 1.7e-2, 94.8 % of queries with the same blank/character decision); the character error
    rate of the decoded lines against the fp32 reference decode is printed (random weights put many decisions and cx orderings at
    near-ties, so this is a pessimistic stand-in for trained weights)."""
    fx = fixture("dino_A_b2")
    model, crit, _ = build_model(900)
    x = synth.synth_images(2, 40, 1024, seed=0).cuda()
    out, st = run_engine(model, x, force=torch.from_numpy(fx["topk_idx"]).long(), dtype=dt)
    e_log, e_box = rel(out["pred_logits"], fx["pred_logits"]), rel(out["pred_boxes"], fx["pred_boxes"])
    ref_logits, ref_boxes = torch.from_numpy(fx["pred_logits"]), torch.from_numpy(fx["pred_boxes"])

    def per_query_decision(logits):
        p = logits.sigmoid()
        s = p.sum(-1)
        blank = torch.where(s < 1 - 0.003, 1 - s, torch.full_like(s, 0.003))
        pm, arg = p.max(-1)
        pm = torch.where(s < 1 - 0.003, pm, (1 - 0.003) * pm / s)
        return torch.where(blank >= pm, torch.full_like(arg, -1), arg)

    dec = per_query_decision(out["pred_logits"].float().cpu())
    dec_ref = per_query_decision(ref_logits)
    agree = (dec == dec_ref).float().mean().item()
    seq = dino.convert_output_to_pred(dino.ctc_view(out["pred_logits"].float(), out["pred_boxes"].float()))
    seq_ref = dino.convert_output_to_pred(dino.ctc_view(ref_logits, ref_boxes))
    cer = sum(_edit_distance(a, b) for a, b in zip(seq, seq_ref)) / max(1, sum(len(b) for b in seq_ref))
    print("%s: logits %.3e boxes %.3e memory %.3e per-query decision agreement %.4f CER-vs-fp32-decode %.4f" % (
        dt, e_log, e_box, rel(st["memory"][:, ::8, ::4].float(), fx["memory_s"]), agree, cer))
    assert e_log < bound and e_box < bound and agree >= min_agree


def _oracle_run(model, x, cfg_kw):
    """the CPU oracle (oracle/dino_ref.py, pinned to the reference-generated fixtures by tests/test_oracle_dino.py) on the model's
    own weights -- runs on the host cores of the GPU box (about 15 images/s)"""
    from oracle import dino_ref
    sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
    st = {}
    ref = dino_ref.dino_forward(sd, dino_ref.default_cfg(**cfg_kw), x, stages=st)
    return ref, st


def _decidable_frames(ref, out, cx_margin, p_margin):
    """frames of the CTC view (reference dino.py:472-502 + engine.py:523-529) that the reference decides by a margin: neighbouring
    cx gap >= cx_margin (the reading order) and top-2 probability gap >= p_margin (the label).
    Returns (#frames, #decidable, #decidable mismatching)"""
    from oracle import dino_ref
    ref_new, idx = dino_ref.ctc_view(ref["pred_logits"], ref["pred_boxes"])
    cx = torch.gather(ref["pred_boxes"][:, :, 0], 1, idx)
    gap = torch.minimum(torch.diff(cx, dim=1, prepend=cx[:, :1] - 1), torch.diff(cx, dim=1, append=cx[:, -1:] + 1))
    top2 = ref_new.topk(2, dim=-1)[0]
    decidable = (gap >= cx_margin) & ((top2[..., 0] - top2[..., 1]) >= p_margin)
    mine = dino.decode_frames(out).long().cpu()
    mism = mine != ref_new.argmax(-1)
    return mism.numel(), int(decidable.sum()), int((mism & decidable).sum())


def _query_decisions(logits, eps=0.003):
    """blank-or-class decision of every query BEFORE the cx sort, with the top-2 margin of the CTC-view probabilities"""
    p = logits.sigmoid()
    s = p.sum(-1, keepdim=True)
    low = s < 1 - eps
    new = torch.cat([torch.where(low, 1 - s, torch.full_like(s, eps)), torch.where(low, p, (1 - eps) * p / s)], -1)
    top2, arg = new.topk(2, dim=-1)
    return arg[..., 0], top2[..., 0] - top2[..., 1]


@pytest.fixture(scope="module")
def bench_shape():
    """BASELINE config 2 at the shape bench.py times: 64 x 3x40x1024, Q = 900, C = 166, the bench's own synthetic batch"""
    model, _, _ = build_model(900)
    x = synth.synth_images(64, 40, 1024, seed=100)
    ref, st = _oracle_run(model, x, dict(num_queries=900))
    return model, x, ref, st


def test_bench_shape_fp32_vs_oracle(bench_shape):
    """parity AT the bench shape (B = 64 selects other tile shapes / persistent-grid schedules / weight-stationary
    slicing than the B <= 3 fixtures): fp32 mode within 1e-3 of the oracle on every stage and output, rankings identical wherever the
    reference's score gap is decidable, identical character frames, and CUDA-graph replay == eager launches."""
    model, x, ref, rst = bench_shape
    xg = x.cuda()
    out, st = run_engine(model, xg, force=rst["topk_idx"])
    for a, b, name in ((st["memory"], rst["memory"], "memory"), (st["topk_scores"], rst["topk_scores"], "scores"),
                       (st["hs"][5], rst["hs"][5], "hs5"), (out["pred_logits"], ref["pred_logits"], "logits"),
                       (out["pred_boxes"], ref["pred_boxes"], "boxes"),
                       (out["aux_outputs"][0]["pred_logits"], ref["aux_outputs"][0]["pred_logits"], "aux0 logits"),
                       (out["interm_outputs"]["pred_boxes"], ref["interm_outputs"]["pred_boxes"], "interm boxes")):
        e = rel(a.float(), b)
        print("B=64 fp32 %-12s rel-to-max %.2e" % (name, e))
        assert e < TOL, name
    n, dec, bad = _decidable_frames(ref, out, 1e-5, 1e-4)
    print("B=64 fp32 frames %d decidable (cx gap >= 1e-5, top-2 gap >= 1e-4) %d mismatching-decidable %d" % (n, dec, bad))
    assert bad == 0 and dec > 0.9 * n
    # un-forced ranking: every mismatch sits on a reference near-tie
    out_u, st_u = run_engine(model, xg)
    sc = torch.gather(rst["topk_scores"], 1, rst["topk_idx"]).numpy()
    tie = np.minimum(np.abs(np.diff(sc, axis=1, prepend=np.inf)), np.abs(np.diff(sc, axis=1, append=-np.inf))) < 1e-4
    mism = st_u["topk_idx"].cpu().numpy() != rst["topk_idx"].numpy()
    print("B=64 fp32 un-forced ranking: %d of %d positions differ, all on near-ties: %s" % (mism.sum(), mism.size, bool(tie[mism].all())))
    assert tie[mism].all() and mism.mean() < 0.05
    # CUDA-graph replay (what bench.py times) returns what the eager launches return
    model.use_cuda_graph = True
    try:
        with torch.no_grad():
            model(xg)
            og = model(xg)
        assert torch.equal(og["pred_logits"], out_u["pred_logits"]) and torch.equal(og["pred_boxes"], out_u["pred_boxes"])
    finally:
        model.use_cuda_graph = False


def test_split_precision_mode_config2_fixture_within_1e3():
    """the tensor-core parity mode (model.split_precision: fp32 activations, every Linear / conv a 3-term fp16 split product on tcgen05,
    fp16 attention core, exact fp32 deformable-attention core) against the reference-generated fixture: EVERY compared stage and output
    within the north-star 1e-3 (predicted by tests/precision_sim.py: logits 1.7e-4, boxes 1.1e-4), frames identical wherever the
    reference decides by a margin."""
    fx = fixture("dino_A_b2")
    model, _, _ = build_model(900)
    x = synth.synth_images(2, 40, 1024, seed=0).cuda()
    out, st = run_engine(model, x, force=torch.from_numpy(fx["topk_idx"]).long(), split=True)
    check(fx, out, st, forced=True)
    e_log, e_box = rel(out["pred_logits"], fx["pred_logits"]), rel(out["pred_boxes"], fx["pred_boxes"])
    ref = {"pred_logits": torch.from_numpy(fx["pred_logits"]), "pred_boxes": torch.from_numpy(fx["pred_boxes"])}
    n, dec, bad = _decidable_frames(ref, out, 1e-3, 1e-2)
    # un-forced two-stage ranking: mismatches only where the reference's score gap is below the mode's score accuracy
    out_u, st_u = run_engine(model, x, split=True)
    mism = st_u["topk_idx"].cpu().numpy() != fx["topk_idx"]
    ref_scores = np.take_along_axis(fx["topk_scores"], fx["topk_idx"].astype(np.int64), 1)
    tie = np.minimum(np.abs(np.diff(ref_scores, axis=1, prepend=np.inf)), np.abs(np.diff(ref_scores, axis=1, append=-np.inf))) < 1e-3
    print("split mode: logits %.2e boxes %.2e memory %.2e scores %.2e; frames %d decidable %d mismatching-decidable %d; un-forced rank positions "
          "differing %d of %d (all on gaps < 1e-3: %s)" % (e_log, e_box, rel(st["memory"][:, ::8, ::4], fx["memory_s"]),
                                                            rel(st["topk_scores"], fx["topk_scores"]), n, dec, bad, mism.sum(), mism.size,
                                                            bool(tie[mism].all())))
    assert bad == 0 and dec > 0.05 * n       # (random synthetic weights leave most frames at near-ties; trained weights do not)
    assert tie[mism].all() and mism.mean() < 0.05


def test_bench_shape_split_precision_vs_oracle(bench_shape):
    """the same mode at the shape bench.py times (its `parity_mode` leg): B = 64, every stage within 1e-3 of the oracle, decidable frames
    identical, CUDA-graph replay == eager launches"""
    model, x, ref, rst = bench_shape
    xg = x.cuda()
    out, st = run_engine(model, xg, force=rst["topk_idx"], split=True)
    for a, b, name in ((st["feats"][2][0], rst["feats"][2], "feat_c5"), (st["memory"], rst["memory"], "memory"),
                       (st["topk_scores"], rst["topk_scores"], "scores"), (st["hs"][5], rst["hs"][5], "hs5"),
                       (out["pred_logits"], ref["pred_logits"], "logits"), (out["pred_boxes"], ref["pred_boxes"], "boxes"),
                       (out["aux_outputs"][0]["pred_logits"], ref["aux_outputs"][0]["pred_logits"], "aux0 logits"),
                       (out["interm_outputs"]["pred_boxes"], ref["interm_outputs"]["pred_boxes"], "interm boxes")):
        e = rel(a.float(), b)
        print("B=64 split %-12s rel-to-max %.2e" % (name, e))
        assert e < TOL, name
    n, dec, bad = _decidable_frames(ref, out, 1e-3, 1e-2)
    print("B=64 split frames %d decidable (cx gap >= 1e-3, top-2 gap >= 1e-2) %d mismatching-decidable %d" % (n, dec, bad))
    assert bad == 0 and dec > 0.05 * n       # (random synthetic weights leave most frames at near-ties; trained weights do not)
    model.split_precision = True
    model.use_cuda_graph = True
    try:
        with torch.no_grad():
            o1 = model(xg)
            l1, b1 = o1["pred_logits"].clone(), o1["pred_boxes"].clone()
            o2 = model(xg)
        assert torch.equal(o2["pred_logits"], l1) and torch.equal(o2["pred_boxes"], b1)
        model.use_cuda_graph = False
        with torch.no_grad():
            oe = model(xg)
        assert torch.equal(oe["pred_logits"], l1) and torch.equal(oe["pred_boxes"], b1)
    finally:
        model.use_cuda_graph = False
        model.split_precision = False


@pytest.mark.parametrize("dt,MODE_TOL,min_agree", [(torch.float16, 1e-2, 0.97), (torch.bfloat16, 5e-2, 0.9)])
def test_bench_shape_throughput_mode_vs_oracle(bench_shape, dt, MODE_TOL, min_agree):
    """the modes bench.py times (16-bit tensor-core operands and activations, fp32 accumulation), at the bench shape, against the
    oracle with the reference ranking forced.  The bounds asserted here are the modes' documented error budgets (DESIGN.md 2.1,
    regenerated by tests/precision_sim.py): no single-pass 16-bit operand format reaches 1e-3 through ~100 dependent layers; fp16
    (11 significand bits, the default throughput mode) is ~6x closer than bf16 (8 bits) at the same speed."""
    model, x, ref, rst = bench_shape
    out, st = run_engine(model, x.cuda(), force=rst["topk_idx"], dtype=dt)
    errs = {}
    for a, b, name in ((st["feats"][2][0], rst["feats"][2], "feat_c5"), (st["memory"], rst["memory"], "memory"),
                       (st["topk_scores"], rst["topk_scores"], "scores"), (st["hs"][5], rst["hs"][5], "hs5"),
                       (out["pred_logits"], ref["pred_logits"], "logits"), (out["pred_boxes"], ref["pred_boxes"], "boxes")):
        errs[name] = rel(a.float(), b)
    mine, _ = _query_decisions(out["pred_logits"].float().cpu())
    want, margin = _query_decisions(ref["pred_logits"])
    dec = margin >= 0.1
    agree_all = (mine == want).float().mean().item()
    bad = int(((mine != want) & dec).sum())
    print("B=64 %s per-stage rel-to-max: %s ; per-query blank/label decisions equal %.4f; of the %d decided by >= 0.1 in the reference, %d differ" % (
        dt, " ".join("%s %.2e" % kv for kv in errs.items()), agree_all, int(dec.sum()), bad))
    assert errs["logits"] < MODE_TOL and errs["boxes"] < MODE_TOL
    assert bad <= 0.005 * int(dec.sum()) and agree_all >= min_agree


def test_hwdb_wide_head_engine_vs_oracle():
    """BASELINE config 3 (config/HWDB_full.py, 7356 classes) against the ORACLE: the C-wide class heads
    (N = 7356 GEMMs on 912 tokens + 6 x 900 queries, 7357-column CTC view) in fp32 at 1e-3 with identical frames, and the
    throughput mode within its documented bound."""
    from dtlr_b200 import config
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, _, _ = dino.build_dino(config.hwdb_args(num_queries=300))
    synth.load_synth_weights(model, 0)
    model = model.cuda().eval()
    x = synth.synth_images(2, 40, 1024, seed=7)
    ref, rst = _oracle_run(model, x, dict(num_queries=300, num_classes=7356))
    out, st = run_engine(model, x.cuda(), force=rst["topk_idx"])
    assert out["pred_logits"].shape == (2, 300, 7356)
    for a, b, name in ((st["topk_scores"], rst["topk_scores"], "scores"), (out["pred_logits"], ref["pred_logits"], "logits"),
                       (out["pred_boxes"], ref["pred_boxes"], "boxes"),
                       (out["interm_outputs"]["pred_logits"], ref["interm_outputs"]["pred_logits"], "interm logits")):
        e = rel(a.float(), b)
        print("HWDB fp32 %-13s rel-to-max %.2e" % (name, e))
        assert e < TOL, name
    n, dec, bad = _decidable_frames(ref, out, 2e-6, 1e-5)
    print("HWDB fp32 frames %d decidable (cx gap >= 2e-6, top-2 gap >= 1e-5) %d mismatching-decidable %d" % (n, dec, bad))
    assert bad == 0
    for dt, tol in ((torch.bfloat16, 5e-2), (torch.float16, 1e-2)):
        out16, _ = run_engine(model, x.cuda(), force=rst["topk_idx"], dtype=dt)
        e = rel(out16["pred_logits"].float(), ref["pred_logits"])
        print("HWDB %s logits rel-to-max %.2e" % (dt, e))
        assert e < tol


def test_host_pipeline_matches_direct_calls():
    """dtlr_b200.pipeline.HostPipeline (overlapped H2D / forward+decode / D2H) returns, in order, exactly what direct calls do,
    also with CUDA-graph replay (static output buffers) underneath."""
    from dtlr_b200.pipeline import HostPipeline
    model, _, _ = build_model(300)
    model.eval()
    model.compute_dtype = torch.bfloat16
    batches = [synth.synth_images(2, 40, 704, seed=40 + i).pin_memory() for i in range(5)]
    with torch.no_grad():
        model.use_cuda_graph = False
        direct = [dino.decode_frames(model(b.cuda())).cpu() for b in batches]
        for graph in (False, True):
            model.use_cuda_graph = graph
            got = list(HostPipeline(model).run(iter(batches)))       # default copy=True: results are not aliased ring buffers
            assert len(got) == 5 and all(torch.equal(a, b) for a, b in zip(got, direct))


def test_real_resolution_line_vs_oracle():
    """evaluation-time geometry (reference datasets/IAM.py:225-229 resizes lines to ~94x1333 -> S = 2676 tokens, odd feature
    widths 167/84/42/21): exercises the large-map MSDA paths (fp32 slab does not fit shared memory -> L1/L2 gather), the
    im2col fallback for non-power-of-two widths and ragged level sizes, against the CPU oracle on the same weights."""
    from oracle import dino_ref
    model, _, _ = build_model(900)
    x = synth.synth_images(1, 94, 1333, seed=11)
    sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
    st = {}
    ref = dino_ref.dino_forward(sd, dino_ref.default_cfg(num_queries=900), x, stages=st)
    assert st["src_flatten"].shape[1] == 2676
    out, est = run_engine(model, x.cuda(), force=st["topk_idx"])
    assert rel(est["memory"], st["memory"]) < TOL and rel(est["topk_scores"], st["topk_scores"]) < TOL
    assert rel(out["pred_logits"], ref["pred_logits"]) < TOL and rel(out["pred_boxes"], ref["pred_boxes"]) < TOL
    out16, _ = run_engine(model, x.cuda(), force=st["topk_idx"], dtype=torch.bfloat16)
    assert rel(out16["pred_logits"], ref["pred_logits"]) < 5e-2 and rel(out16["pred_boxes"], ref["pred_boxes"]) < 5e-2


def test_ddp_wrapped_finetune_step_single_rank_nccl():
    """the model is a plain nn.Module: DistributedDataParallel over NCCL (reference finetuning.py:211-215) wraps it and a CTC
    fine-tuning step (forward with targets -> loss_CTC -> backward -> clip -> AdamW) runs; world size 1 here (the 1-GPU test
    box), the gradient all-reduce path is the same code."""
    import os
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    model, crit, _ = build_model(300)
    model.train()
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        ddp = DDP(model, device_ids=[0], find_unused_parameters=True)
        opt = torch.optim.AdamW([p for p in ddp.parameters() if p.requires_grad], lr=1e-5)
        tg = [{k: v.cuda() for k, v in t.items()} for t in synth.synth_targets(2, 166, seed=3)]
        before = model.class_embed[0].weight.detach().clone()
        out = ddp(synth.synth_images(2, 40, 1024, seed=3).cuda(), tg)
        loss = crit.loss_CTC(out, tg, None, None)["loss_CTC"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ddp.parameters(), 0.01)            # clip_max_norm of config/Latin_CTC.py:18
        opt.step()
        assert torch.isfinite(loss) and not torch.equal(before, model.class_embed[0].weight.detach())
    finally:
        if created:
            dist.destroy_process_group()
