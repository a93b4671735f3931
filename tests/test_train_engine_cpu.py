"""CPU check of the native fine-tune step's orchestration (dtlr_b200/train_engine.py): with the torch stand-in of the kernel namespace
(tests/train_ops_double.py) the explicit backward must reproduce torch autograd of the reference-shaped module path -- loss, every
parameter gradient, and the parameters after clip_grad_norm_ + AdamW (reference engine.py:192-241)."""
import copy
import warnings

import pytest
import torch

from dtlr_b200 import config, dino, ms_deform_attn, synth, train_engine
from dtlr_b200.misc import nested_tensor_from_tensor_list

import train_ops_double as KD


class _TorchMSDA(torch.autograd.Function):
    """differentiable torch core in place of the C-ABI op, so that the module path runs on the CPU"""

    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step=64):
        hw = [tuple(int(v) for v in r) for r in shapes.tolist()]
        return KD.msda_core(value, loc, attn, hw)


@pytest.fixture(scope="module")
def setup():
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, crit, _ = dino.build_dino(config.latin_ctc_args(num_queries=60, device="cpu"))
    synth.load_synth_weights(model, 0)
    model.train()
    model.use_engine = False
    B = 2
    x = synth.synth_images(B, 40, 256, seed=5)
    tg = synth.synth_targets(B, 166, seed=5, mean_len=6.0, std_len=2.0, min_len=3, max_len=9)
    return model, crit, x, tg


def _reference_step(model, crit, x, tg, lr, lr_backbone, wd, max_norm):
    ref = copy.deepcopy(model)
    old = ms_deform_attn.MSDeformAttnFunction
    ms_deform_attn.MSDeformAttnFunction = _TorchMSDA
    try:
        out = ref.forward_modules(nested_tensor_from_tensor_list(x), tg)
        crit.fused_ctc = False
        loss = crit.loss_CTC(out, tg, None, None)["loss_CTC"]
        loss.backward()
    finally:
        ms_deform_attn.MSDeformAttnFunction = old
    named = [(n, p) for n, p in ref.named_parameters() if p.requires_grad]
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in named}
    opt = torch.optim.AdamW([{"params": [p for n, p in named if "backbone" not in n], "lr": lr},
                             {"params": [p for n, p in named if "backbone" in n], "lr": lr_backbone}], lr=lr, weight_decay=wd)
    torch.nn.utils.clip_grad_norm_([p for _, p in named], max_norm)
    opt.step()
    return float(loss), grads, {n: p.detach().clone() for n, p in ref.named_parameters()}


def test_native_step_matches_autograd_on_cpu(setup):
    model, crit, x, tg = setup
    lr, lrb, wd, mn = 1e-3, 1e-4, 1e-2, 0.1
    loss_ref, g_ref, p_ref = _reference_step(model, crit, x, tg, lr, lrb, wd, mn)
    m2 = copy.deepcopy(model)
    eng = train_engine.TrainEngine(m2, lr=lr, lr_backbone=lrb, weight_decay=wd, max_norm=mn, dtype=torch.float32, K=KD)
    eng.zero_grad()
    loss = eng.forward_backward(x, tg)
    assert abs(float(loss) - loss_ref) < 1e-4 * abs(loss_ref)
    live = {n for n, *_ in eng.layout}
    worst = 0.0
    for n, p in m2.named_parameters():
        if not p.requires_grad:
            continue
        gr = g_ref[n]
        if n not in live:
            assert gr is None or float(gr.abs().max()) == 0.0, "parameter %s left out of the arena has a reference gradient" % n
            continue
        assert gr is not None, n
        g = eng.grad(p)
        scale = float(gr.abs().max())
        err = float((g - gr).abs().max())
        # fp32 round-off of two summation orders, plus the odd ReLU unit whose pre-activation sits at zero and flips (one column of a
        # weight gradient moves by O(1e-3) of the maximum): bound the maximum loosely, the direction tightly
        assert err <= 5e-3 * scale + 1e-7, (n, err, scale)
        cos = float(torch.nn.functional.cosine_similarity(g.flatten().double(), gr.flatten().double(), dim=0))
        assert cos > 1 - 1e-5 or scale < 1e-6, (n, cos)
        worst = max(worst, err / (scale + 1e-12))
    # optimizer: clip_grad_norm_ + AdamW with the backbone learning-rate group.  Adam's first step is lr * g / (|g| + eps): elements
    # whose gradient is round-off noise around zero move by +-lr either way, so the update is compared on IDENTICAL gradients
    for n, p in m2.named_parameters():
        if n in live:
            eng.grad(p).copy_(g_ref[n])
    eng.optimizer_step()
    for n, p in m2.named_parameters():
        assert float((p.detach() - p_ref[n]).abs().max()) <= 2e-6 * (1 + float(p_ref[n].abs().max())), n


def test_arena_layout_and_dead_parameters(setup):
    model, _, _, _ = setup
    m2 = copy.deepcopy(model)
    eng = train_engine.TrainEngine(m2, dtype=torch.float32, K=KD)
    names = [n for n, *_ in eng.layout]
    assert not any(n.startswith(train_engine.DEAD_PREFIXES) for n in names)
    assert any(n.startswith("transformer.decoder.class_embed.") for n in names)
    segs = [s for _, _, _, s, _, _ in eng.layout]
    gis = [g for _, _, g, _, _, _ in eng.layout]
    assert gis == sorted(gis) and all(segs[i] <= segs[i + 1] or gis[i] < gis[i + 1] for i in range(len(segs) - 1))
    for n, p, gi, seg, o, k in eng.layout:          # parameters and gradients are views of the arenas
        assert p.data_ptr() == eng.flat_p.data_ptr() + 4 * o and p.grad.data_ptr() == eng.flat_g.data_ptr() + 4 * o
    # frozen stem / layer1 are not in the arena
    assert not any("layer1" in n or n.endswith("body.conv1.weight") for n in names)


def test_head_only_finetuning_step(setup):
    """step-1 fine-tuning (reference finetuning.py:531-539): only the class heads require grad"""
    model, crit, x, tg = setup
    m2 = copy.deepcopy(model)
    for n, p in m2.named_parameters():
        p.requires_grad_("class_embed" in n and "enc_out" not in n)
    eng = train_engine.TrainEngine(m2, dtype=torch.float32, K=KD)
    assert all(n.startswith(train_engine.HEAD_PREFIXES) for n, *_ in eng.layout)
    before = m2.class_embed[0].weight.detach().clone()
    loss = eng.step(x, tg)
    assert torch.isfinite(loss) and float((m2.class_embed[0].weight.detach() - before).abs().max()) > 0


def _ddp_worker(rank, world, port, q):
    import os
    import sys
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    import warnings
    import torch.distributed as dist
    import train_ops_double as KD_
    from dtlr_b200 import config as cfg_, dino as dino_, synth as synth_, train_engine as te_
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, _, _ = dino_.build_dino(cfg_.latin_ctc_args(num_queries=60, device="cpu"))
    synth_.load_synth_weights(model, 0)
    model.train()
    res = {}
    for w_, tag in ((1, "local"), (world, "reduced")):
        m2 = copy.deepcopy(model)
        eng = te_.TrainEngine(m2, dtype=torch.float32, K=KD_, world_size=w_)
        x = synth_.synth_images(1, 40, 256, seed=50 + rank)
        tg = synth_.synth_targets(1, 166, seed=50 + rank, mean_len=5.0, std_len=1.0, min_len=3, max_len=7)
        eng.zero_grad()
        eng.forward_backward(x, tg)
        if w_ > 1:
            for wk in eng._works:
                wk.wait()
            assert len(eng._reduced) == len(eng.chunks) == 4          # decoder | encoder | input_proj (lr group 0), backbone (group 1)
        res[tag] = eng.flat_g.clone()
    gathered = [torch.zeros_like(res["local"]) for _ in range(world)]
    dist.all_gather(gathered, res["local"])
    mean = sum(gathered) / world
    err = float((res["reduced"] - mean).abs().max() / mean.abs().max())
    q.put((rank, err, float(gathered[0].abs().sum()), float(gathered[1].abs().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_all_reduce_gloo():
    """N > 1 path of the native step (world size 2, gloo, CPU stand-in kernels): each rank's arena after the segmented asynchronous
    all-reduces = the mean over ranks of the single-rank gradients (DistributedDataParallel semantics, finetuning.py:211-215)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, 29547, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank, err, s0, s1 in res:
        assert err < 1e-5, (rank, err)
        assert s0 != s1                                               # the two ranks really had different lines


def test_finetune_loop_scheduler_and_checkpoint_resume(setup):
    """dtlr_b200/finetune.train_one_epoch_CTC (mirror of reference engine.py:172-274) + TrainEngine.state_dict round trip: two steps in
    one run == one step, checkpoint, resume in a fresh engine, one more step"""
    from types import SimpleNamespace
    from dtlr_b200 import finetune
    model, crit, x, tg = setup
    batches = [(x, tg), (x.flip(0), list(reversed(tg)))]

    def fresh():
        m = copy.deepcopy(model)
        return m, train_engine.TrainEngine(m, lr=1e-3, lr_backbone=1e-4, weight_decay=1e-2, max_norm=0.1, dtype=torch.float32, K=KD)

    m_a, e_a = fresh()
    st = finetune.train_one_epoch_CTC(e_a, batches, "cpu", 0, args=SimpleNamespace(max_iterations=100, onecyclelr=False))
    assert st["steps"] == 2 and st["loss"] > 0
    m_b, e_b = fresh()
    finetune.train_one_epoch_CTC(e_b, batches[:1], "cpu", 0)
    ck_model, ck_opt = copy.deepcopy(m_b.state_dict()), e_b.state_dict()
    m_c = copy.deepcopy(model)
    m_c.load_state_dict(ck_model)
    e_c = train_engine.TrainEngine(m_c, lr=1e-3, lr_backbone=1e-4, weight_decay=1e-2, max_norm=0.1, dtype=torch.float32, K=KD)
    e_c.load_state_dict(ck_opt)
    finetune.train_one_epoch_CTC(e_c, batches[1:], "cpu", 1)
    pa = dict(m_a.named_parameters())
    for n, p in m_c.named_parameters():
        assert float((p.detach() - pa[n].detach()).abs().max()) <= 1e-6 * (1 + float(pa[n].abs().max())), n
    # the iteration budget counts lines (reference engine.py:259-261) and a scheduler only has to write param_groups[i]["lr"]
    m_d, e_d = fresh()
    st = finetune.train_one_epoch_CTC(e_d, batches, "cpu", 0, args=SimpleNamespace(max_iterations=2, onecyclelr=True),
                                      lr_scheduler=lambda: e_d.param_groups[0].__setitem__("lr", e_d.param_groups[0]["lr"] * 0.5))
    assert st["steps"] == 1 and abs(e_d.param_groups[0]["lr"] - 5e-4) < 1e-12
