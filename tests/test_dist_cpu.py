"""CPU: the N>1 plumbing with world_size 2 over gloo (127.0.0.1), and the rank behaviour of bench.py --impl reference."""
import json
import os
import subprocess
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from dtlr_b200 import dist_util, synth
    r, w = dist_util.init("gloo")
    dist_util.barrier()
    slow = dist_util.max_over_ranks(10.0 + 5.0 * rank)              # rank 1 is the slow one
    lo, hi = dist_util.shard_range(13, r, w)
    x = synth.synth_images(2, 8, 16, seed=100 + r)                    # each rank draws its own shard of synthetic lines
    q.put((r, w, slow, lo, hi, float(x.sum())))
    dist_util.barrier()
    dist_util.shutdown()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [2, 2]
    assert all(abs(r[2] - 15.0) < 1e-9 for r in res)                  # max over ranks, seen by both
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 7, 7, 13)
    assert res[0][5] != res[1][5]                                     # different shards
    from dtlr_b200 import dist_util
    assert dist_util.whole_job_throughput(64, 2, 10, 1000.0) == 1280.0


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env["RANK"] = "0"; env["LOCAL_RANK"] = "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["n_gpus"] == 2


def _eval_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import numpy as np
    from dtlr_b200 import dist_util, evaluation
    dist_util.init("gloo")
    rng = np.random.default_rng(0)                                     # every rank sees the SAME data set
    heights, widths = rng.integers(48, 200, 101).tolist(), rng.integers(900, 1334, 101).tolist()
    batches = evaluation.bucket_batches(widths, 8, 64, heights=heights, height_multiple=16)
    costs = [len(b) * max(heights[i] for i in b) * max(widths[i] for i in b) for b in batches]
    mine = evaluation.shard_batches(batches, costs, rank, world)
    preds = [None] * 101
    for b in mine:
        for i in b:
            preds[i] = [i % 7, heights[i] % 5]                         # stand-in for the GPU forward + decode of line i
    full = evaluation.gather_predictions(preds)
    q.put((rank, sum(len(b) for b in mine), sum(costs[batches.index(b)] for b in mine), full))
    dist_util.barrier()
    dist_util.shutdown()


def test_sharded_evaluation_two_ranks_gloo():
    """multi-GPU evaluation shards whole batches over the ranks (no data-path collective) and gathers the host-side label lists:
    every line evaluated exactly once, both ranks end with the complete, identical list, loads balanced."""
    from dtlr_b200 import evaluation
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, 29541, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == 101 and res[0][1] > 0 and res[1][1] > 0
    assert res[0][3] == res[1][3] and all(p is not None for p in res[0][3])
    assert abs(res[0][2] - res[1][2]) <= 0.15 * (res[0][2] + res[1][2])          # greedy balance by padded pixels
    # single-process properties of the assignment
    batches = [[0, 1], [2], [3, 4, 5], [6]]
    parts = [evaluation.shard_batches(batches, [5, 1, 9, 2], r, 3) for r in range(3)]
    assert sorted(b for p in parts for b in map(tuple, p)) == sorted(map(tuple, batches))
    assert evaluation.shard_batches(batches, [1, 1, 1, 1], 0, 1) == batches
    import pytest
    with pytest.raises(ValueError):
        evaluation.shard_batches(batches, [1], 0, 2)


def test_bench_secondary_views_arithmetic():
    """the helper arithmetic behind roofline_gemm.hbm_view and roofline_msda.binding_* (bench.py), on known numbers"""
    sys.path.insert(0, ROOT)
    import bench
    v = bench.gemm_hbm_view(3.0e9, 1.5, 3, 6000.0)                 # 3 GB over 1.5 ms = 2000 GB/s
    assert v["achieved"] == 2000.0 and v["frac"] == round(2000.0 / 6000.0, 4) and v["algorithmic_bytes_per_step"] == 1.0e9
    assert bench.gemm_hbm_view(0.0, 0.0, 0, 6000.0)["achieved"] == 0.0
    b = bench.msda_binding_view(107.0, 148, 1965.0, 20217621, "profiles/x.txt")
    assert abs(b["binding_floor_us"] - 69.5) < 0.1 and abs(b["binding_frac"] - 0.65) < 0.01
    # the ncu-derived constants are parsed from the committed summaries, not pasted into bench.py
    t, src = bench.profile_traffic(r"ffn_ln_(sk|tcgen05)_kernel", "ffn")
    assert src is not None and 1e6 < t < 1e9
    m, src = bench.profile_metrics(r"msda_fwd", ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",), "msda")
    assert src is not None and m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"] > 1e6
    assert bench.profile_metrics(r"no_such_kernel", ("dram__bytes_read.sum",)) == ({}, None)
