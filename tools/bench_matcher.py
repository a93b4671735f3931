"""Times the Hungarian matching of one training step (7 matchings: 6 decoder layers + interm) at the Latin CTC shape:
dtlr_match_cost + dtlr_lsap (2 launches) against the reference procedure (full cost matrix on the GPU with torch ops, copy to
the CPU, scipy per image) on the same inputs."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200.matcher import HungarianMatcher  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--queries", type=int, default=900)
    ap.add_argument("--classes", type=int, default=97)
    ap.add_argument("--layers", type=int, default=7)
    ap.add_argument("--tmin", type=int, default=30)
    ap.add_argument("--tmax", type=int, default=90)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    sizes = torch.randint(a.tmin, a.tmax + 1, (a.batch,), generator=g).tolist()
    targets = [{"labels": torch.randint(0, a.classes, (n,), generator=g).to(dev),
                "boxes": torch.cat([torch.rand(n, 2, generator=g), torch.rand(n, 2, generator=g) * 0.1 + 0.01], -1).to(dev)} for n in sizes]
    layers = [{"pred_logits": (torch.randn(a.batch, a.queries, a.classes, generator=g) * 2 - 3).to(dev),
               "pred_boxes": torch.cat([torch.rand(a.batch, a.queries, 2, generator=g),
                                        torch.rand(a.batch, a.queries, 2, generator=g) * 0.1 + 0.01], -1).to(dev)} for _ in range(a.layers)]
    m = HungarianMatcher(2.0, 5.0, 2.0, 0.25)
    for _ in range(3):
        ours = m.match_layers(layers, targets)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    m.cpu_indices = False
    e0.record()
    cost, t_cnt, _, Tmax = m.block_costs(layers, targets)
    e1.record()
    from dtlr_b200.matcher import _lsap_target_major
    _lsap_target_major(cost, t_cnt, a.batch, Tmax)
    e2.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.cpu_indices = True
    ours = m.match_layers(layers, targets)
    torch.cuda.synchronize()
    t_ours = time.perf_counter() - t0
    from scipy.optimize import linear_sum_assignment
    t0 = time.perf_counter()
    ref = []
    for o in layers:
        C = m.cost_matrix(o, targets).cpu()
        ref.append([linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))])
    t_ref = time.perf_counter() - t0
    same = all(ours[l][b][0].tolist() == ref[l][b][0].tolist() and ours[l][b][1].tolist() == ref[l][b][1].tolist()
               for l in range(a.layers) for b in range(a.batch))
    print("B=%d Q=%d layers=%d targets/img %d..%d: cost kernel %.3f ms, lsap kernel %.3f ms, matcher wall %.2f ms | "
          "reference procedure (torch cost + scipy) %.1f ms | identical matchings: %s"
          % (a.batch, a.queries, a.layers, a.tmin, a.tmax, e0.elapsed_time(e1), e1.elapsed_time(e2), t_ours * 1e3, t_ref * 1e3, same))


if __name__ == "__main__":
    main()
