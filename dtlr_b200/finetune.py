"""The fine-tuning loop around TrainEngine: mirror of reference engine.py:172-274 `train_one_epoch_CTC` (what finetuning.py:623 calls once
per epoch).  The reference passes (model, criterion, optimizer) and does forward / loss_CTC / backward / clip / step itself; here the
engine owns those five steps (dtlr_b200/train_engine.py), so the loop keeps only what is loop: device transfer, the finite-loss check,
the iteration budget (`args.max_iterations` counts LINES, engine.py:259-261), the per-step scheduler and the running statistics.
"""
import math
import sys

import torch


def train_one_epoch_CTC(engine, data_loader, device, epoch, lr_scheduler=None, args=None, logger=None, run=None, print_freq=10):
    """-> {"loss": mean loss of the epoch, "lr": last learning rate of group 0, "steps": optimizer steps taken}

    engine: dtlr_b200.train_engine.TrainEngine (its max_norm is the reference's `max_norm` argument, config/Latin_CTC.py:18);
    data_loader yields (NestedTensor | list of images, list of target dicts) as the reference's collate_fn does (util/misc.py:285-289);
    lr_scheduler: a callable invoked after every step when args.onecyclelr (the reference steps OneCycleLR per batch, engine.py:243-244);
    run: an object with .log(dict) (the reference's wandb run) or None."""
    engine.model.train()
    max_iterations = getattr(args, "max_iterations", None) if args is not None else None
    onecycle = bool(getattr(args, "onecyclelr", False)) if args is not None else False
    iterations, steps, loss_sum = 0, 0, 0.0
    for it, (samples, targets) in enumerate(data_loader):
        samples = samples.to(device) if hasattr(samples, "to") else [s.to(device) for s in samples]
        targets = [{k: v.to(device) for k, v in t.items()} for t in targets]
        loss = engine.step(samples, targets)
        loss_value = float(loss)                                    # (the reference's loss.item(): one sync per step)
        if not math.isfinite(loss_value):
            print("Loss is {}, stopping training".format(loss_value))
            sys.exit(1)
        if run is not None:
            run.log({"train_loss_CTC": loss_value, "global_step": iterations})
        if onecycle and lr_scheduler is not None:
            lr_scheduler()
        loss_sum += loss_value
        steps += 1
        if logger is not None and it % print_freq == 0:
            logger.info("Epoch: [%d] step %d loss %.4f lr %.2e" % (epoch, it, loss_value, engine.param_groups[0]["lr"]))
        iterations += len(targets)
        if max_iterations is not None and iterations >= max_iterations:
            break
    stats = torch.tensor([loss_sum, float(steps)], dtype=torch.float64)
    if engine.world > 1:                                            # MetricLogger.synchronize_between_processes (engine.py:265)
        import torch.distributed as dist
        stats = stats.to(engine.device)
        dist.all_reduce(stats, group=engine.pg)
        stats = stats.cpu()
    return {"loss": float(stats[0] / max(float(stats[1]), 1.0)), "lr": engine.param_groups[0]["lr"], "steps": steps}
