"""Step-glue helpers for the reference's training loop (tool/train.py:278-304).

The reference's loop ends every iteration with SEVEN blocking all-reduces (main_loss, aux_loss, loss, count, then
intersection, union, target) followed by three `.item()` / `.cpu()` synchronisations. `reduce_step_logging` packs the
same quantities into ONE fp32 buffer, issues ONE all-reduce and returns device tensors; nothing synchronises the host
until the caller actually formats a log line (every `print_freq` iterations). The arithmetic is the reference's:

    main_loss, aux_loss, loss  <-  sum_r(x_r * n_r) / sum_r(n_r)          (tool/train.py:280-284)
    intersection, union, target <- sum over ranks                          (tool/train.py:286-288)

`intersection / union / target` come from semseg_b200.metrics.intersectionAndUnionGPU (one-pass kernel, exact counts).
Drop-in use inside tool/train.py's loop (the three blocks :278-284, :286-289 and the .item() calls collapse into):

    m = reduce_step_logging(main_loss, aux_loss, loss, input.size(0), intersection, union, target)
    intersection_meter.update(m.intersection), ...      # device tensors; AverageMeter works on tensors unchanged
"""
import collections

import torch
import torch.distributed as dist

StepLog = collections.namedtuple("StepLog", "main_loss aux_loss loss n intersection union target")


def reduce_step_logging(main_loss, aux_loss, loss, n, intersection, union, target, group=None):
    """One all-reduce for the seven per-iteration logging reductions. All returned values are device tensors (no host
    sync); `n` is the global image count as a 0-d tensor."""
    k = intersection.numel()
    dev = intersection.device
    buf = torch.empty((4 + 3 * k,), dtype=torch.float32, device=dev)
    nf = float(n)
    buf[0] = main_loss.detach().float() * nf
    buf[1] = aux_loss.detach().float() * nf
    buf[2] = loss.detach().float() * nf
    buf[3] = nf
    buf[4:4 + k] = intersection.float()
    buf[4 + k:4 + 2 * k] = union.float()
    buf[4 + 2 * k:] = target.float()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, group=group)
    tot = buf[3]
    return StepLog(buf[0] / tot, buf[1] / tot, buf[2] / tot, tot, buf[4:4 + k], buf[4 + k:4 + 2 * k], buf[4 + 2 * k:])
