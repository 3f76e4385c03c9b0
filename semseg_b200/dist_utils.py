"""Host-side helpers for the data-parallel path (one process per GPU, torch.distributed).

The only collectives on the training hot path are (a) the per-BatchNorm statistics exchange of SyncBatchNorm
(tool/train.py:141-142) and (b) DistributedDataParallel's gradient buckets (tool/train.py:157), which torch owns.
"""
import torch
import torch.distributed as dist


def gather_rank_stats(stats, pg=None):
    """stats [3][C] = local (mean, M2, count) -> [world][3][C], one all_gather (equal shapes on every rank)."""
    world = dist.get_world_size(pg)
    rows = stats.shape[0]
    out = torch.empty((world * rows,) + tuple(stats.shape[1:]), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(out, stats.contiguous(), group=pg)  # concatenated along dim 0 (gloo and nccl)
    return out.view((world, rows) + tuple(stats.shape[1:]))


def max_over_ranks(value, device):
    """Max of a python float over all ranks (bench timing: the slowest rank defines the step time)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_batch(global_batch, world, rank):
    """Images of a global batch owned by `rank` (DistributedSampler semantics with drop_last, tool/train.py:154,204)."""
    per = global_batch // world
    return range(rank * per, (rank + 1) * per)
