"""Multi-GPU sharding of the path: shapes are independent, so each rank owns a contiguous block of
the batch and the only exchange is one all-reduce of [sum of per-shape losses, number of shapes with
a fitted ellipsoid] (8 bytes) -- the mean of src/utils.py:425 over the global batch.  Input
gradients stay with their shapes.  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the
CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_shapes, rank, world):
    """Contiguous block [lo, hi) of the global batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_shapes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_masked_mean(loss_b, has, group=None):
    """loss_b[B_local], has[B_local] in {0,1} -> (L_global (value), L_for_backward).

    L_for_backward = sum_local(loss_b * has) / n_valid_global: calling .backward() on it on every rank
    yields exactly the gradient of the global mean for the local shapes."""
    local = torch.stack([(loss_b * has).sum(), has.sum()])
    tot = local.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    n = tot[1].clamp(min=1.0)
    return tot[0] / n, local[0] / n


def global_mean_from_local(local_mean, n_local, group=None):
    """Same reduction when a rank only has its local masked mean and its count of valid shapes."""
    local = torch.stack([local_mean.reshape(()) * n_local, n_local])
    tot = local.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    n = tot[1].clamp(min=1.0)
    return tot[0] / n, local[0] / n


def global_loss(out, group=None):
    """Same reduction from pipeline.fit_loss()'s fused outputs (`loss_sum`, `n_valid`, `loss`): on one rank the
    local mean is the global mean and nothing is enqueued.  fit_loss(dist_reduce=True) has already done it."""
    if "loss_global" in out:
        return out["loss_global"], out["loss_backward"]
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return out["loss"].detach(), out["loss"]
    tot = torch.stack([out["loss_sum"].detach(), out["n_valid"]])
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    n = tot[1].clamp(min=1.0)
    return tot[0] / n, out["loss_sum"] / n
