"""Multi-GPU partitioning of the hot path's independent work units: PCA components, optimizer
restarts and query-point blocks.  One process per GPU; no collective in the data path -- only the final
gather of thetas / likelihoods / predictions crosses ranks (torch.distributed: NCCL on the GPUs, gloo in the
CPU tests)."""
import numpy as np


def block_range(n_items, world, rank):
    """Contiguous block [lo, hi) of n_items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def round_robin(n_items, world, rank):
    """Item indices owned by `rank` under static round-robin (component c -> GPU c mod world)."""
    return list(range(rank, n_items, world))


def shard_map_rows(fn, rows, group=None, device=None):
    """Apply `fn(block) -> array (len(block) x k)` to this rank's contiguous block of `rows` and gather the
    blocks of all ranks in order.  Every rank returns the full (len(rows) x k) result.  With no initialised
    process group this is just fn(rows)."""
    import torch
    import torch.distributed as dist
    rows = np.asarray(rows)
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(fn(rows), dtype=np.float64)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = block_range(len(rows), world, rank)
    local = np.asarray(fn(rows[lo:hi]), dtype=np.float64)
    local = local.reshape(hi - lo, -1)
    k = local.shape[1]
    sizes = [block_range(len(rows), world, r) for r in range(world)]
    maxlen = max(h - l for l, h in sizes)
    pad = np.zeros((maxlen, k))
    pad[:hi - lo] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.concatenate([o.cpu().numpy()[:h - l] for o, (l, h) in zip(outs, sizes)], axis=0)


def max_over_ranks(x, group=None, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
