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


def gather_row_blocks(local_rows, counts, group=None, device=None):
    """The final gather of a sharded job: every rank contributes a (counts[rank] x k) block of rows (its thetas, its
    block of predictions); returns the list of all ranks' blocks on every rank.  The only cross-rank traffic of the
    path; NCCL on the GPUs (pass device), gloo in the CPU tests."""
    import torch
    import torch.distributed as dist
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64).reshape(len(local_rows), -1)
    if not (dist.is_available() and dist.is_initialized()):
        return [local_rows]
    world = dist.get_world_size(group)
    pad = np.zeros((max(max(counts), 1), local_rows.shape[1]))
    pad[:local_rows.shape[0]] = local_rows
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [o.cpu().numpy()[:c] for o, c in zip(outs, counts)]


def scatter_round_robin(blocks, n_items):
    """Undo round_robin: blocks[r] holds the rows of the items rank r owns; returns them in item order."""
    world = len(blocks)
    k = blocks[0].shape[1]
    out = np.zeros((n_items, k))
    for r in range(world):
        out[round_robin(n_items, world, r)] = blocks[r]
    return out


def max_over_ranks(x, group=None, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
