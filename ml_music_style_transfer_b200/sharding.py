"""Multi-GPU partitioning of the hot path: clips / pieces are independent (preprocess.py:63-75,172-196;
inference.py:89-91), so rank r of W owns a contiguous balanced block and there is no data-path collective.
The only exchange is the optional final gather of features over NCCL (NVLink / NVSwitch)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous balanced partition: first (n % W) ranks get one extra item.  -> (start, stop)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world_size):
    return [shard_range(n_items, r, world_size)[1] - shard_range(n_items, r, world_size)[0] for r in range(world_size)]


def gather_features(local, n_items_total, group=None):
    """All-gather per-rank feature blocks (first dim = items of this rank) into the full (n_items_total, ...) tensor.

    Equal shards use one ``all_gather_into_tensor``; ragged shards are padded to the largest shard first.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_items_total, world)
    mx = max(sizes)
    pad = local
    if local.shape[0] != mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)
