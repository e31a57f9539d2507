"""Multi-GPU plumbing: images are independent (decoder.ml:422-427), so a batch is split by index across
ranks (one process per GPU) with no data-path collective; torch.distributed is used only to line the ranks
up for timing and to combine per-rank results."""


def shard_range(total, rank, world):
    """Contiguous range [lo, hi) of batch indices handled by `rank` (GPU g of G takes
    [g*B/G, (g+1)*B/G))."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return total * rank // world, total * (rank + 1) // world


def barrier(dist=None, cuda_sync=None):
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if cuda_sync is not None:
        cuda_sync()


def max_over_ranks(x, dist=None, device="cpu"):
    """The slowest rank defines a multi-GPU time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch

    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj, dist=None):
    """All ranks' per-shard results in rank order (used by tests and tools, not by the timed path)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
