"""Multi-GPU plumbing: one process per GPU, independent graphs sharded by rank, no data-path collective.

OfflineAudioContexts share nothing, so a batch of G graphs is split into contiguous rank shards; every
rank renders its shard with its own engine.  The only collective is the optional final gather of rendered PCM
(north_star) — torch.distributed `gather` over NCCL (NVLink/NVSwitch) on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_graphs, rank, world):
    """Contiguous, balanced shard [g0, g1) of rank `rank` (sizes differ by at most one)."""
    g0 = n_graphs * rank // world
    g1 = n_graphs * (rank + 1) // world
    return g0, g1


def gather_pcm(local, n_graphs_total, dst=0):
    """Gathers the ranks' rendered PCM shards ([g_local, ch, length] tensors, possibly different g_local) on `dst`.

    Returns the [n_graphs_total, ch, length] tensor on dst, None elsewhere.  Uneven shards are padded to the largest
    shard for the collective and trimmed afterwards."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return local
    sizes = [shard_range(n_graphs_total, r, world) for r in range(world)]
    max_n = max(g1 - g0 for g0, g1 in sizes)
    pad = local
    if local.shape[0] < max_n:
        pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def all_gather_group(full_k, shard_slice):
    """One graph group of the pipelined gather bench.py runs (north_star "NCCL gather of rendered PCM"): every rank contributes the
    PCM of the same local graph range, `full_k` ([world, graphs_in_group, ch, length]) receives all of them.  Called per graph group on
    a side stream while the next group renders, so that the NVLink transfer hides behind the render instead of following it.
    NCCL: one all_gather_into_tensor; gloo (CPU tests): the list form."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        full_k[0].copy_(shard_slice)
        return
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(full_k.view(-1), shard_slice.reshape(-1))
    else:
        dist.all_gather([full_k[r] for r in range(full_k.shape[0])], shard_slice.contiguous())


def group_ranges(n_graphs, n_groups):
    """Contiguous graph groups [(first, last)] the way wae_batch_prepare cuts a batch without suspend points (ceil-sized pieces)."""
    n_groups = max(1, min(n_groups, n_graphs))
    target = (n_graphs + n_groups - 1) // n_groups
    return [(g0, min(n_graphs, g0 + target)) for g0 in range(0, n_graphs, target)]


def max_over_ranks(value, device="cpu"):
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
