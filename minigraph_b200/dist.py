"""Multi-GPU plumbing of the mapping path: reads are sharded, the index is replicated, and the
only collective is an all-gather of per-rank GAF byte counts that turns into output offsets.

One process per GPU (torch.distributed: NCCL on the GPU box, gloo in the CPU tests)."""


def shard_bounds(lengths, world):
    """Contiguous blocks of reads balanced by cumulative bases: returns `world`+1 boundaries (reference order is kept,
    so concatenating the shards' GAF in rank order reproduces the single-process output, gmap.c:107-135)."""
    tot = sum(lengths)
    bounds, acc, r = [0], 0, 1
    for i, l in enumerate(lengths):
        while r < world and acc >= tot * r / world:
            bounds.append(i)
            r += 1
        acc += l
    while len(bounds) < world:
        bounds.append(len(lengths))
    bounds.append(len(lengths))
    return bounds


def gaf_offsets(n_bytes, device=None):
    """All-gather this rank's GAF byte count; returns (offset of this rank, list of all counts)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, [n_bytes]
    mine = torch.tensor([n_bytes], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allc, mine)
    counts = [int(c.item()) for c in allc]
    return sum(counts[:dist.get_rank()]), counts
