"""Channel sharding across the GPUs of one node.

Every (channel, band) cascade is independent (sosfilt.c:59-60), so a job shards by giving each rank
a contiguous block of channels, its slice of the states and a replica of the (tiny) coefficient
bank; there is no collective on the filtering path.  Because the output layout is [C][B][N]
(channel slowest, == the reference's (N, B, C) Fortran array), each rank's result is one contiguous
slab of the full tensor and a plain all-gather reassembles it -- that gather is optional and timed
separately from the filtering throughput.
"""
import torch
import torch.distributed as dist


def shard_channels(num_channels, world_size, rank):
    """Contiguous channel range [c0, c1) of `rank`; the first (num_channels % world_size) ranks
    carry one extra channel."""
    base, extra = divmod(int(num_channels), int(world_size))
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


def gather_bands(y_local, num_channels=None, group=None):
    """All-gather the per-rank (C_local, B, n) slabs into the full (C, B, n) tensor on every rank.
    Ranks may hold different channel counts (ragged shards are padded for the collective)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if num_channels is None:
        counts = torch.tensor([y_local.shape[0]], device=y_local.device)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        sizes = [int(c.item()) for c in all_counts]
    else:
        sizes = [shard_channels(num_channels, world, r)[1] - shard_channels(num_channels, world, r)[0]
                 for r in range(world)]
    cmax = max(sizes)
    slab = y_local
    if slab.shape[0] < cmax:
        pad = torch.zeros((cmax - slab.shape[0],) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
        slab = torch.cat([slab, pad], dim=0)
    out = torch.empty((world * cmax,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
    dist.all_gather_into_tensor(out, slab.contiguous(), group=group)
    if all(s == cmax for s in sizes):
        return out
    parts = [out[r * cmax:r * cmax + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
