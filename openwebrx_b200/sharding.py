"""Multi-GPU plumbing of the hot path: client channels are independent given the
wideband IQ block, so they are partitioned across ranks and the block is broadcast from the ingest rank
each hop.  No other collective exists on this path.  torch.distributed is only the transport."""


def shard_channels(n_channels, world_size, rank):
    """contiguous, balanced partition of channel ids: returns range(lo, hi) for this rank"""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_channels, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return range(lo, hi)


def owner_of(channel, n_channels, world_size):
    """rank that owns `channel` under shard_channels"""
    base, extra = divmod(n_channels, world_size)
    cut = extra * (base + 1)
    if channel < cut:
        return channel // (base + 1)
    return extra + (channel - cut) // max(base, 1)


def broadcast_block(block, src=0, group=None, async_op=False):
    """One hop: the ingest rank's IQ block (a torch tensor, same shape on every rank) reaches all ranks.
    NCCL over NVLink/NVSwitch on GPUs, gloo in CPU tests."""
    import torch.distributed as dist
    return dist.broadcast(block, src, group=group, async_op=async_op)
