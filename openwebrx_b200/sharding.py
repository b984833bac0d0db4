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


class MulticastHop:
    """The per-block IQ hop over NVSwitch multicast (NVLS), double-buffered.

    Every rank holds two block buffers inside ONE symmetric allocation (torch.distributed symmetric memory: the same
    offset on every GPU, plus a multicast address that maps them all).  `send(j, src)` is called by every rank for block j:
    the ingest rank streams `src` through the multicast address with libowrx_b200's owrx_iq_multicast_store — the block
    leaves its NVLink port once and the switch replicates it into every rank's buffer j & 1 (no copy kernel, no SM time on
    the receivers) — bracketed by two cross-GPU barriers on the hop stream: "every rank has finished reading this buffer"
    before, "the block has landed everywhere" after.  `recv(j, stream)` makes `stream` wait for the landing and returns
    the local buffer; `release(j, stream)` marks the consumer's last read.  torch is only the transport set-up here.
    Raises RuntimeError when the GPUs have no multicast support (the caller falls back to broadcast_block)."""

    def __init__(self, block_floats, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        self._N, self._torch = N, torch
        self.block_bytes = int(block_floats) * 4
        if self.block_bytes % 16:
            raise ValueError("block must be a multiple of 16 bytes")
        self.mem = symm.empty(2 * int(block_floats), dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.mem, group if group is not None else dist.group.WORLD)
        self.mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if not self.mc:
            raise RuntimeError("no NVSwitch multicast support for this group")
        self.bufs = [self.mem[:block_floats], self.mem[block_floats:]]
        self.stream = torch.cuda.Stream(device=device)
        self.landed = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]

    def send(self, j, src=None):
        """all ranks; `src` (a device tensor of one block) on the ingest rank only"""
        b = j & 1
        torch = self._torch
        self.stream.wait_event(self.free[b])                     # this rank's consumer is done with buffer b
        with torch.cuda.stream(self.stream):
            self.handle.barrier(channel=0)                       # ... and so is every other rank's
            if src is not None:
                self._N.check(self._N.lib.owrx_iq_multicast_store(src.data_ptr(), self.mc + b * self.block_bytes, self.block_bytes,
                                                                  self.stream.cuda_stream))
            self.handle.barrier(channel=1)                       # the block has landed in every rank's buffer b
            self.landed[b].record(self.stream)

    def recv(self, j, stream):
        stream.wait_event(self.landed[j & 1])
        return self.bufs[j & 1]

    def release(self, j, stream):
        self.free[j & 1].record(stream)
