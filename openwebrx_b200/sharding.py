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


def gather_block(out, my_slice, group=None):
    """One hop of the SHARDED ingest: every rank holds 1/N of the wideband block (what it uploaded over its own PCIe link);
    afterwards every rank holds the whole block.  `out`: flat tensor of N x len(my_slice) elements.  NCCL over NVLink on GPUs,
    gloo in CPU tests."""
    import torch.distributed as dist
    return dist.all_gather_into_tensor(out, my_slice, group=group)


class _HopBase:
    """double-buffered per-block hop: gather(j, my_slice) on the hop's own stream, one block ahead of the consumer;
    recv(j, stream) makes the consumer wait for the landing and returns the whole block; release(j, stream) marks its last read"""

    def _init_common(self, block_samples, world, rank, device):
        import torch
        self._torch = torch
        self.world, self.rank = world, rank
        self.block_floats = 2 * int(block_samples)
        if self.block_floats % world:
            raise ValueError("the block must divide evenly among the ranks")
        self.shard_floats = self.block_floats // world
        self.stream = torch.cuda.Stream(device=device)
        self.landed = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self._timing = []

    def recv(self, j, stream):
        stream.wait_event(self.landed[j & 1])
        return self.bufs[j & 1]

    def release(self, j, stream):
        self.free[j & 1].record(stream)

    def _timed(self):
        torch = self._torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if len(self._timing) < 256:
            self._timing.append((e0, e1))
        return e0, e1

    def mean_ms(self):
        """mean device time of a gather on the hop's stream (waits for the peers included), first calls excluded"""
        self.stream.synchronize()
        t = [a.elapsed_time(b) for a, b in self._timing[3:]]
        self._timing = []
        return sum(t) / len(t) if t else None


class NcclGatherHop(_HopBase):
    """the slices are all-gathered by NCCL (copy kernels on every rank's SMs, NCCL_MAX_NCHANNELS bounds how many)"""

    def __init__(self, block_samples, world, rank, device, group=None):
        import torch
        self._init_common(block_samples, world, rank, device)
        self.group = group
        self.bufs = [torch.empty(self.block_floats // 2, 2, device=device), torch.empty(self.block_floats // 2, 2, device=device)]
        import os
        self.kind = "ncclAllGather (NCCL_MAX_NCHANNELS=%s)" % os.environ.get("NCCL_MAX_NCHANNELS", "default")

    def gather(self, j, my_slice):
        b = j & 1
        torch = self._torch
        self.stream.wait_event(self.free[b])                     # this rank's consumer is done with buffer b
        with torch.cuda.stream(self.stream):
            e0, e1 = self._timed()
            e0.record(self.stream)
            gather_block(self.bufs[b].view(-1), my_slice.reshape(-1), self.group)
            e1.record(self.stream)
            self.landed[b].record(self.stream)


class PullGatherHop(_HopBase):
    """SM-free all-gather over NVSwitch: both block buffers live in ONE symmetric allocation (torch.distributed symmetric memory:
    every rank's buffer is mapped into every other rank's address space), a rank copies its own slice in and PULLS the other
    N - 1 slices with peer-to-peer cudaMemcpyAsync — copy engines, no copy kernel beside the DSP kernels — each rank reading
    from N - 1 different peers at once, so no single NVLink port carries more than one block per hop.  Ordering across GPUs
    is by the symmetric memory's stream-ordered signals: "my slice of block j is in place" (channel b) before a peer pulls it,
    "I have pulled your slice" (channel 2 + b) before its owner overwrites it two blocks later.  Every wait carries a time-out
    (a lost peer traps the waiting stream instead of hanging the GPU)."""

    TIMEOUT_MS = 20000

    def __init__(self, block_samples, world, rank, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self._init_common(block_samples, world, rank, device)
        self.mem = symm.empty(2 * self.block_floats, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.mem, group if group is not None else dist.group.WORLD)
        self.peer = [self.handle.get_buffer(r, (2 * self.block_floats,), torch.float32, 0) for r in range(world)]
        self.bufs = [self.mem[:self.block_floats].view(-1, 2), self.mem[self.block_floats:].view(-1, 2)]
        # the N - 1 pulls go out on several streams: one peer copy keeps ONE copy engine busy (~200 GB/s over NVLink 5 measured,
        # 1.19 ms for the 235 MB a rank pulls of a 2^25-sample block at N = 8), several run on different engines side by side
        import os
        self.lanes = max(1, min(world - 1, int(os.environ.get("OWRX_HOP_LANES", "4"))))
        self.side = [torch.cuda.Stream(device=device) for _ in range(self.lanes - 1)]
        self.fork = torch.cuda.Event()
        self.joined = [torch.cuda.Event() for _ in self.side]
        self.kind = ("copy-engine pull all-gather over symmetric memory (peer cudaMemcpyAsync on %d streams, signal-ordered)" % self.lanes)
        self.calls = [0, 0]

    def gather(self, j, my_slice):
        b = j & 1
        torch, h, S = self._torch, self.handle, self.shard_floats
        base = b * self.block_floats
        self.stream.wait_event(self.free[b])                     # this rank's consumer is done with buffer b
        with torch.cuda.stream(self.stream):
            e0, e1 = self._timed()
            e0.record(self.stream)
            if self.calls[b]:
                for k in range(1, self.world):                   # every peer has pulled the slice this one replaces
                    h.wait_signal((self.rank + k) % self.world, channel=2 + b, timeout_ms=self.TIMEOUT_MS)
            self.calls[b] += 1
            self.mem[base + self.rank * S: base + (self.rank + 1) * S].copy_(my_slice.reshape(-1), non_blocking=True)
            for k in range(1, self.world):
                h.put_signal((self.rank + k) % self.world, channel=b, timeout_ms=self.TIMEOUT_MS)
            self.fork.record(self.stream)
            lanes = [self.stream] + self.side
            for k in range(1, self.world):                       # staggered: rank r starts with r + 1, so the sources differ at any moment
                p = (self.rank + k) % self.world
                st = lanes[(k - 1) % self.lanes]
                with torch.cuda.stream(st):
                    if st is not self.stream and k - 1 < self.lanes:
                        st.wait_event(self.fork)
                    h.wait_signal(p, channel=b, timeout_ms=self.TIMEOUT_MS)
                    self.mem[base + p * S: base + (p + 1) * S].copy_(self.peer[p][base + p * S: base + (p + 1) * S], non_blocking=True)
            for st, ev in zip(self.side, self.joined):
                ev.record(st)
                self.stream.wait_event(ev)
            for k in range(1, self.world):
                h.put_signal((self.rank + k) % self.world, channel=2 + b, timeout_ms=self.TIMEOUT_MS)
            e1.record(self.stream)
            self.landed[b].record(self.stream)


class PushGatherHop(PullGatherHop):
    """the same all-gather with the copies turned round: a rank WRITES its slice into the N - 1 peers' buffers (posted writes over
    NVLink: no read round trip per request).  Signals: "my buffer b may be overwritten" (channel 2 + b) to every peer before,
    "my slice has landed in your buffer" (channel b) after each copy."""

    def __init__(self, block_samples, world, rank, device, group=None):
        super().__init__(block_samples, world, rank, device, group)
        self.kind = "copy-engine push all-gather over symmetric memory (peer cudaMemcpyAsync on %d streams, signal-ordered)" % self.lanes

    def gather(self, j, my_slice):
        b = j & 1
        torch, h, S = self._torch, self.handle, self.shard_floats
        base = b * self.block_floats
        mine = my_slice.reshape(-1)
        self.stream.wait_event(self.free[b])                     # this rank's consumer is done with buffer b
        with torch.cuda.stream(self.stream):
            e0, e1 = self._timed()
            e0.record(self.stream)
            for k in range(1, self.world):
                h.put_signal((self.rank + k) % self.world, channel=2 + b, timeout_ms=self.TIMEOUT_MS)
            self.mem[base + self.rank * S: base + (self.rank + 1) * S].copy_(mine, non_blocking=True)
            self.fork.record(self.stream)
            lanes = [self.stream] + self.side
            for k in range(1, self.world):                       # staggered: the destinations differ at any moment
                p = (self.rank + k) % self.world
                st = lanes[(k - 1) % self.lanes]
                with torch.cuda.stream(st):
                    if st is not self.stream and k - 1 < self.lanes:
                        st.wait_event(self.fork)
                    h.wait_signal(p, channel=2 + b, timeout_ms=self.TIMEOUT_MS)
                    self.peer[p][base + self.rank * S: base + (self.rank + 1) * S].copy_(mine, non_blocking=True)
                    h.put_signal(p, channel=b, timeout_ms=self.TIMEOUT_MS)
            for st, ev in zip(self.side, self.joined):
                ev.record(st)
                self.stream.wait_event(ev)
            for k in range(1, self.world):
                h.wait_signal((self.rank + k) % self.world, channel=b, timeout_ms=self.TIMEOUT_MS)
            e1.record(self.stream)
            self.landed[b].record(self.stream)


class MulticastGatherHop(_HopBase):
    """the all-gather over NVSwitch multicast (NVLS): every rank streams ITS slice through the multicast address of the symmetric
    allocation (libowrx_b200's owrx_iq_multicast_store: multimem.st, a few dozen CTAs) — the slice leaves its NVLink port ONCE and
    the switch replicates it into every rank's buffer, so a rank sends 1/N of the block instead of (N-1)/N.  Two cross-GPU barriers
    on the hop stream per block: "every rank's consumer has finished with buffer b" before, "every slice has landed" after."""

    def __init__(self, block_samples, world, rank, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        self._N = N
        self._init_common(block_samples, world, rank, device)
        self.mem = symm.empty(2 * self.block_floats, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.mem, group if group is not None else dist.group.WORLD)
        self.mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if not self.mc:
            raise RuntimeError("no NVSwitch multicast support for this group")
        self.bufs = [self.mem[:self.block_floats].view(-1, 2), self.mem[self.block_floats:].view(-1, 2)]
        self.kind = "NVSwitch multicast all-gather (multimem.st of each rank's slice, two cross-GPU barriers per block)"

    def gather(self, j, my_slice):
        b = j & 1
        torch, S = self._torch, self.shard_floats
        mine = my_slice.reshape(-1)
        self.stream.wait_event(self.free[b])                     # this rank's consumer is done with buffer b
        with torch.cuda.stream(self.stream):
            e0, e1 = self._timed()
            e0.record(self.stream)
            self.handle.barrier(channel=0)                       # ... and so is every other rank's
            dst = self.mc + 4 * (b * self.block_floats + self.rank * S)
            self._N.check(self._N.lib.owrx_iq_multicast_store(mine.data_ptr(), dst, 4 * S, self.stream.cuda_stream))
            self.handle.barrier(channel=1)                       # every slice has landed in every rank's buffer b
            e1.record(self.stream)
            self.landed[b].record(self.stream)


def make_hop(kind, block_samples, world, rank, device, group=None):
    """kind: "nccl" | "pull" | "push" | "auto".  auto = nccl: measured on 8 x B200 with 2^25-sample blocks (profiles/r2_hop.md), alone
    ncclAllGather needs 0.41 ms per block, the copy-engine all-gathers 0.62-0.78 ms (pull) / 0.69-0.71 ms (push) whatever the number
    of copy streams; beside the DSP pass all three stretch to 1.0-1.2 ms and the step is 1.08 ms with NCCL, 1.19-1.25 ms with the
    copy engines — the SM-free transports do not pay for themselves here.  Every rank takes the same branch."""
    import sys
    import torch
    import torch.distributed as dist
    if kind in ("pull", "push", "multicast"):
        hop, ok = None, 1
        try:
            hop = {"push": PushGatherHop, "pull": PullGatherHop, "multicast": MulticastGatherHop}[kind](block_samples, world, rank, device, group)
        except Exception as e:                               # no symmetric memory on this box / torch build
            print("[hop] pull all-gather unavailable (%s: %s)" % (type(e).__name__, e), file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            return hop
        if kind in ("pull", "push", "multicast"):
            raise RuntimeError("OWRX_HOP=%s but symmetric memory is not available on every rank" % kind)
    return NcclGatherHop(block_samples, world, rank, device, group)
