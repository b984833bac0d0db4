"""CPU test of the N > 1 path (world_size 2, gloo): channel sharding covers every channel exactly once and the
per-hop IQ broadcast delivers the ingest rank's block bit-identically to the other rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_channels, q):
    sys.path.insert(0, ROOT)
    from openwebrx_b200.sharding import broadcast_block, shard_channels
    from openwebrx_b200.synth import carrier_plan, make_iq
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 50000
        if rank == 0:
            iq = make_iq(n, 2.4e6, carrier_plan(4, 2.4e6, seed=9), seed=9)
            block = torch.from_numpy(iq.view(np.float32).copy())
        else:
            block = torch.zeros(2 * n, dtype=torch.float32)
        for hop in range(3):                       # several hops, the buffer is reused like in bench.py
            if rank == 0:
                block.mul_(1.0)                    # ingest rank owns the data
            broadcast_block(block, 0)
        # the sharded-ingest hop: every rank holds 1/N of the block, afterwards every rank holds all of it
        from openwebrx_b200.sharding import gather_block
        shard = block.numel() // world
        out = torch.zeros_like(block)
        gather_block(out, block[rank * shard:(rank + 1) * shard].clone())
        assert torch.equal(out, block)
        mine = list(shard_channels(n_channels, world, rank))
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, float(block.double().sum()), int(block.numel())))
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_channels", [64, 1024, 7])
def test_two_rank_broadcast_and_sharding(n_channels):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    chans = sorted(c for part, _, _ in got for c in part)
    assert chans == list(range(n_channels))                              # every channel exactly once
    sizes = [len(part) for part, _, _ in got]
    assert max(sizes) - min(sizes) <= 1                                  # balanced
    assert got[0][1] == got[1][1] and got[0][2] == got[1][2]             # identical block on both ranks


def test_owner_matches_partition():
    from openwebrx_b200.sharding import owner_of, shard_channels
    for n, w in ((1024, 8), (1000, 8), (7, 2), (64, 1), (5, 8)):
        for r in range(w):
            for c in shard_channels(n, w, r):
                assert owner_of(c, n, w) == r
