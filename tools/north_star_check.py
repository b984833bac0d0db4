#!/usr/bin/env python
"""The north-star load, paced at real time, on N GPUs of one box (run under torchrun):

  61.44 MS/s complex-float IQ arrives on rank 0 in 1/30 s blocks (pinned host memory) -> H2D -> NCCL broadcast to every GPU;
  every GPU runs its shard of the client channels (128 x 12 kHz NFM/AM/USB each: 1024 on 8 GPUs) through
  owrx_bank_process_device + owrx_bank_drain and pops every channel's audio on the host; rank 0 also runs the 65536-point,
  30 fps waterfall through the host API (ADPCM lines read back).

Rank 0 paces the blocks at the true rate for --seconds and every rank records how long a block keeps it busy against the
33.3 ms budget.  Rank 0 prints one JSON object (max over ranks)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwebrx_b200 import ChannelBank, Waterfall                       # noqa: E402
from openwebrx_b200.sharding import broadcast_block, shard_channels     # noqa: E402
from openwebrx_b200.synth import BANDPASS, carrier_plan                 # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--channels-per-gpu", type=int, default=128)
    ap.add_argument("--fs", type=float, default=61.44e6)
    ap.add_argument("--fps", type=int, default=30)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fs, fps = args.fs, args.fps
    block = int(fs / fps)
    n_blocks = int(args.seconds * fps)
    total = args.channels_per_gpu * world
    mine = shard_channels(total, world, rank)
    cars = carrier_plan(total, fs, seed=7)
    bank = ChannelBank(fs, device=local)
    chans = [bank.add_channel(12000, demod=cars[c]["kind"], offset=cars[c]["offset"], bandpass=BANDPASS[cars[c]["kind"]]) for c in mine]
    audio = np.empty((len(chans), 1 << 12), np.float32)
    # the device path consumes whole decimation steps of what it is given and keeps no wideband history: the caller hands it
    # a window [carry | new block] and carries the unconsumed tail (< T + D samples) over to the next block
    win = torch.empty(block + (1 << 18), 2, device=dev)
    carry = [0]
    ring, wf = [], None
    if rank == 0:
        g = torch.Generator(); g.manual_seed(1)
        t = torch.arange(block, dtype=torch.float32)
        for k in range(8):
            x = 1e-3 * torch.randn(block, 2, generator=g)
            x[:, 0] += 0.1 * torch.cos(0.01 * (k + 1) * t); x[:, 1] += 0.1 * torch.sin(0.01 * (k + 1) * t)
            ring.append(x.pin_memory())
        wf = Waterfall(fs, 65536, 0.3, fps, "adpcm", device=local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def one(k):
        buf = win[carry[0]:carry[0] + block]
        if rank == 0:
            buf.copy_(ring[k % len(ring)], non_blocking=True)                # ingest: H2D of the block
        if world > 1:
            broadcast_block(buf, 0)                                          # the hop (NVLink)
            if rank != 0:
                stream.synchronize()                                         # the block is here: this rank's clock starts
        t_have = time.perf_counter()
        n = carry[0] + block
        before = bank.stats()["channel_samples"]
        bank.process_device(win, n, stream=stream.cuda_stream)
        consumed = (bank.stats()["channel_samples"] - before) // len(chans)
        carry[0] = n - consumed
        assert 0 <= carry[0] <= (1 << 18)
        win[:carry[0]].copy_(win[consumed:n], non_blocking=True)
        bank.drain()
        got = sum(bank.read_audio_all(chans, audio))
        lines = len(wf.feed(ring[k % len(ring)].numpy().view(np.complex64).reshape(-1))) if rank == 0 else 0
        return got, lines, t_have

    for k in range(4):
        one(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    busy, n_audio, n_lines, late = [], 0, 0, 0
    period = 1.0 / fps
    t0 = time.perf_counter()
    for k in range(n_blocks):
        s = time.perf_counter()
        if rank == 0:
            deadline = t0 + k * period
            if s < deadline:
                time.sleep(deadline - s)
            elif s - deadline > period:
                late += 1
            s = time.perf_counter()
        a, l, t_have = one(k)
        # rank 0: everything from the block's arrival; other ranks: from the moment the hop call returned
        busy.append(time.perf_counter() - (s if rank == 0 else t_have))
        n_audio += a; n_lines += l
    wall = time.perf_counter() - t0
    busy = np.asarray(busy) * 1e3
    stats = torch.tensor([busy.mean(), np.percentile(busy, 99), busy.max(), float(late), n_audio / max(len(chans), 1) / wall], device=dev, dtype=torch.float64)
    if world > 1:
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
        allst = torch.stack(allst).cpu().numpy()
    else:
        allst = stats.cpu().numpy()[None]
    if rank == 0:
        print(json.dumps({
            "workload": "%d x 12 kHz channels on %d GPU(s) (%d each) + 65536-pt %d fps waterfall on rank 0, from %.2f MS/s in %d blocks of 1/%d s; "
                        "H2D on rank 0, NCCL broadcast, audio of every channel read on the host" % (total, world, args.channels_per_gpu, fps, fs / 1e6, n_blocks, fps),
            "budget_ms_per_block": period * 1e3, "block_ms_mean_rank0": float(allst[0, 0]), "block_ms_p99_rank0": float(allst[0, 1]),
            "block_ms_max_rank0": float(allst[0, 2]), "block_ms_mean_max_over_ranks": float(allst[:, 0].max()),
            "block_ms_p99_max_over_ranks": float(allst[:, 1].max()), "headroom_x": float(period * 1e3 / allst[:, 0].max()),
            "late_blocks": int(allst[0, 3]), "audio_samples_per_channel_per_s_min_over_ranks": float(allst[:, 4].min()),
            "waterfall_lines_per_s": n_lines / wall, "wall_s": wall}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
