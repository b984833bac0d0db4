#!/usr/bin/env python
"""The per-block all-gather ALONE (no DSP kernels beside it): device time per hop for each transport, N ranks under torchrun.
  torchrun --nproc-per-node 8 tools/hop_probe.py [log2 block samples]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openwebrx_b200.sharding import make_hop                     # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    block = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 25)
    shard = block // world
    mine = torch.randn(shard, 2, device=dev)
    st = torch.cuda.Stream(device=dev)
    out = []
    kinds = (("multicast", 0), ("nccl", 0)) if os.environ.get("OWRX_HOP_PROBE") == "mc" else (
        ("pull", 1), ("pull", 4), ("pull", 7), ("push", 1), ("push", 7), ("multicast", 0), ("nccl", 0))
    for kind, lanes in kinds:
        os.environ["OWRX_HOP_LANES"] = str(max(lanes, 1))
        hop = make_hop(kind, block, world, rank, dev)
        for j in range(25):
            hop.gather(j, mine)
            hop.recv(j, st)
            hop.release(j, st)
        st.synchronize()
        ms = hop.mean_ms()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        inb = block * 8.0 * (world - 1) / world
        out.append(dict(kind=kind, lanes=lanes, ms=float(t.item()), GBps_in_per_gpu=inb / (float(t.item()) * 1e-3) / 1e9))
        del hop
        torch.cuda.empty_cache()
        dist.barrier()
    if rank == 0:
        print(json.dumps(dict(world=world, block_samples=block, hops=out)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
