#!/usr/bin/env python
"""Peer receive buffers (engine.peer_receive_buffers): can a kernel of this rank store into another rank's buffer?
usage: torchrun --nproc-per-node 2 tools/p2p_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pymn_b200 import _lib, engine

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = engine.peer_receive_buffers(1 << 20, dev, rank, world)
print(rank, "peer buffers:", "ok" if res is not None else ("FAILED " + str(engine._PEER_ERR)), flush=True)
if res is not None:
    recv, table = res
    ent = engine._PEER[local]
    print(rank, "ptrs", [hex(int(p.data_ptr())) for p in ent["peers"]], "devices", [str(p.device) for p in ent["peers"]],
          flush=True)
    n = 1000
    src = torch.arange(n, dtype=torch.float32, device=dev) + 1000.0 * rank
    off = torch.arange(n, dtype=torch.int64, device=dev)
    dist.barrier()
    torch.cuda.synchronize()
    for r in range(world):                       # write my vector into everybody's buffer at slot `rank`
        dst = ent["peers"][r].view(torch.float32)[rank * n:(rank + 1) * n]
        _lib.call("mn_votes_unpack", src, off, n, 1, n, dst, n)
    torch.cuda.synchronize()
    dist.barrier()
    got = recv[:world * n].view(world, n).cpu()
    ok = all(torch.equal(got[r], torch.arange(n, dtype=torch.float32) + 1000.0 * r) for r in range(world))
    print(rank, "kernel peer stores:", "ok" if ok else "WRONG", flush=True)
dist.barrier()
dist.destroy_process_group()
