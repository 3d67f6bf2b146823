"""CPU test of the multi-tile host logic with torch.distributed (gloo, world_size 2 and 4): partition bounds, tile
slicing, the ring neighbour table with the periodic wrap, the message order of the halo exchange (3 columns eastward, 2
westward -- the numpy model of csrc/api_nccl.cu) and the unique-id broadcast helper."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
from roms_trunk_mgh_b200 import multigpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, Lm, Mm, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        G = rng.standard_normal((3, Mm + 2, Lm + 5))                    # global A(-2:Lm+2, 0:Mm+1, 3 planes)
        G[..., Lm + 3:Lm + 5] = G[..., 3:5]; G[..., 0:3] = G[..., Lm:Lm + 3]   # periodic images
        b = orc.bounds(Lm, Mm, world, 1, rank, distribute=True)
        a = np.ascontiguousarray(G[..., b["LBi"] + 2:b["UBi"] + 3])
        own = multigpu.interior_columns(b)
        a[..., :own.start] = np.nan; a[..., own.stop:] = np.nan           # forget the ghosts
        multigpu.ring_exchange_numpy(dist, a, b, rank, world)
        ok = np.array_equal(a, G[..., b["LBi"] + 2:b["UBi"] + 3])
        w, e = multigpu.ring_neighbours(rank, world)
        ok = ok and w == (rank - 1) % world and e == (rank + 1) % world
        payload = multigpu.broadcast_bytes(dist, bytes(range(128)) if rank == 0 else b"", 0)
        ok = ok and payload == bytes(range(128))
        # gather + reassembly used by the multi-GPU verification
        parts = [None] * world
        dist.all_gather_object(parts, a)
        allb = [orc.bounds(Lm, Mm, world, 1, r, distribute=True) for r in range(world)]
        ok = ok and np.array_equal(multigpu.assemble_global(parts, allb, Lm), G)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Lm,Mm", [(2, 41, 12), (4, 64, 8), (2, 512, 4)])
def test_ring_exchange_gloo(world, Lm, Mm):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Lm, Mm, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
