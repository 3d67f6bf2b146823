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


# ---------------------------------------------------------------------------------------------------------------------
# Model of the halo exchange fused into the step2d kernel (csrc/dev.cuh Xchg, csrc/k_step2d.cu): every sub-step kernel k of
# a tile PULLS, per CTA (row block rb), the rows rb-1..rb+1 its two neighbours pushed in their kernel k-1 (slot (k-1) % S of
# the own mailbox, tag k-1), and PUSHES its own rows with tag k into slot k % S of both neighbours' mailboxes.  The only
# ordering the hardware gives is: a tile's kernel k+1 starts after ALL CTAs of its kernel k have finished; CTAs of a kernel
# and different tiles progress in any interleaving.  The model replays random interleavings and checks that a pull never
# finds a LATER epoch in its slot (data overwritten before it was consumed) and that the ring never deadlocks.
def _simulate_fused_exchange(world, K, R, S, seed):
    import random
    rnd = random.Random(seed)
    # box[r][slot][half][rb] = tag; half 0: written by the west neighbour, 1: by the east neighbour
    box = [[[[0] * R for _ in range(2)] for _ in range(S)] for _ in range(world)]
    kern = [1] * world                                  # current kernel of each tile
    state = [{rb: 0 for rb in range(R)} for _ in range(world)]     # per CTA: 0 = must pull, 1 = must push, (absent) = done
    while True:
        if all(k > K for k in kern):
            return "ok"
        moves = [(r, rb) for r in range(world) if kern[r] <= K for rb in state[r]]
        rnd.shuffle(moves)
        progressed = False
        for r, rb in moves:
            k = kern[r]
            if state[r][rb] == 0:
                if k >= 2:
                    need = [b for b in (rb - 1, rb, rb + 1) if 0 <= b < R]
                    tags = [box[r][(k - 1) % S][half][b] for half in (0, 1) for b in need]
                    if any(t > k - 1 for t in tags):
                        return "overwritten"
                    if any(t < k - 1 for t in tags):
                        continue                         # not delivered yet: this CTA spins
                state[r][rb] = 1
            else:
                east, west = (r + 1) % world, (r - 1) % world
                box[east][k % S][0][rb] = k              # my eastern columns -> east neighbour's "from the west" half
                box[west][k % S][1][rb] = k
                del state[r][rb]
                if not state[r]:
                    kern[r] += 1
                    state[r] = {b: 0 for b in range(R)}
            progressed = True
            break
        if not progressed:
            return "deadlock"


@pytest.mark.parametrize("world", [2, 3, 4])
def test_fused_exchange_protocol_model(world):
    for seed in range(150):
        assert _simulate_fused_exchange(world, K=9, R=5, S=4, seed=seed) == "ok"      # XSLOTS = 4 (the library's value)
        assert _simulate_fused_exchange(world, K=9, R=5, S=2, seed=seed) == "ok"      # two slots are already enough
    # the model is able to see the hazard: with a single slot some interleaving overwrites unconsumed data
    assert any(_simulate_fused_exchange(world, K=9, R=5, S=1, seed=seed) == "overwritten" for seed in range(150))


# ---- model of the persistent barotropic-loop kernel's ordering (csrc/k_step2d_loop.cu) ------------------------------------
def _simulate_loop_kernel(nbx, nby, nfast, seed, wait_for_neighbours=True):
    """CTAs of a nbx x nby grid each own one tile for the whole LOOP_2D.  Call c of a CTA READS, at its start, levels krhs(c) and
    kstp(c) of zeta/ubar/vbar on its own tile and on its 8 neighbours (halo), and WRITES level knew(c) of its own tile at its
    end.  The kernel's rule: a CTA starts call c once every neighbour has published call c-1.  The model runs the CTAs under a
    random scheduler (reads and writes of one call are separate events), tags every tile/level with the call that wrote it,
    and checks that each read sees exactly the version a sequential run sees (RAW) -- which also fails if a neighbour has
    already overwritten the level with a later call (WAR)."""
    rng = random.Random(seed)
    # the 2-D time indices of LOOP_2D (main3d.F:592-700), starting with indx1 = 1
    calls, indx1, pred = [], 1, False
    for my_iif in range(1, nfast + 2):
        nxt = 3 - indx1
        if not pred:
            pred, iif = True, my_iif
            kstp, knew, krhs = (indx1 if iif == 1 else 3 - indx1), 3, indx1
        calls.append((krhs, kstp, knew))
        if pred:
            pred, knew = False, nxt
            kstp, krhs = 3 - knew, 3
            if iif < nfast + 1:
                indx1 = nxt
        if iif < nfast + 1:
            calls.append((krhs, kstp, knew))
    ncall = len(calls)
    tiles = [(x, y) for y in range(nby) for x in range(nbx)]
    nbrs = {t: [((t[0] + dx) % nbx, t[1] + dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)
                if 0 <= t[1] + dy < nby and ((t[0] + dx) % nbx, t[1] + dy) != t] for t in tiles}
    # sequential reference: version of (tile, level) seen by call c = last call < c that wrote that level (0 = initial state)
    expect, last = [], {1: 0, 2: 0, 3: 0}
    for c, (krhs, kstp, knew) in enumerate(calls, start=1):
        expect.append((last[krhs], last[kstp]))
        last[knew] = c
    ver = {t: {1: 0, 2: 0, 3: 0} for t in tiles}
    done = {t: 0 for t in tiles}                  # published completion flag
    phase = {t: "idle" for t in tiles}            # idle -> read done ("busy") -> written + published
    cur = {t: 1 for t in tiles}
    while any(cur[t] <= ncall for t in tiles):
        ready = [t for t in tiles if cur[t] <= ncall and
                 (phase[t] == "busy" or not wait_for_neighbours or all(done[n] >= cur[t] - 1 for n in nbrs[t]))]
        if not ready:
            return "deadlock"
        t = rng.choice(ready)
        c = cur[t]
        krhs, kstp, knew = calls[c - 1]
        if phase[t] == "idle":
            for n in nbrs[t] + [t]:
                if (ver[n][krhs], ver[n][kstp]) != expect[c - 1]:
                    return "stale-or-overwritten"
            phase[t] = "busy"
        else:
            ver[t][knew] = c
            done[t] = c
            cur[t] += 1
            phase[t] = "idle"
    return "ok"


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (4, 3), (8, 4)])
def test_loop_kernel_neighbour_flag_ordering_model(shape):
    import random as _r
    globals()["random"] = _r
    for seed in range(60):
        assert _simulate_loop_kernel(shape[0], shape[1], nfast=6, seed=seed) == "ok"
    if shape != (1, 1):
        # the model sees the hazards: without the neighbour wait some interleaving reads a stale or already overwritten level
        assert any(_simulate_loop_kernel(shape[0], shape[1], nfast=6, seed=seed, wait_for_neighbours=False) != "ok" for seed in range(60))
