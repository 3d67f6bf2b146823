"""Host-side plumbing for the NtileI x 1 ring of tiles (one process per GPU, torch.distributed for the rendezvous only).

The halo exchange itself (mp_exchange2d/3d/4d, ROMS/Utility/mp_exchange.F:1413-2128) runs inside the CUDA library with
NCCL send/recv over NVLink (csrc/api_nccl.cu).  This module only (a) distributes the 128-byte NCCL unique id, (b) attaches
the communicator to a Tile, (c) scatters / gathers whole global arrays for set-up and verification, and (d) provides a
numpy model of the ring exchange that the CPU (gloo) tests run.
"""
import ctypes as C

import numpy as np

NW, NE = 3, 2          # west / east ghost columns carried on the device (csrc/api_nccl.cu)


def ring_neighbours(rank, nranks):
    """(west, east) ranks of a tile on the periodic xi ring (tile_neighbors, mp_exchange.F:73-286)."""
    return (rank + nranks - 1) % nranks, (rank + 1) % nranks


def broadcast_bytes(dist, payload, src=0):
    """Broadcast a bytes object from `src` with any torch.distributed backend."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = torch.tensor([len(payload) if dist.get_rank() == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def attach(tile, dist, rank, world, peer=True):
    """Create the NCCL ring communicator for `tile` (rank == tile index) and fill all ghost columns; then, unless
    peer=False, move the exchanges to the NVLink peer path.  Collective."""
    L = tile.L
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        rc = L.roms_b200_nccl_unique_id(idbuf)
        if rc:
            raise RuntimeError(f"roms_b200_nccl_unique_id -> {rc}")
    payload = broadcast_bytes(dist, idbuf.raw, 0)
    comm = C.c_void_p()
    rc = L.roms_b200_nccl_init_rank(payload, rank, world, C.byref(comm))
    if rc:
        raise RuntimeError(f"roms_b200_nccl_init_rank -> {rc}")
    rc = L.roms_b200_attach_nccl(tile.h, comm, rank, world)
    if rc:
        raise RuntimeError(f"roms_b200_attach_nccl -> {rc}")
    tile._nccl_comm = comm
    tile.peer = False
    if peer:
        tile.peer = enable_peer(tile, dist, rank, world)
    return comm


def enable_peer(tile, dist, rank, world):
    """Switch the ring exchanges of `tile` to the NVLink peer path (CUDA IPC mailboxes written by the neighbours' kernels).
    Collective; falls back to NCCL send/recv on every rank unless every rank could map both neighbours.  Returns the mode."""
    import torch
    L = tile.L
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    hbuf = C.create_string_buffer(64)
    ok = L.roms_b200_peer_export(tile.h, hbuf) == 0
    mine = torch.frombuffer(bytearray(hbuf.raw), dtype=torch.uint8).to(dev)
    allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine)
    west, east = ring_neighbours(rank, world)
    if ok:
        ok = L.roms_b200_peer_attach(tile.h, bytes(allh[west].cpu().numpy().tobytes()), bytes(allh[east].cpu().numpy().tobytes())) == 0
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
    if ok:
        ok = L.roms_b200_peer_enable(tile.h, 1) == 0
    dist.barrier()
    return ok


def interior_columns(bounds):
    """Column slice of a tile-local array (origin LBi) that holds the tile's own points Istr..Iend."""
    lo = bounds["Istr"] - bounds["LBi"]
    return slice(lo, lo + bounds["Iend"] - bounds["Istr"] + 1)


def assemble_global(parts, all_bounds, Lm):
    """Rebuild a global array A(-2:Lm+2, ...) from the tiles' own columns, then fill the periodic images."""
    shape = parts[0].shape[:-1] + (Lm + 5,)
    G = np.zeros(shape)
    for a, b in zip(parts, all_bounds):
        G[..., b["Istr"] + 2:b["Iend"] + 3] = a[..., interior_columns(b)]
    G[..., Lm + 3:Lm + 5] = G[..., 3:5]
    G[..., 0:3] = G[..., Lm:Lm + 3]
    return G


def ring_exchange_numpy(dist, a, bounds, rank, world):
    """numpy model of one halo exchange on the tile-local array `a` (origin LBi = Istr-2, or -2 on tile 0):
    eastward message = last NW own columns, westward message = first NE own columns (same order as k_pack)."""
    import torch
    west, east = ring_neighbours(rank, world)
    own = interior_columns(bounds)
    lo, hi = own.start, own.stop
    nw_have = lo                                 # west ghost columns present in the host-shaped array (2, or 3 on tile 0)
    sendE = np.ascontiguousarray(a[..., hi - NW:hi]); sendW = np.ascontiguousarray(a[..., lo:lo + NE])
    recvW = np.empty_like(sendE); recvE = np.empty_like(sendW)
    reqs = [dist.isend(torch.from_numpy(sendE), east), dist.irecv(torch.from_numpy(recvW), west),
            dist.isend(torch.from_numpy(sendW), west), dist.irecv(torch.from_numpy(recvE), east)]
    for r in reqs:
        r.wait()
    a[..., lo - nw_have:lo] = recvW[..., NW - nw_have:]
    a[..., hi:hi + NE] = recvE
    return a
