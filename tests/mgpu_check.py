"""Multi-GPU tiling-invariance check (run under torchrun on N GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank owns one xi-tile of a BENCHMARK-shaped grid (NtileI = N, NtileJ = 1) and steps it with NCCL halo exchanges; rank 0
also steps the whole domain as a single tile.  The reference's acceptance criterion (ROMS/Bin/verify.sh:985-1045) is that
results do not depend on the tiling: all prognostic fields must agree BITWISE."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roms_trunk_mgh_b200 import _lib, multigpu, synth  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    pos = [a for a in sys.argv[1:] if "=" not in a]
    opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)          # roms_b200_set_option switches, e.g. step2d_exchange=0
    Lm, Mm, N = (int(x) for x in (pos[0:3] if len(pos) >= 3 else (256, 64, 30)))
    nsteps = int(pos[3]) if len(pos) >= 4 else 6
    peer = opts.pop("peer", "1") != "0"
    opts_e2e = opts.pop("e2e", "0") != "0"     # step through roms_b200_step_fields / step_forced (uploads without halo exchange)
    # physics=full: the shipped benchmark.h cpp set (bulk_flux + lmd_vmix on the device, synth.FULL_BENCHMARK)
    phys = dict(synth.FULL_BENCHMARK) if opts.pop("physics", "reduced") == "full" else {}
    # cpp variants of the chain (roms_b200_config members): uv_adv=1|2|3, dj_gradps=0..3, ts_dif4=1 tnu4=..., vadv=3, ...
    for k in ("uv_adv", "dj_gradps", "ts_dif4", "hadv", "vadv", "uv_qdrag", "mix_geo_ts", "nonlin_eos"):
        if k in opts:
            phys[k] = int(opts.pop(k))
    if "tnu4" in opts:
        phys["tnu4"] = float(opts.pop("tnu4"))
    t = synth.make_tile(synth.APP_BENCHMARK, Lm, Mm, N, NtileI=world, tile=rank, device=local, **phys)
    for k, v in opts.items():
        t.set_option(k, float(v))
    multigpu.attach(t, dist, rank, world, peer=peer)
    if rank == 0:
        print(f"exchange path: {'NVLink peer mailboxes' if t.peer else 'NCCL send/recv'}; options {opts}", flush=True)
    # redo the start-up phases now that ghosts can be exchanged (make_tile ran them before the ring existed)
    for ph in ("set_depth", "set_massflux", "omega", "rho_eos"):
        t.run_phase(ph)
    e2e = opts_e2e

    def advance(tt, n):
        """n steps: resident, or (e2e=1) through the forced-step API with this tile's slice of the forcing uploaded every step"""
        if not e2e:
            tt.main3d(n)
            return
        g, b = tt.synth["grid"], tt.synth["bounds"]
        for _ in range(n):
            if phys:
                atm = {k: synth.tile_slice(v, Lm, b) for k, v in synth.atmosphere_at(g, tt.cfg, tt.indices()["time"] / 86400.0).items()}
                tt.step_fields(atm)
            else:
                su = synth.tile_slice(synth.sustr_at(synth.APP_BENCHMARK, g, tt.cfg, 0.0), Lm, b)
                tt.step_forced(su, np.zeros_like(su), np.zeros_like(su))

    advance(t, nsteps)
    d = t.diag()
    names = ["zeta1", "zeta2", "ubar1", "vbar1", "u1", "u2", "v1", "v2", "t1_0", "t2_0", "t1_1", "t2_1", "Huon", "Hvom", "W", "rho"]
    if phys:
        names += ["Akv", "Akt_0", "Akt_1", "hsbl", "sustr", "svstr", "stflux_0", "lrflx", "lhflx", "shflx", "bvf"]
    allb = [_lib.bounds(Lm, Mm, world, 1, r, distribute=True) for r in range(world)]
    ok = True
    ref = None
    if rank == 0:
        ref = synth.make_tile(synth.APP_BENCHMARK, Lm, Mm, N, device=local, **phys)
        advance(ref, nsteps)
    for n in names:
        a = torch.from_numpy(t.get(n)).cuda()
        # gather variable-width tiles: pad to the widest
        w = max(b["UBi"] - b["LBi"] + 1 for b in allb)
        pad = torch.zeros(a.shape[:-1] + (w,), dtype=a.dtype, device="cuda"); pad[..., :a.shape[-1]] = a
        bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bufs, dst=0)
        if rank == 0:
            parts = [bufs[r][..., :allb[r]["UBi"] - allb[r]["LBi"] + 1].cpu().numpy() for r in range(world)]
            G = multigpu.assemble_global(parts, allb, Lm)
            R = ref.get(n)
            same = np.array_equal(G, R)
            if not same:
                ok = False
                dd = np.abs(G - R)
                print(f"  {n}: MISMATCH max {dd.max():.3e} of {np.abs(R).max():.3e} at {np.unravel_index(dd.argmax(), dd.shape)}", flush=True)
    if rank == 0:
        dr = ref.diag()
        print("diag tiled :", {k: f"{v:.12e}" for k, v in d.items()})
        print("diag single:", {k: f"{v:.12e}" for k, v in dr.items()})
        # diag.F: maxima (and the Courant components AT the largest Courant number) do not depend on the tiling; the three
        # sums are added tile by tile, so only their last bits may differ from the single-tile summation order
        for k in d:
            if k in ("avgke", "avgpe", "avgkp", "volume"):
                if abs(d[k] - dr[k]) > 1e-13 * abs(dr[k]):
                    ok = False; print(f"  diag {k}: {d[k]!r} vs {dr[k]!r}", flush=True)
            elif d[k] != dr[k]:
                ok = False; print(f"  diag {k}: {d[k]!r} vs {dr[k]!r}", flush=True)
        print(f"MGPU_CHECK world={world} e2e={int(e2e)} physics={'full' if phys else 'reduced'} grid={Lm}x{Mm}x{N} steps={nsteps}:", "BITWISE-IDENTICAL" if ok else "FAILED", flush=True)
    perr = t.L.roms_b200_peer_error(t.h)
    if perr:
        print(f"rank {rank}: peer exchange timed out", flush=True)
    flag = torch.tensor([0 if ok and not perr else 1], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
