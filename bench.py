#!/usr/bin/env python
"""bench.py -- throughput of the ROMS nonlinear 3-D baroclinic step (main3d chain) on B200.

Metric (BASELINE.json): grid-point-steps/s = Lm*Mm*N*steps / seconds on a synthetic BENCHMARK-shaped grid.
  value : device-resident main3d loop (inputs in HBM), CUDA-event timed, max over ranks
  e2e   : the same metric through the host-facing call roms_b200_step_forced (H2D of the step's surface forcing from
          pinned memory, the step, D2H of the diag scalars) -- host wall clock around the calls
  roofline     : dominant kernel, algorithmic bytes per launch / CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the C++ oracle (one tile per host thread) on a bounded sample of the same workload
  state_digest : sha256 of the owned points of the prognostic fields after spinup + warmup + steps, gathered in global
                 order -- identical at every GPU count (the reference's tiling-invariance criterion, verify.sh:985-1045)
  phase_ms     : per-phase device time taken INSIDE the captured step graph (roms_b200_profile_enable(2)), so it describes
                 the timed configuration at every N and sums to the step
`--impl reference` times the CPU restatement of the reference (the Fortran reference cannot be built here: no Fortran
compiler, no NetCDF) with all host threads, on the same grid, spin-up, warm-up and step count, and prints the same line
with "impl": "reference".
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRIDS = {"benchmark1": (512, 64, 30), "benchmark2": (1024, 128, 30), "benchmark3": (2048, 256, 30),
         # tuning aid: on 2 GPUs this gives every rank the 256x256 tile BENCHMARK3 has on 8 GPUs
         "b3tile8x2": (512, 256, 30),
         # ... the same 256x256 tile as a periodic single-GPU domain (no ring): kernel tuning at the 8-GPU tile size
         "b3tile8": (256, 256, 30),
         # ... and this one the 512x256 tile of 4 GPUs
         "b3tile4x2": (1024, 256, 30)}
# weak scaling: BENCHMARK1/2/3 grow x4 in points per step; grid run at N GPUs
WEAK = {1: "benchmark1", 2: "benchmark2", 4: "benchmark2", 8: "benchmark3"}
METRIC = "grid-point-steps/sec (3D baroclinic)"
DIGEST_FIELDS = ["zeta1", "zeta2", "ubar1", "ubar2", "vbar1", "vbar2", "u1", "u2", "v1", "v2", "t1_0", "t2_0", "t1_1", "t2_1"]
HISTORY_FIELDS = ["zeta1", "ubar1", "vbar1", "u1", "v1", "t1_0", "t1_1", "rho", "W"]     # what a history record holds (wrt_his.F)


# The shipped ROMS/Include/benchmark.h cpp set on top of the reduced one (oracle option names == roms_b200_config members):
# BULK_FLUXES (+LONGWAVE), LMD_MIXING (+RIMIX, CONVEC, SKPP, NONLOCAL, RI_SPLINES), SOLAR_SOURCE, BV_FREQUENCY, MIX_GEO_TS
FULL_BENCHMARK = dict(bv_frequency=1, eos_tderivative=1, solar_source=1, lmd_nonlocal=1, bulk_fluxes=1, lmd_mixing=1, mix_geo_ts=1)


def workload(grid, Lm, Mm, N, ndtfast=20, nfast=29, mix_geo=False, physics="full"):
    """The workload string both arms print (config.workload)."""
    if physics == "full":
        return (f"{grid.upper()} {Lm}x{Mm}x{N}, NT=2, the shipped benchmark.h cpp set (NONLIN_EOS, DJ_GRADPS, U3/C4, UV_VIS2, TS_DIF2 MIX_GEO_TS, UV_QDRAG, "
                f"CURVGRID, SOLAR_SOURCE, BULK_FLUXES+LONGWAVE, LMD_MIXING+RIMIX+CONVEC+SKPP+NONLOCAL+RI_SPLINES; analytical atmosphere), "
                f"ndtfast={ndtfast}, nfast={nfast}")
    mix = "TS_DIF2 MIX_GEO_TS" if mix_geo else "TS_DIF2 MIX_S_TS"
    return (f"{grid.upper()} {Lm}x{Mm}x{N}, NT=2, reduced physics set (nonlinear EOS, DJ_GRADPS, U3/C4, UV_VIS2, {mix}, UV_QDRAG, CURVGRID), "
            f"ndtfast={ndtfast}, nfast={nfast}")


def b_alg_bytes(N, nfast, curvgrid=True, nonlin_eos=True, wvelocity=True, mix_geo=False, full_physics=False):
    """Algorithmic bytes per grid-point-step (SURVEY.md section 8d / BASELINE.md): 8*(U3D + S2D/N).
    full_physics (the shipped benchmark.h set): lmd_vmix reads Hz, u, v, pden, bvf, z_w and writes Akv, Akt x2, ghats x2 (11 units),
    rho_eos also writes bvf (1), pre_step3d also reads ghats x2 and z_w (3); bulk_flux is 2-D only."""
    u3d = 98 + (7 if wvelocity else 0) + (3 if nonlin_eos else 0) + (1 if mix_geo else 0) + (15 if full_physics else 0)
    per_pred, per_corr = 42 + (2 if curvgrid else 0), 39 + (2 if curvgrid else 0)
    s2d = nfast * per_pred + 16 + nfast * per_corr
    return 8.0 * (u3d + s2d / N), s2d


# wclock_on / wclock_off region of each phase (the ids a -DPROFILE build of the reference reports, ROMS/Utility/timers.F with the
# names of ROMS/Modules/mod_strings.F:162-222): phase_wclock maps this bench's phase_ms onto that table.  wvelocity and ana_vmix
# carry no region of their own (they fall into the caller's "residual"); t3dmix is 24 with MIX_S_TS, 25 with MIX_GEO_TS.
WCLOCK = {"set_data": 4, "set_avg": 5, "set_vbc": 6, "diag": 7, "step2d_loop": 9, "set_zeta": 12, "set_depth": 12, "set_massflux": 12, "omega": 13,
          "omega2": 13, "rho_eos": 14, "bulk_flux": 17, "lmd_vmix": 18, "rhs3d": 21, "pre_step3d": 22, "prsgrd": 23, "t3dmix": 24, "uv3dmix": 30,
          "step3d_uv": 34, "step3d_t": 35}


def wclock_ms(prof, mix_geo):
    """phase_ms summed per timers.F region id."""
    out = {}
    for k, v in prof.items():
        r = WCLOCK.get(k)
        if r is None:
            continue
        if k == "t3dmix" and mix_geo:
            r = 25
        out[str(r)] = out.get(str(r), 0.0) + v
    return dict(sorted(out.items(), key=lambda kv: int(kv[0])))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML in-process
    (a sample every 5 ms, so that even a 100 ms timed region is covered); nvidia-smi -lms as the fall-back."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, dev=0):
        self.dev, self.sm, self.mx, self.reasons, self.p, self.run, self.th = dev, [], [], set(), None, False, None

    def _nvml_loop(self, nv, h):
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        self.run = True
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.dev]) if vis and vis.split(",")[self.dev].isdigit() else self.dev
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.th = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.th.start()
            return
        except Exception:
            self.th = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            r = [c.strip() for c in line.split(",")]
            if r and r[0].replace(".", "").isdigit():
                self.sm.append(float(r[0]))
            if len(r) > 1 and r[1].replace(".", "").isdigit():
                self.mx.append(float(r[1]))
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        self.run = False
        if self.th:
            self.th.join(timeout=1.0)
        if self.p:
            self.p.terminate()
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def cpu_tiles(nthreads, Lm, Mm):
    """Factor the host threads into NtileI x NtileJ (reference OpenMP mode: NtileI*NtileJ = threads), NtileI >= NtileJ."""
    best = (max(1, min(nthreads, Lm // 16)), 1)
    for a in range(1, nthreads + 1):
        if nthreads % a == 0:
            b = nthreads // a
            if a >= b and Lm // a >= 8 and Mm // b >= 8:
                best = (a, b)
                break
    return best


def cpu_run(grid, steps, warmup, nthreads, spinup=0, kinds=("fast", "fastmath"), physics="full"):
    """Time the CPU restatement (oracle/) on the host cores: one tile per thread, barrier per phase.  Two timing builds
    (-O3 -march=native, and the same plus the reference's own -ffast-math, Compilers/Linux-gfortran.mk:98-99) are probed on
    two steps each; the faster one runs the sample.  Returns (gp-steps/s, seconds, tiles, build kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    Lm, Mm, N = grid
    ni, nj = cpu_tiles(nthreads, Lm, Mm)
    models, probe = {}, {}
    for k in kinds:
        try:
            models[k] = orc.Oracle(orc.APP_BENCHMARK, Lm=Lm, Mm=Mm, N=N, NtileI=ni, NtileJ=nj, kind=k, **(FULL_BENCHMARK if physics == "full" else {}))
            models[k].step(1, nthreads)
            probe[k] = models[k].timed_steps(2, nthreads)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"bench.py: oracle build '{k}' unavailable: {e}\n")
    kind = min(probe, key=probe.get)
    o = models[kind]
    for k in list(models):
        if k != kind:
            del models[k]
    done = 3                                     # steps the probe already advanced the model
    if spinup > done:
        o.step(spinup - done, nthreads)
    if warmup:
        o.step(warmup, nthreads)
    sec = o.timed_steps(steps, nthreads)
    return Lm * Mm * N * steps / sec, sec, (ni, nj), kind, probe


def cpu_flags(kind):
    return "-O3 -march=native" + (" -ffast-math" if kind == "fastmath" else "")


def gather_global(t, dist, name, rank, world, allb):
    """Owned columns of field `name` of every tile, concatenated along xi on rank 0 (None elsewhere)."""
    import numpy as np
    a = t.get(name)
    b = allb[rank]
    lo = b["Istr"] - b["LBi"]
    own = np.ascontiguousarray(a[..., lo:lo + b["Iend"] - b["Istr"] + 1])
    if dist is None:
        return own
    import torch
    w = max(bb["Iend"] - bb["Istr"] + 1 for bb in allb)
    pad = torch.zeros(own.shape[:-1] + (w,), dtype=torch.float64, device="cuda")
    pad[..., :own.shape[-1]] = torch.from_numpy(own).cuda()
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    return np.concatenate([bufs[r][..., :allb[r]["Iend"] - allb[r]["Istr"] + 1].cpu().numpy() for r in range(world)], axis=-1)


def state_digest(t, dist, rank, world, Lm, Mm):
    from roms_trunk_mgh_b200 import _lib
    allb = [_lib.bounds(Lm, Mm, world, 1, r, distribute=world > 1) for r in range(world)]
    h = hashlib.sha256()
    for n in DIGEST_FIELDS:
        G = gather_global(t, dist, n, rank, world, allb)
        if rank == 0:
            assert G.shape[-1] == Lm, (n, G.shape)
            h.update(n.encode()); h.update(G.tobytes())
    return h.hexdigest() if rank == 0 else None


def make_ring_tile(synth, grid, rank, world, local, dist, **overrides):
    Lm, Mm, N = grid
    t = synth.make_tile(synth.APP_BENCHMARK, Lm, Mm, N, NtileI=world, tile=rank, device=local, **overrides)
    for kv in os.environ.get("ROMS_B200_BENCH_OPTS", "").split():          # tuning aid (tools/): roms_b200_set_option switches
        k, v = kv.split("=")
        t.set_option(k, float(v))
    xchg = "none (single tile, periodic images written by the producing kernel)"
    if world > 1:
        from roms_trunk_mgh_b200 import multigpu
        multigpu.attach(t, dist, rank, world)
        xchg = ("nvlink-peer-mailbox+edge-first-overlap+step2d-exchange-fused-into-kernel" if t.peer else "nccl-send-recv+edge-first-overlap")
        for ph in ("set_depth", "set_massflux", "omega", "rho_eos"):      # start-up phases again, now with live ghosts
            t.run_phase(ph)
    return t, xchg


def timed_steps(t, dist, steps):
    """`steps` resident steps bracketed by barrier + synchronize; device time (CUDA events on the library stream), max over ranks."""
    if dist is not None:
        dist.barrier()
    t.sync()
    t.main3d(steps, sync=True)
    ms = t.last_step_ms()
    if dist is not None:
        dist.barrier()
        import torch
        tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--grid", default="benchmark3")
    ap.add_argument("--spinup", type=int, default=20, help="untimed steps before warm-up so that all upstream branches are live")
    ap.add_argument("--physics", default="full", choices=["full", "reduced"],
                    help="full: the shipped benchmark.h cpp set (bulk_flux + lmd_vmix on the device); reduced: round 1's set (analytical stress, constant mixing)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the MIX_GEO_TS row, the weak-scaling row and the history-write timing")
    a = ap.parse_args()
    # keep stdout clean for the single JSON line: libraries (NCCL's version banner, ...) that print to fd 1 go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")

    def emit(line):
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    Lm, Mm, N = GRIDS[a.grid]
    ncores = os.cpu_count() or 1
    W = max(a.warmup, 3)

    if a.impl == "reference":
        if rank != 0:
            return 0
        nth = ncores
        val, sec, tiles, kind, probe = cpu_run((Lm, Mm, N), a.steps, W, nth, spinup=a.spinup, physics=a.physics)
        sample = (f"{a.steps} steps after {a.spinup} spin-up + {W} warm-up steps on the full {a.grid.upper()} grid, {tiles[0]}x{tiles[1]} tiles on {nth} host "
                  f"threads; C++ restatement of the Fortran (scratch arrays heap-allocated per call), g++ {cpu_flags(kind)} "
                  f"(2-step probe: " + ", ".join(f"{k} {v / 2:.2f} s/step" for k, v in probe.items()) + ")")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "grid-point-steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": W,
                "ms_per_step": 1e3 * sec / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": workload(a.grid, Lm, Mm, N, physics=a.physics), "spinup_steps": a.spinup, "physics": a.physics},
                "cpu_tiles": f"{tiles[0]}x{tiles[1]}",
                "cpu_baseline": {"value": val, "unit": "grid-point-steps/s", "cores": nth, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "grid-point-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import numpy as np
    from roms_trunk_mgh_b200 import synth
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
    NtileI = world
    full = a.physics == "full"
    phys = dict(synth.FULL_BENCHMARK) if full else {}
    assert phys == (FULL_BENCHMARK if full else {})
    t, xchg = make_ring_tile(synth, (Lm, Mm, N), rank, world, local, dist, **phys)
    nfast = synth.set_weights(t.cfg.ndtfast)[0]
    t.main3d(a.spinup)
    t.main3d(W)

    clk = ClockSampler(local); clk.start()
    l0 = t.launch_count()
    ms = timed_steps(t, dist, a.steps)
    launches = t.launch_count() - l0
    clocks = clk.stop()
    value = Lm * Mm * N * a.steps / (ms * 1e-3)

    # ---- digest of the prognostic state after spinup + warmup + steps (before anything else touches the state)
    digest = state_digest(t, dist, rank, world, Lm, Mm)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- e2e: host forcing in, diag scalars out, every step
    g = t.synth["grid"]; b = t.synth["bounds"]
    ke = min(a.steps, 10)
    if full:
        # the set_data products of the shipped set: the atmosphere bulk_flux reads (8 arrays), as of the current model time
        atm = {n: np.ascontiguousarray(synth.tile_slice(v, Lm, b), dtype=np.float64) for n, v in synth.atmosphere_at(g, t.cfg, t.indices()["tdays"]).items()}
        t.register_host(*atm.values())        # what the Fortran host does once for its FORCES(ng) module arrays
        forced = lambda: t.step_fields(atm)   # noqa: E731
        h2d = int(sum(v.nbytes for v in atm.values()))
    else:
        sustr = np.ascontiguousarray(synth.tile_slice(synth.sustr_at(synth.APP_BENCHMARK, g, t.cfg, 0.0), Lm, b), dtype=np.float64)
        svstr = np.zeros_like(sustr); stf = np.zeros_like(sustr)
        t.register_host(sustr, svstr, stf)
        forced = lambda: t.step_forced(sustr, svstr, stf)   # noqa: E731
        h2d = int(3 * sustr.size * 8)
    for _ in range(4):                        # untimed: both time-level parities of the with-diag step get their CUDA graph
        forced()
    barrier(); t.sync(); t0 = time.perf_counter()
    for _ in range(ke):
        d, rc = forced()
    t.sync(); barrier(); e2e_sec = time.perf_counter() - t0
    if dist is not None:
        import torch
        tt = torch.tensor([e2e_sec], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_sec = float(tt.item())
    e2e_val = Lm * Mm * N * ke / e2e_sec

    # ---- what one history record costs (wrt_his.F writes every NHIS steps; the fields come back with roms_b200_get_field)
    hist = None
    if not a.no_extras:
        barrier(); t.sync(); t0 = time.perf_counter()
        nbytes = 0
        for n in HISTORY_FIELDS:
            nbytes += t.get(n).nbytes
        hist = {"fields": HISTORY_FIELDS, "bytes": int(nbytes), "ms": 1e3 * (time.perf_counter() - t0),
                "note": "roms_b200_get_field into pageable numpy arrays, per tile; amortised over NHIS steps (BENCHMARK: NHIS = NTIMES)"}

    # ---- per-phase device times taken inside the captured step graph -> dominant kernel roofline
    NP = 4
    t.profile(2)
    t.main3d(2)                               # both time-level parities get their (marked) graph
    t.profile(2)                              # reset the accumulators
    t.main3d(NP)
    prof, _ = t.profile_get()
    t.profile(0)
    prof = {k: v / NP for k, v in prof.items()}
    if dist is not None:                      # max over ranks, phase by phase
        import torch
        keys = sorted(prof)
        tt = torch.tensor([prof[k] for k in keys], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        prof = dict(zip(keys, [float(x) for x in tt.tolist()]))
    peak, peak_kind = measured_peak()
    balg, s2d = b_alg_bytes(N, nfast, curvgrid=bool(t.cfg.curvgrid), nonlin_eos=bool(t.cfg.nonlin_eos), wvelocity=bool(t.cfg.wvelocity_every_step),
                            mix_geo=full, full_physics=full)
    npts2 = (Lm // NtileI) * Mm
    # units (whole 3-D arrays, or 2-D arrays for step2d) touched per launch group, SURVEY.md section 8a
    units3 = {"pre_step3d": 24 + (3 if full else 0), "rhs3d": 10, "step3d_t": 12, "step3d_uv": 12, "prsgrd": 5, "t3dmix": 8 if full else 7, "uv3dmix": 7,
              "wvelocity": 7, "set_massflux": 5, "rho_eos": (8 if t.cfg.nonlin_eos else 5) + (1 if full else 0), "omega": 4, "omega2": 4, "set_depth": 3,
              "lmd_vmix": 11}
    dom = max(prof, key=prof.get) if prof else None
    roof = None
    if dom:
        if dom == "step2d_loop":
            nl = 2 * nfast + 1
            bytes_per_launch = 8.0 * s2d * npts2 / nl
            dur = prof[dom] * 1e-3 / nl
            kname = "k_step2d"
        else:
            bytes_per_launch = 8.0 * units3.get(dom, 5) * npts2 * N
            dur = prof[dom] * 1e-3
            kname = dom
        ach = bytes_per_launch / dur / 1e9
        traffic, tsrc = None, None
        try:   # DRAM bytes per launch: not measurable in-process; the committed ncu --set full capture of this build (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if world == 1 and a.grid == "benchmark3" and kname in tj:
                traffic = tj[kname]["dram_bytes_per_launch"]; tsrc = tj[kname].get("source")
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": dur * 1e3,
                "launches_per_step": (2 * nfast + 1) if dom == "step2d_loop" else 1,
                "share_of_step": prof[dom] / max(sum(prof.values()), 1e-30)}
    step_gbs = balg * value / 1e9
    line = {"metric": METRIC, "value": value, "unit": "grid-point-steps/s", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(a.grid, Lm, Mm, N, t.cfg.ndtfast, nfast, physics=a.physics), "spinup_steps": a.spinup, "physics": a.physics},
            "tiles": f"{NtileI}x1", "halo_exchange": xchg, "launch": "cuda-graph per step",
            "l2": "working set per step (~3.5 GB) exceeds L2 (126 MB); no explicit flush",
            "e2e": {"value": e2e_val, "unit": "grid-point-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12 * 8, "steps": ke,
                    "api": "roms_b200_step_fields (atmosphere: 8 arrays)" if full else "roms_b200_step_forced (sustr, svstr, stflux)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "roofline_step": {"b_alg_bytes_per_gp_step": balg, "achieved_gbs": step_gbs, "frac": step_gbs / (peak * world)},
            "phase_ms": prof, "phase_ms_sum": sum(prof.values()), "phase_ms_by_wclock_region": wclock_ms(prof, full), "phase_ms_mode": "events inside the captured step graph, main stream, max over ranks",
            "state_digest": digest, "state_digest_steps": a.spinup + W + a.steps, "history_write": hist}
    perr = int(t.L.roms_b200_peer_error(t.h))
    t.close()

    # ---- extra rows: the other physics set (continuity with round 1 / SURVEY.md section 8d "second row") and weak scaling (8e)
    if not a.no_extras and a.grid == "benchmark3":
        try:
            if full:
                # round 1's reduced set (analytical stress, constant vertical mixing, MIX_S_TS fused into pre_step3d)
                tg, _ = make_ring_tile(synth, (Lm, Mm, N), rank, world, local, dist)
                tg.main3d(10)
                msg = timed_steps(tg, dist, 10)
                bg, _ = b_alg_bytes(N, nfast)
                vg = Lm * Mm * N * 10 / (msg * 1e-3)
                line["reduced_physics_row"] = {"workload": workload(a.grid, Lm, Mm, N, t.cfg.ndtfast, nfast, physics="reduced"), "ms_per_step": msg / 10, "value": vg,
                                               "steps": 10, "spinup_steps": 10, "b_alg_bytes_per_gp_step": bg, "roofline_step_frac": bg * vg / 1e9 / (peak * world)}
            else:
                tg, _ = make_ring_tile(synth, (Lm, Mm, N), rank, world, local, dist, **synth.FULL_BENCHMARK)
                tg.main3d(10)
                msg = timed_steps(tg, dist, 6)
                bg, _ = b_alg_bytes(N, nfast, mix_geo=True, full_physics=True)
                vg = Lm * Mm * N * 6 / (msg * 1e-3)
                line["full_benchmark_row"] = {"workload": workload(a.grid, Lm, Mm, N, t.cfg.ndtfast, nfast, physics="full"), "ms_per_step": msg / 6, "value": vg,
                                              "steps": 6, "spinup_steps": 10, "b_alg_bytes_per_gp_step": bg, "roofline_step_frac": bg * vg / 1e9 / (peak * world)}
            perr |= int(tg.L.roms_b200_peer_error(tg.h))
            tg.close()
            wg = WEAK.get(world)
            if wg:
                wl = GRIDS[wg]
                tw, _ = make_ring_tile(synth, wl, rank, world, local, dist, **phys)
                tw.main3d(10)
                msw = timed_steps(tw, dist, 10)
                line["weak_scaling_row"] = {"grid": wg, "n_gpus": world, "ms_per_step": msw / 10, "value": wl[0] * wl[1] * wl[2] * 10 / (msw * 1e-3), "steps": 10,
                                            "note": "BENCHMARK1 on 1, BENCHMARK2 on 2 and 4, BENCHMARK3 on 8 GPUs (x4 points per grid)"}
                perr |= int(tw.L.roms_b200_peer_error(tw.h))
                tw.close()
        except Exception as e:  # noqa: BLE001
            line["extras_error"] = str(e)
    if dist is not None:
        # a neighbour that never delivered its halo invalidates the run
        import torch
        bad = torch.tensor([perr], device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            sys.stderr.write("bench.py: a halo exchange timed out; results are invalid, no line printed\n")
            dist.destroy_process_group()
            return 1
    if rank == 0 and world == 1 and not a.no_cpu:
        try:
            cv, csec, tiles, kind, probe = cpu_run((Lm, Mm, N), 4, 1, ncores, spinup=0, physics=a.physics)
            line["cpu_baseline"] = {"value": cv, "unit": "grid-point-steps/s", "cores": ncores, "kind": "port",
                                    "sample": f"4 steps after 4 warm-up steps on the full {a.grid.upper()} grid, {tiles[0]}x{tiles[1]} tiles on {ncores} host threads, "
                                              f"g++ {cpu_flags(kind)} (the faster of the two timing builds)"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "grid-point-steps/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
