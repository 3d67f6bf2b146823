#!/usr/bin/env python
"""bench.py -- throughput of the ROMS nonlinear 3-D baroclinic step (main3d chain) on B200.

Metric (BASELINE.json): grid-point-steps/s = Lm*Mm*N*steps / seconds on a synthetic BENCHMARK-shaped grid.
  value : device-resident main3d loop (inputs in HBM), CUDA-event timed, max over ranks
  e2e   : the same metric through the host-facing call roms_b200_step_forced (H2D of the step's surface forcing from
          pinned memory, the step, D2H of the diag scalars) -- host wall clock around the calls
  roofline     : dominant kernel, algorithmic bytes per launch / CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the C++ oracle (-O3 -march=native, one tile per host thread) on a bounded sample of the same workload
`--impl reference` times the CPU restatement of the reference (the Fortran reference cannot be built here: no Fortran
compiler, no NetCDF) with all host threads and prints the same line with "impl": "reference".
"""
import argparse
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
         # ... and this one the 512x256 tile of 4 GPUs
         "b3tile4x2": (1024, 256, 30)}
METRIC = "grid-point-steps/sec (3D baroclinic)"


def b_alg_bytes(N, nfast, curvgrid=True, nonlin_eos=True, wvelocity=True):
    """Algorithmic bytes per grid-point-step (SURVEY.md section 8d / BASELINE.md): 8*(U3D + S2D/N)."""
    u3d = 98 + (7 if wvelocity else 0) + (3 if nonlin_eos else 0)
    per_pred, per_corr = 42 + (2 if curvgrid else 0), 39 + (2 if curvgrid else 0)
    s2d = nfast * per_pred + 16 + nfast * per_corr
    return 8.0 * (u3d + s2d / N), s2d


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


def cpu_run(grid, steps, warmup, nthreads, kind="fast"):
    """Time the CPU restatement (oracle/) on the host cores: one tile per thread, barrier per phase."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    Lm, Mm, N = grid
    ni = max(1, min(nthreads, Lm // 16))
    nj = 1
    # factor threads into NtileI x NtileJ (reference OpenMP mode: NtileI*NtileJ = threads)
    best = (ni, 1)
    for a in range(1, nthreads + 1):
        if nthreads % a == 0:
            b = nthreads // a
            if a >= b and Lm // a >= 8 and Mm // b >= 8:
                best = (a, b)
                break
    ni, nj = best
    o = orc.Oracle(orc.APP_BENCHMARK, Lm=Lm, Mm=Mm, N=N, NtileI=ni, NtileJ=nj, kind=kind)
    if warmup:
        o.step(warmup, nthreads)
    sec = o.timed_steps(steps, nthreads)
    return Lm * Mm * N * steps / sec, sec, (ni, nj)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--grid", default="benchmark3")
    ap.add_argument("--spinup", type=int, default=20, help="untimed steps before warm-up so that all upstream branches are live")
    ap.add_argument("--no-cpu", action="store_true")
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
        nth = min(ncores, 32)
        steps = max(1, min(a.steps, 3)); warm = min(W, 1)
        val, sec, tiles = cpu_run((Lm, Mm, N), steps, warm, nth)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "grid-point-steps/s", "n_gpus": a.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": f"{a.grid.upper()} {Lm}x{Mm}x{N}, reduced physics set", "tiles": f"{tiles[0]}x{tiles[1]}"},
                "cpu_baseline": {"value": val, "unit": "grid-point-steps/s", "cores": nth, "kind": "port",
                                 "sample": f"{steps} steps after {warm} warm-up on the full {a.grid.upper()} grid, {tiles[0]}x{tiles[1]} tiles on {nth} threads"},
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
    xchg = "none (single tile, periodic images written by the producing kernel)"
    t = synth.make_tile(synth.APP_BENCHMARK, Lm, Mm, N, NtileI=NtileI, tile=rank, device=local)
    nfast = synth.set_weights(t.cfg.ndtfast)[0]
    if world > 1:
        from roms_trunk_mgh_b200 import multigpu
        multigpu.attach(t, dist, rank, world)
        xchg = ("nvlink-peer-mailbox" if t.peer else "nccl-send-recv") + ("" if os.environ.get("ROMS_B200_NO_OVERLAP") == "1" else "+edge-first-overlap")
        if t.peer and os.environ.get("ROMS_B200_FUSED_XCHG", "2") != "0":
            xchg += "+step2d-exchange-fused-into-kernel"
        for ph in ("set_depth", "set_massflux", "omega", "rho_eos"):      # start-up phases again, now with live ghosts
            t.run_phase(ph)
    t.main3d(a.spinup)
    t.main3d(W)

    def barrier():
        if dist is not None:
            dist.barrier()

    clk = ClockSampler(local); clk.start()
    barrier(); t.sync()
    l0 = t.launch_count()
    t.main3d(a.steps, sync=True)
    ms = t.last_step_ms()
    barrier()
    launches = t.launch_count() - l0
    clocks = clk.stop()
    if dist is not None:
        import torch
        tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
    value = Lm * Mm * N * a.steps / (ms * 1e-3)

    # ---- e2e: host forcing in, diag scalars out, every step
    g = t.synth["grid"]; b = t.synth["bounds"]
    sustr = np.ascontiguousarray(synth.tile_slice(synth.sustr_at(synth.APP_BENCHMARK, g, t.cfg, 0.0), Lm, b), dtype=np.float64)
    svstr = np.zeros_like(sustr); stf = np.zeros_like(sustr)
    ke = min(a.steps, 10)
    t.register_host(sustr, svstr, stf)        # what the Fortran host does once for its FORCES(ng) module arrays
    for _ in range(4):                        # untimed: both time-level parities of the with-diag step get their CUDA graph
        t.step_forced(sustr, svstr, stf)
    barrier(); t.sync(); t0 = time.perf_counter()
    for _ in range(ke):
        d, rc = t.step_forced(sustr, svstr, stf)
    t.sync(); barrier(); e2e_sec = time.perf_counter() - t0
    if dist is not None:
        import torch
        tt = torch.tensor([e2e_sec], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_sec = float(tt.item())
    e2e_val = Lm * Mm * N * ke / e2e_sec

    # ---- per-phase device times (CUDA events on the library's stream) -> dominant kernel roofline
    t.profile(True)
    t.main3d(3)
    prof, _ = t.profile_get()
    t.profile(False)
    prof = {k: v / 3.0 for k, v in prof.items()}
    peak, peak_kind = measured_peak()
    balg, s2d = b_alg_bytes(N, nfast, curvgrid=bool(t.cfg.curvgrid), nonlin_eos=bool(t.cfg.nonlin_eos), wvelocity=bool(t.cfg.wvelocity_every_step))
    npts2 = (Lm // NtileI) * Mm
    # units (whole 3-D arrays, or 2-D arrays for step2d) touched per launch group, SURVEY.md section 8a
    units3 = {"pre_step3d": 24, "rhs3d": 10, "step3d_t": 12, "step3d_uv": 12, "prsgrd": 5, "t3dmix": 7, "uv3dmix": 7, "wvelocity": 7,
              "set_massflux": 5, "rho_eos": 8 if t.cfg.nonlin_eos else 5, "omega": 4, "omega2": 4, "set_depth": 3}
    dom = max(prof, key=prof.get) if prof else None
    roof = None
    if dom:
        if dom == "step2d_loop":
            bytes_per_launch = 8.0 * s2d * npts2 / (2 * nfast + 1)
            dur = prof[dom] * 1e-3 / (2 * nfast + 1)
            kname = "k_step2d"
        else:
            bytes_per_launch = 8.0 * units3.get(dom, 5) * npts2 * N
            dur = prof[dom] * 1e-3
            kname = dom
        ach = bytes_per_launch / dur / 1e9
        traffic = None
        try:   # DRAM bytes per launch from the committed ncu --set full capture (profiles/), single-GPU full-size tile only
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if world == 1 and a.grid == "benchmark3" and kname in tj:
                traffic = tj[kname]["dram_bytes_per_launch"]
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": dur * 1e3, "share_of_step": prof[dom] / max(sum(prof.values()), 1e-30)}
    step_gbs = balg * value / 1e9
    line = {"metric": METRIC, "value": value, "unit": "grid-point-steps/s", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{a.grid.upper()} {Lm}x{Mm}x{N}, NT=2, reduced physics set (nonlinear EOS, DJ_GRADPS, U3/C4, UV_VIS2, TS_DIF2 MIX_S_TS, "
                                   f"UV_QDRAG, CURVGRID), ndtfast={t.cfg.ndtfast}, nfast={nfast}", "tiles": f"{NtileI}x1", "halo_exchange": xchg,
                       "launch": "stream" if os.environ.get("ROMS_B200_NO_GRAPH") == "1" else "cuda-graph per step",
                       "l2": "working set per step (~3.5 GB) exceeds L2 (126 MB); no explicit flush", "spinup_steps": a.spinup},
            "e2e": {"value": e2e_val, "unit": "grid-point-steps/s", "h2d_bytes_per_step": int(3 * sustr.size * 8), "d2h_bytes_per_step": 12 * 8,
                    "steps": ke},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "roofline_step": {"b_alg_bytes_per_gp_step": balg, "achieved_gbs": step_gbs, "frac": step_gbs / (peak * world)},
            "phase_ms": prof}
    if dist is not None:
        # a neighbour that never delivered its halo (NVLink peer path: the spin gives up after ~3 s) invalidates the run
        import torch
        bad = torch.tensor([int(t.L.roms_b200_peer_error(t.h))], device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            sys.stderr.write("bench.py: a halo exchange timed out; results are invalid, no line printed\n")
            dist.destroy_process_group()
            return 1
    if rank == 0 and world == 1 and not a.no_cpu:
        try:
            nth = min(ncores, 32)
            cv, csec, tiles = cpu_run((Lm, Mm, N), 2, 1, nth)
            line["cpu_baseline"] = {"value": cv, "unit": "grid-point-steps/s", "cores": nth, "kind": "port",
                                    "sample": f"2 steps after 1 warm-up on the full {a.grid.upper()} grid, {tiles[0]}x{tiles[1]} tiles on {nth} host threads"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "grid-point-steps/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
