"""CPU check of the CUDA kernels' SOURCE (no GPU needed): tests/emu compiles the unmodified text of the kernels of the chain
(csrc/k_glue.cu, k_pre.cu, k_rhs.cu, k_step3d.cu, k_mixgeo.cu, k_step2d.cu: set_massflux, rho_eos, set_vbc, ana_vmix, omega,
wvelocity, set_zeta, pre_step3d, prsgrd31/32, t3dmix2_s, t3dmix2_geo, t3dmix4_s, rhs3d, uv3dmix2, step2d, set_depth, step3d_uv,
step3d_t, bvf_mix) for the host and runs every thread of every launch -- in turn, or as a pool of host threads with a real barrier
for the kernels that call __syncthreads().  Built like the oracle's parity build (-O2 -ffp-contract=off), each phase must reproduce
the oracle BIT FOR BIT from the oracle's own inputs: loop ranges, wall / periodic-image handling, upstream selects, shared-memory
tile indexing and the operation order of the kernel text are pinned without a device.  What it cannot see: the cooperative loop
kernel (k_step2d_loop), the multi-GPU exchange, device libm, data races.  The -m gpu tests remain the parity
tests proper."""
import os
import sys

import numpy as np
import pytest

import orc
from helpers import all_names, fill_flux_data, optional_names

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from emu import EMULATED, EmuTile  # noqa: E402

STEP_PHASES = ["set_massflux", "rho_eos", "set_vbc", "ana_vmix", "omega", "wvelocity", "set_zeta", "pre_step3d", "prsgrd", "t3dmix",
               "rhs3d", "uv3dmix", "step2d_loop", "set_depth", "step3d_uv", "omega2", "step3d_t"]
CASES = {
    "seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8)),
    "upwelling": (orc.APP_UPWELLING, dict(Lm=20, Mm=24, N=8)),
    "benchmark": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7)),
    "benchmark_p31": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, dj_gradps=0, nonlin_eos=0)),
    "benchmark_p40": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, dj_gradps=2)),                       # PJ_GRADP (prsgrd40.h)
    "seamount_p40": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, dj_gradps=2)),
    "benchmark_wj": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, dj_gradps=3)),                        # WJ_GRADP (prsgrd31.h, weighted)
    "seamount_wj": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, dj_gradps=3)),
    "benchmark_limit": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, limit_bstress=1, uv_qdrag=0, rdrg=5.0)),   # LIMIT_BSTRESS, with a linear drag strong enough to hit the limit
    "upwelling_limit": (orc.APP_UPWELLING, dict(Lm=20, Mm=24, N=8, limit_bstress=1, rdrg=0.5)),
    "benchmark_nospl": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, nospl_vvisc=1, nospl_vdiff=1)),     # SPLINES_VVISC / SPLINES_VDIFF undefined
    "seamount_nospl": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, nospl_vvisc=1, nospl_vdiff=1)),
    "upwelling_nospl_n30": (orc.APP_UPWELLING, dict(Lm=20, Mm=24, N=30, nospl_vvisc=1, nospl_vdiff=0)),
    "flux_corr": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, qcorrection=1, limit_stflx_cooling=1, scorrection=1, Tnudg_salt=1.0e-6)),
    "flux_relax": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, scorrection=2, Tnudg_salt=2.0e-7)),
    "bodyforce": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, bodyforce=1, levsfrc=5, levbfrc=2)),
    "bodyforce_upwelling_c4": (orc.APP_UPWELLING, dict(Lm=20, Mm=24, N=8, bodyforce=1, levsfrc=8, levbfrc=1, uv_adv=1)),
    "vtransform1": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, Vtransform=1)),
    "vtransform1_seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, Vtransform=1)),
    "atm_press": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, atm_press=1)),                          # ATM_PRESS with prsgrd32
    "atm_press_p31": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, atm_press=1, dj_gradps=3)),          # ... the weighted prsgrd31
    "atm_press_p40": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, atm_press=1, dj_gradps=2)),           # ... prsgrd40
    "benchmark_splines": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, vadv=3)),
    "benchmark_bvf": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, bv_frequency=1, bvf_mixing=1)),
    "benchmark_geo": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, mix_geo_ts=1)),
    "benchmark_a4": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, hadv=1, vadv=1)),
    "benchmark_c2": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, hadv=3, vadv=2)),
    "benchmark_full": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=30, bv_frequency=1, eos_tderivative=1, solar_source=1, lmd_nonlocal=1,
                                               bulk_fluxes=1, lmd_mixing=1, mix_geo_ts=1)),      # the shipped benchmark.h cpp set
    "benchmark_full_n12": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=12, bv_frequency=1, eos_tderivative=1, solar_source=1, lmd_nonlocal=1,
                                                   bulk_fluxes=1, lmd_mixing=1)),
    "uv_c4": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, uv_adv=1)),
    "uv_c4_seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, uv_adv=1)),
    "uv_c2": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, uv_adv=3)),
    "uv_c2_seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, uv_adv=3)),
    "uv_sadv": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, uv_adv=2)),
    "uv_sadv_seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, uv_adv=2)),
    "ts_dif4": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, ts_dif4=1, tnu4=1.0e15)),
    "ts_dif4_upwelling": (orc.APP_UPWELLING, dict(Lm=20, Mm=24, N=8, ts_dif4=1, tnu4=4.0e8)),
    "both_n30": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=30, uv_adv=1, ts_dif4=1, tnu4=1.0e15)),
}


# every case on the AB3 branch (three spin-up steps); the Euler start-up branch (iic == ntfirst) for a core subset
_FIRST_STEP_TOO = ("seamount", "upwelling", "benchmark", "benchmark_full", "benchmark_geo", "uv_c2", "ts_dif4", "benchmark_nospl", "bodyforce")


@pytest.mark.parametrize("case,spinup", [(c, 3) for c in sorted(CASES)] + [(c, 0) for c in _FIRST_STEP_TOO])
def test_kernel_source_bit_exact_against_oracle(case, spinup):
    app, kw = CASES[case]
    o = orc.Oracle(app, **kw)
    o.run_phase("set_data"); o.run_phase("ini")
    if spinup:
        o.step(spinup)
    d = o.indices()
    d["nstp"] = 1 + ((d["iic"] - d["ntstart"]) % 2); d["nnew"] = 3 - d["nstp"]; d["nrhs"] = d["nstp"]
    d["tdays"] = d["time"] / 86400.0
    o.set_indices(d)
    o.run_phase("set_data")
    fill_flux_data(o)                                         # after set_data (which resets the analytical surface fluxes)
    t = EmuTile(o)
    names = all_names(int(o.opt("NT"))) + optional_names(o)
    phases = list(STEP_PHASES)
    if o.opt("bvf_mixing"):
        phases[phases.index("ana_vmix")] = "bvf_mix"
    if o.opt("lmd_mixing"):
        phases[phases.index("ana_vmix")] = "lmd_vmix"
    if o.opt("bulk_fluxes"):
        phases.insert(phases.index("set_vbc"), "bulk_flux")
    if o.opt("ts_dif4"):
        phases.insert(phases.index("t3dmix") + 1, "t3dmix4")
    ran = 0

    AVG = ("Zt_avg1", "DU_avg1", "DV_avg1", "DU_avg2", "DV_avg2")

    def both(ph, interior_avg=False):
        for n in names:                                       # the oracle's state BEFORE the phase: every phase is checked on its own
            t.set(n, o.field(n))
        t.set_indices(o.indices())
        o.run_phase(ph); t.run_phase(ph)
        for n in names:
            a, b = o.field(n), t.get(n)
            if interior_avg and n in AVG:
                # the fast-time averages are only accumulated point by point during the loop: the kernel fills their periodic images in
                # the last call of the loop (where every consumer finds them), the oracle after every call -- compare i = 1..Lm
                a, b = a[:, :, 3:-2], b[:, :, 3:-2]
            if not np.array_equal(a, b):
                dif = np.abs(a - b)
                raise AssertionError(f"{case} spinup={spinup} phase {ph} {o.indices()} field {n}: {np.count_nonzero(dif > 0)} points differ, max {np.nanmax(dif)}")

    for ph in phases:
        if ph == "step2d_loop":
            # LOOP_2D (main3d.F:592-700) call by call through the oracle's step2d; the first five calls and the last three (the
            # final corrector and the averaging-only call iif = nfast + 1) also run in the emulation
            twin = None
            if spinup == 0:
                twin = orc.Oracle(app, **kw); twin.run_phase("set_data"); twin.run_phase("ini"); twin.set_indices(o.indices())
                for n in names:
                    twin.field(n)[:] = o.field(n)
                twin.run_phase("step2d_loop")
            nfast, call, ncall = int(o.opt("nfast")), 0, 2 * int(o.opt("nfast")) + 1
            d = o.indices(); d["PREDICTOR"] = 0

            def sub():
                nonlocal call, ran
                call += 1
                o.set_indices(d)
                if call <= 5 or call >= ncall - 2:
                    both("step2d", interior_avg=(call < ncall)); ran += 1
                else:
                    o.run_phase("step2d")
            for my_iif in range(1, nfast + 2):                # csrc/api.cu loop2d_machine
                nxt = 3 - d["indx1"]
                d["PREDICTOR"] = 1; d["iif"] = my_iif
                d["kstp"] = d["indx1"] if my_iif == 1 else 3 - d["indx1"]; d["knew"] = 3; d["krhs"] = d["indx1"]
                sub()
                d["PREDICTOR"] = 0; d["knew"] = nxt; d["kstp"] = 3 - nxt; d["krhs"] = 3
                if my_iif < nfast + 1:
                    d["indx1"] = nxt
                    sub()
            o.set_indices(d)
            assert call == ncall
            if twin is not None:                              # the Python call sequence above IS the oracle's LOOP_2D
                assert twin.indices() == o.indices()
                for n in names:
                    assert np.array_equal(twin.field(n), o.field(n)), n
        elif ph in EMULATED:
            both(ph); ran += 1
        else:
            o.run_phase(ph)
    assert ran >= 24
    t.close()


@pytest.mark.parametrize("case", ["seamount", "benchmark_geo", "uv_c2"])
def test_kernel_chain_whole_steps_without_resync(case):
    """Two whole baroclinic steps with the state kept in the emulated device arrays (no re-upload between phases, LOOP_2D as
    2 nfast + 1 per-call launches): every ghost row / periodic image / wall value that a later kernel consumes must have been
    written by the kernel that produces it.  Only the forcing (set_data) comes from the oracle, as in the resident C-ABI form."""
    app, kw = CASES[case]
    o = orc.Oracle(app, **kw)
    o.run_phase("set_data"); o.run_phase("ini")
    t = EmuTile(o)
    names = all_names(int(o.opt("NT"))) + optional_names(o)
    for n in names:
        t.set(n, o.field(n))
    nfast = int(o.opt("nfast"))
    for step in range(2):
        d = o.indices()
        d["nstp"] = 1 + ((d["iic"] - d["ntstart"]) % 2); d["nnew"] = 3 - d["nstp"]; d["nrhs"] = d["nstp"]
        d["tdays"] = d["time"] / 86400.0
        o.set_indices(d)
        o.run_phase("set_data")
        t.set("sustr", o.field("sustr")); t.set("svstr", o.field("svstr"))
        for ph in (STEP_PHASES[:STEP_PHASES.index("t3dmix") + 1] + ["t3dmix4"] * int(o.opt("ts_dif4")) + STEP_PHASES[STEP_PHASES.index("t3dmix") + 1:]):
            if ph == "step2d_loop":
                e = o.indices(); e["PREDICTOR"] = 0
                for my_iif in range(1, nfast + 2):            # csrc/api.cu loop2d_machine
                    nxt = 3 - e["indx1"]
                    e["PREDICTOR"] = 1; e["iif"] = my_iif
                    e["kstp"] = e["indx1"] if my_iif == 1 else 3 - e["indx1"]; e["knew"] = 3; e["krhs"] = e["indx1"]
                    t.set_indices(e); t.run_phase("step2d")
                    e["PREDICTOR"] = 0; e["knew"] = nxt; e["kstp"] = 3 - nxt; e["krhs"] = 3
                    if my_iif < nfast + 1:
                        e["indx1"] = nxt
                        t.set_indices(e); t.run_phase("step2d")
                o.run_phase("step2d_loop")
                assert {k: o.indices()[k] for k in ("indx1", "kstp", "krhs", "knew")} == {k: e[k] for k in ("indx1", "kstp", "krhs", "knew")}
            else:
                t.set_indices(o.indices())
                o.run_phase(ph); t.run_phase(ph)
        # advance the oracle's clock exactly as main3d does at the end of a step
        d = o.indices(); d["iic"] += 1; d["time"] += o.opt("dt"); o.set_indices(d)
        for n in names:
            a, b = o.field(n), t.get(n)
            assert np.array_equal(a, b), f"{case} step {step} field {n}: {np.count_nonzero(a != b)} points differ"
    t.close()


@pytest.mark.parametrize("case", ["benchmark", "seamount", "benchmark_full_n12", "uv_c2", "uv_sadv", "ts_dif4", "benchmark_p40", "benchmark_wj",
                                  "benchmark_limit", "benchmark_geo"])
def test_kernel_source_on_ring_tiles(case):
    """The same kernels as xi-TILES of a ring (NtileI = 3, uneven widths; ew_wrap = 0, arrays Istr-3 .. Iend+2 as on the device): every
    tile starts each phase from the oracle's state -- i.e. after a perfect halo exchange -- and must reproduce the oracle's owned
    columns bit for bit.  This is the tiling-invariance of the kernels' index handling (tile offsets, no periodic self-images, the
    eastern-edge copy of lmd_finish) checked without GPUs; the exchange itself is what tests/mgpu_check.py checks on a multi-GPU box."""
    app, kw = CASES[case]
    o = orc.Oracle(app, **kw)
    o.run_phase("set_data"); o.run_phase("ini")
    o.step(2)
    d = o.indices()
    d["nstp"] = 1 + ((d["iic"] - d["ntstart"]) % 2); d["nnew"] = 3 - d["nstp"]; d["nrhs"] = d["nstp"]
    d["tdays"] = d["time"] / 86400.0
    o.set_indices(d)
    o.run_phase("set_data")
    NtileI = 3
    tiles = [EmuTile(o, NtileI=NtileI, tile=q) for q in range(NtileI)]
    own = [(t.LBi + 3, t.LBi + t.ni - 3) for t in tiles]                       # Istr, Iend
    assert own[0][0] == 1 and own[-1][1] == int(o.opt("Lm")) and all(own[q][1] + 1 == own[q + 1][0] for q in range(NtileI - 1))
    names = all_names(int(o.opt("NT"))) + optional_names(o)
    phases = list(STEP_PHASES)
    if o.opt("lmd_mixing"):
        phases[phases.index("ana_vmix")] = "lmd_vmix"
    if o.opt("bulk_fluxes"):
        phases.insert(phases.index("set_vbc"), "bulk_flux")
    if o.opt("ts_dif4"):
        phases.insert(phases.index("t3dmix") + 1, "t3dmix4")

    def both(oo, ph):
        for t in tiles:
            for n in names:
                t.set(n, oo.field(n))
            t.set_indices(oo.indices())
        oo.run_phase(ph)
        for q, t in enumerate(tiles):
            t.run_phase(ph)
            Istr, Iend = own[q]
            for n in names:
                a = oo.field(n)[:, :, Istr + 2:Iend + 3]
                b = t.get(n)[:, :, Istr - t.LBi:Iend - t.LBi + 1]
                if not np.array_equal(a, b):
                    raise AssertionError(f"{case} tile {q} phase {ph} {oo.indices()} field {n}: {np.count_nonzero(a != b)} owned points differ")

    # a twin of the oracle, advanced to the start of LOOP_2D, plays the first four step2d calls with the tiles; the main oracle runs
    # the whole loop on its own and then the rest of the step with the tiles
    twin = orc.Oracle(app, **kw)
    twin.run_phase("set_data"); twin.run_phase("ini"); twin.step(2)
    twin.set_indices(d); twin.run_phase("set_data")
    for ph in phases[:phases.index("step2d_loop")]:
        twin.run_phase(ph)
    e = twin.indices(); e["PREDICTOR"] = 0
    for my_iif in (1, 2):
        nxt = 3 - e["indx1"]
        e["PREDICTOR"] = 1; e["iif"] = my_iif
        e["kstp"] = e["indx1"] if my_iif == 1 else 3 - e["indx1"]; e["knew"] = 3; e["krhs"] = e["indx1"]
        twin.set_indices(e); both(twin, "step2d")
        e["PREDICTOR"] = 0; e["knew"] = nxt; e["kstp"] = 3 - nxt; e["krhs"] = 3; e["indx1"] = nxt
        twin.set_indices(e); both(twin, "step2d")
    ran = 0
    for ph in phases:
        if ph in EMULATED:
            both(o, ph); ran += 1
        else:
            o.run_phase(ph)                                                    # step2d_loop
    assert ran >= 16
    for t in tiles:
        t.close()


class _EmuAsTile:
    """Adapter that lets the body of a GPU parity test run against the emulation: the subset of ocean.Tile the every-phase tests use
    (set / get / set_indices / indices / run_phase incl. the whole LOOP_2D / close)."""

    def __init__(self, o):
        self.o, self.e, self.idx = o, EmuTile(o), dict(o.indices())

    def set(self, name, arr):
        self.e.set(name, arr)

    def get(self, name):
        return self.e.get(name)

    def set_indices(self, d):
        self.idx = dict(d); self.e.set_indices(self.idx)

    def indices(self):
        return dict(self.idx)

    def run_phase(self, ph):
        if ph != "step2d_loop":
            self.e.set_indices(self.idx); self.e.run_phase(ph)
            return
        e, nfast = self.idx, self.e.nfast
        e["PREDICTOR"] = 0
        for my_iif in range(1, nfast + 2):                        # csrc/api.cu loop2d_machine
            nxt = 3 - e["indx1"]
            e["PREDICTOR"] = 1; e["iif"] = my_iif
            e["kstp"] = e["indx1"] if my_iif == 1 else 3 - e["indx1"]; e["knew"] = 3; e["krhs"] = e["indx1"]
            self.e.set_indices(e); self.e.run_phase("step2d")
            e["PREDICTOR"] = 0; e["knew"] = nxt; e["kstp"] = 3 - nxt; e["krhs"] = 3
            if my_iif < nfast + 1:
                e["indx1"] = nxt
                self.e.set_indices(e); self.e.run_phase("step2d")

    def close(self):
        self.e.close()


@pytest.mark.parametrize("case", ["flux_corr", "flux_relax", "bodyforce", "limit_bstress", "nospl", "uv_c2"])
def test_gpu_variant_test_bodies_dry_run_on_the_emulation(case, monkeypatch):
    """The BODY of tests/test_zz_variants_gpu.py::test_variants_strict_bit_exact_every_phase executed with the emulation standing in
    for the device tile (same make_pair / begin_step / fill / compare sequence, small grids): checks the test's own host logic for
    the cases that had no GPU run when the round closed, and once more that the kernel chain keeps its state consistent phase after
    phase WITHOUT re-upload."""
    import helpers
    import test_zz_variants_gpu as G
    small = {"flux_corr": dict(Lm=24, Mm=16, N=7), "flux_relax": dict(Lm=24, Mm=16, N=7), "bodyforce": dict(Lm=24, Mm=16, N=7, levsfrc=5),
             "limit_bstress": dict(Lm=24, Mm=16, N=7), "nospl": dict(Lm=24, Mm=16, N=7), "uv_c2": dict(Lm=24, Mm=16, N=7)}[case]
    app, kw = G.VARIANTS[case]
    monkeypatch.setitem(G.VARIANTS, case, (app, dict(kw, **small)))

    def fake_make_pair(app, strict=True, spinup=0, **kw):
        o = orc.Oracle(app, **kw)
        o.run_phase("set_data"); o.run_phase("ini")
        if spinup:
            o.step(spinup)
        t = _EmuAsTile(o)
        NT = int(o.opt("NT"))
        for n in all_names(NT) + optional_names(o):
            t.set(n, o.field(n))
        t.set_indices(o.indices())
        return o, t
    monkeypatch.setattr(G, "make_pair", fake_make_pair)
    G.test_variants_strict_bit_exact_every_phase(case, 3)


@pytest.mark.parametrize("case", ["seamount", "benchmark", "upwelling"])
def test_diag_kernels_bit_exact(case):
    """diag.F through the three device kernels (column sums, row sums in the reference's order, the final CTA whose warp 1 reduces the
    maxima with shuffles -- emulated with a mailbox per warp) and the host-side finish of roms_b200_diag: the 12 scalars (KE, PE,
    volume, maximum speed, the Courant numbers AT the location of the largest one, ...) equal the oracle's exactly."""
    app, kw = CASES[case]
    o = orc.Oracle(app, **kw)
    o.run_phase("set_data"); o.run_phase("ini")
    o.step(5)
    t = EmuTile(o)
    for n in all_names(int(o.opt("NT"))) + optional_names(o):
        t.set(n, o.field(n))
    t.set_indices(o.indices())
    want, got = o.diag(), t.diag()
    assert want["max_speed"] > 0 and want["avgke"] > 0
    for k in want:
        assert want[k] == got[k], (k, want[k], got[k])
    t.close()


def test_emulation_negative_controls():
    """The comparison has teeth: a relative change of 1e-7 in ONE diff4 value, or the emulated kernel running the default advection
    while the oracle runs UV_C4ADVECTION, makes the phase differ."""
    o = orc.Oracle(orc.APP_BENCHMARK, Lm=37, Mm=19, N=7, uv_adv=1, ts_dif4=1, tnu4=1.0e15)
    o.run_phase("set_data"); o.run_phase("ini"); o.step(3)
    d = o.indices(); d["nstp"] = 1 + ((d["iic"] - d["ntstart"]) % 2); d["nnew"] = 3 - d["nstp"]; d["nrhs"] = d["nstp"]; o.set_indices(d)
    names = all_names(2) + optional_names(o)
    for ph in STEP_PHASES[:STEP_PHASES.index("t3dmix") + 1]:
        o.run_phase(ph)
    t = EmuTile(o)
    for n in names:
        t.set(n, o.field(n))
    t.set_indices(o.indices())
    a = o.field("diff4_0").copy(); a[0, 5, 7] *= 1.0000001; t.set("diff4_0", a)
    o.run_phase("t3dmix4"); t.run_phase("t3dmix4")
    new = "t%d_" % d["nnew"]
    assert not np.array_equal(o.field(new + "0"), t.get(new + "0"))                 # the perturbed tracer differs ...
    assert np.array_equal(o.field(new + "1"), t.get(new + "1"))                     # ... the other one does not
    t.close()

    class WrongScheme:                                                              # an oracle front that reports uv_adv = 0
        def __init__(self, o): self.o = o
        def opt(self, k): return 0 if k == "uv_adv" else self.o.opt(k)
        def vector(self, *a): return self.o.vector(*a)
    t = EmuTile(WrongScheme(o))
    for n in names:
        t.set(n, o.field(n))
    t.set_indices(o.indices())
    o.run_phase("rhs3d"); t.run_phase("rhs3d")
    assert not np.array_equal(o.field("rv%d" % d["nrhs"]), t.get("rv%d" % d["nrhs"]))
    t.close()
