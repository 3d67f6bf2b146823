"""CPU check of the CUDA kernels' SOURCE (no GPU needed): tests/emu compiles the unmodified text of the barrier-free kernels
(csrc/k_glue.cu, k_pre.cu, k_rhs.cu: set_massflux, rho_eos, set_vbc, ana_vmix, omega, wvelocity, set_zeta, pre_step3d, prsgrd31/32,
t3dmix2_s, t3dmix4_s, rhs3d, uv3dmix2, set_depth, bvf_mix) for the host and runs every thread of every launch in turn.  Built like
the oracle's parity build (-O2 -ffp-contract=off), each phase must reproduce the oracle BIT FOR BIT from the oracle's own inputs:
loop ranges, wall / periodic-image handling, upstream selects and operation order of the kernel text are pinned without a device.
What it cannot see: kernels with shared memory / barriers (step2d, step3d_uv, step3d_t, t3dmix2_geo, diag, lmd_vmix's east column),
device libm, races.  The -m gpu tests remain the parity tests proper."""
import os
import sys

import numpy as np
import pytest

import orc
from helpers import all_names, optional_names

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from emu import EMULATED, EmuTile  # noqa: E402

STEP_PHASES = ["set_massflux", "rho_eos", "set_vbc", "ana_vmix", "omega", "wvelocity", "set_zeta", "pre_step3d", "prsgrd", "t3dmix",
               "rhs3d", "uv3dmix", "step2d_loop", "set_depth", "step3d_uv", "omega2", "step3d_t"]
CASES = {
    "seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8)),
    "upwelling": (orc.APP_UPWELLING, dict(Lm=20, Mm=24, N=8)),
    "benchmark": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7)),
    "benchmark_p31": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, dj_gradps=0, nonlin_eos=0)),
    "benchmark_splines": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, vadv=3)),
    "benchmark_bvf": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=8, bv_frequency=1, bvf_mixing=1)),
    "uv_c4": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, uv_adv=1)),
    "uv_c4_seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, uv_adv=1)),
    "uv_sadv": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, uv_adv=2)),
    "uv_sadv_seamount": (orc.APP_SEAMOUNT, dict(Lm=24, Mm=20, N=8, uv_adv=2)),
    "ts_dif4": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, ts_dif4=1, tnu4=1.0e15)),
    "ts_dif4_upwelling": (orc.APP_UPWELLING, dict(Lm=20, Mm=24, N=8, ts_dif4=1, tnu4=4.0e8)),
    "both_n30": (orc.APP_BENCHMARK, dict(Lm=32, Mm=16, N=30, uv_adv=1, ts_dif4=1, tnu4=1.0e15)),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("spinup", [0, 3])
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
    t = EmuTile(o)
    names = all_names(int(o.opt("NT"))) + [n for n in optional_names(o) if n not in ("ksbl",)]
    phases = list(STEP_PHASES)
    if o.opt("bvf_mixing"):
        phases[phases.index("ana_vmix")] = "bvf_mix"
    ran = 0
    for ph in phases:
        if ph in EMULATED and not (ph == "t3dmix" and o.opt("mix_geo_ts")):
            for n in names:                                   # the oracle's state BEFORE the phase: every phase is checked on its own
                t.set(n, o.field(n))
            t.set_indices(o.indices())
            o.run_phase(ph); t.run_phase(ph)
            for n in names:
                a, b = o.field(n), t.get(n)
                if not np.array_equal(a, b):
                    dif = np.abs(a - b)
                    raise AssertionError(f"{case} spinup={spinup} phase {ph} field {n}: {np.count_nonzero(dif > 0)} points differ, max {np.nanmax(dif)}")
            ran += 1
        else:
            o.run_phase(ph)
    assert ran >= 13
    t.close()
