"""GPU parity tests (pytest -m gpu) of the SURVEY 8(f)-3 variants added last in round 2: UV_C4ADVECTION (rhs3d.F:685-921,
:1108-1175, :1362-1429) and TS_DIF4 + MIX_S_TS (t3dmix4_s.h:215-476).  Same bars as tests/test_gpu_parity.py: the strict
(-fmad=false) library is BIT-EXACT against the oracle after every phase and over several steps; the production library is held to
1e-8 (zeta, u, v) / 1e-12 (tracers).  The file sorts after the rest of the suite on purpose: these kernels were written when the
round's GPU budget was spent, so their first run on a B200 is the driver's."""
import ctypes as C

import numpy as np
import pytest

import orc
from helpers import all_names, cfg_from_oracle, compare, fill_flux_data, make_pair, optional_names
from roms_trunk_mgh_b200 import _lib
from test_gpu_parity import STEP_PHASES, begin_step, test_tiling_invariance_across_gpus as _tiling_invariance

pytestmark = pytest.mark.gpu

TNU4 = 1.0e15      # m4/s on the 64 x 32 channel (dy = 69 km): dt * tnu4 * 16 / dy^4 = 0.1, a visible and stable biharmonic term
VARIANTS = {
    "uv_c4": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, uv_adv=1)),
    "uv_c4_ragged": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, uv_adv=1)),
    "uv_c4_seamount": (orc.APP_SEAMOUNT, dict(uv_adv=1)),                                   # no-slip walls, NT = 1
    "uv_c2": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, uv_adv=3)),                        # UV_C2ADVECTION: rhs3d AND step2d (per-call kernels)
    "uv_c2_seamount": (orc.APP_SEAMOUNT, dict(uv_adv=3)),
    "p40": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, dj_gradps=2)),                         # PJ_GRADP: prsgrd40.h
    "p40_seamount": (orc.APP_SEAMOUNT, dict(dj_gradps=2)),
    "wj": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, dj_gradps=3)),                          # WJ_GRADP: prsgrd31.h, weighted Jacobian
    "wj_seamount": (orc.APP_SEAMOUNT, dict(dj_gradps=3)),
    "limit_bstress": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, limit_bstress=1, uv_qdrag=0, rdrg=5.0)),   # LIMIT_BSTRESS with a linear drag that hits the limit
    "nospl": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, nospl_vvisc=1, nospl_vdiff=1)),       # SPLINES_VVISC / SPLINES_VDIFF undefined
    "nospl_seamount": (orc.APP_SEAMOUNT, dict(nospl_vvisc=1, nospl_vdiff=1)),
    "nospl_n30": (orc.APP_BENCHMARK, dict(Lm=96, Mm=40, N=30, nospl_vvisc=0, nospl_vdiff=1, vadv=3)),
    "flux_corr": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, qcorrection=1, limit_stflx_cooling=1, scorrection=1, Tnudg_salt=1.0e-6)),
    "flux_relax": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, scorrection=2, Tnudg_salt=2.0e-7)),
    "bodyforce": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, bodyforce=1, levsfrc=8, levbfrc=2)),
    "atm_press": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, atm_press=1)),                    # ATM_PRESS in prsgrd32
    "atm_press_p40": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, atm_press=1, dj_gradps=2)),
    "vtransform1": (orc.APP_SEAMOUNT, dict(Vtransform=1)),                                   # the original vertical transformation in set_depth
    "uv_sadv": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, uv_adv=2)),                      # UV_SADVECTION: spline vertical advection
    "uv_sadv_seamount": (orc.APP_SEAMOUNT, dict(uv_adv=2)),
    "ts_dif4": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, ts_dif4=1, tnu4=TNU4)),
    "ts_dif4_ragged": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7, ts_dif4=1, tnu4=TNU4)),
    "both_n30": (orc.APP_BENCHMARK, dict(Lm=96, Mm=40, N=30, uv_adv=1, ts_dif4=1, tnu4=TNU4)),
}


@pytest.mark.parametrize("case", sorted(VARIANTS))
@pytest.mark.parametrize("spinup", [0, 3])
def test_variants_strict_bit_exact_every_phase(case, spinup):
    app, kw = VARIANTS[case]
    o, t = make_pair(app, strict=True, spinup=spinup, **kw)
    names = all_names(int(o.opt("NT")))
    begin_step(o, t)
    if o.opt("qcorrection") or o.opt("scorrection") or o.opt("atm_press"):   # host data fields of these options,
        fill_flux_data(o)                                       # set after set_data, which resets the analytical surface fluxes
        for n in ("sst", "dqdt", "sss", "Pair", "stflux_0", "t1_0", "t2_0"):
            if n in ("stflux_0", "t1_0", "t2_0") or n in optional_names(o):
                t.set(n, o.field(n))
    k = STEP_PHASES.index("t3dmix") + 1
    for ph in STEP_PHASES[:k] + ["t3dmix4"] * int(o.opt("ts_dif4")) + STEP_PHASES[k:]:       # rhs3d.F:81-97
        o.run_phase(ph); t.run_phase(ph)
        bad = compare(o, t, names, exact=True)
        assert not bad, f"{case} spinup={spinup} phase {ph}: {bad}"
    t.close()


@pytest.mark.parametrize("case", ["uv_c4", "uv_c2", "uv_sadv", "p40", "wj", "ts_dif4", "both_n30", "nospl", "nospl_seamount", "bodyforce", "vtransform1",
                                  "limit_bstress"])
def test_variants_strict_bit_exact_multistep(case):
    """Whole steps (the captured step graph; t3dmix2_s fused into pre_step3d_t with t3dmix4_s behind it)."""
    app, kw = VARIANTS[case]
    o, t = make_pair(app, strict=True, **kw)
    for _ in range(8):
        o.step(1); t.main3d(1)
    assert not compare(o, t, all_names(int(o.opt("NT"))), exact=True)
    do, dt_ = o.diag(), t.diag()
    for k in do:
        assert do[k] == dt_[k], (k, do[k], dt_[k])
    t.close()


def test_variants_change_the_answer():
    """The switches are live on the device: the state after 6 steps differs from the default branches' and equals the oracle's."""
    base = dict(Lm=64, Mm=32, N=10)
    o0, t0 = make_pair(orc.APP_BENCHMARK, strict=True, **base)
    o1, t1 = make_pair(orc.APP_BENCHMARK, strict=True, uv_adv=1, **base)
    o2, t2 = make_pair(orc.APP_BENCHMARK, strict=True, ts_dif4=1, tnu4=TNU4, **base)
    for t in (t0, t1, t2):
        t.main3d(6)
    assert not np.array_equal(t0.get("u1"), t1.get("u1")) and not np.array_equal(t0.get("u2"), t1.get("u2"))
    assert not np.array_equal(t0.get("t1_0"), t2.get("t1_0"))
    for t in (t0, t1, t2):
        t.close()


def test_variants_production_tolerance():
    app, kw = VARIANTS["both_n30"]
    o, t = make_pair(app, strict=False, **kw)
    for _ in range(20):
        o.step(1); t.main3d(1)
    bad = compare(o, t, ["zeta1", "zeta2", "u1", "u2", "v1", "v2", "W"], exact=False, rtol=1e-8)
    bad += compare(o, t, ["t1_0", "t2_0", "t1_1", "t2_1", "rho"], exact=False, rtol=1e-12)
    assert not bad, bad
    t.close()


def test_ts_dif4_by_routine_and_config_errors():
    """t3dmix4 through roms_b200_routine_tile (its own routine, as in the reference: rhs3d.F:89-97) with the diff4 arrays passed by name; TS_DIF4 with MIX_GEO_TS (t3dmix4_geo.h is
    not built) and an unknown uv_adv are configuration errors (exit_flag 5)."""
    app, kw = VARIANTS["ts_dif4"]
    o, t = make_pair(app, strict=True, spinup=3, **kw)
    t.close()
    L = _lib.load(True)
    cfg = cfg_from_oracle(o)
    NT, N, nd = int(o.opt("NT")), int(o.opt("N")), int(o.opt("ndtfast"))
    d = o.indices()
    d["nstp"] = 1 + ((d["iic"] - d["ntstart"]) % 2); d["nnew"] = 3 - d["nstp"]; d["nrhs"] = d["nstp"]
    d["tdays"] = d["time"] / 86400.0
    o.set_indices(d)
    o.run_phase("set_data")
    for ph in ("set_massflux", "rho_eos", "set_vbc", "ana_vmix", "omega", "wvelocity", "set_zeta", "pre_step3d", "prsgrd", "t3dmix"):
        o.run_phase(ph)
    spec = L.roms_b200_routine_args(_lib.PHASES["t3dmix4"]).decode()
    assert "diff4_*" in spec

    def expand(part):
        out = []
        for n in part.split(":")[1].split(","):
            n = n.lstrip("?")
            out += [n.replace("*", str(it)) for it in range(NT)] if "*" in n else [n]
        return out
    ins, outs = [expand(x) for x in spec.split(";")]
    names = list(dict.fromkeys(ins + outs))
    arrs = [o.field(n).copy() for n in names]
    mode = [(1 if n in ins else 0) | (2 if n in outs else 0) for n in names]
    ta = _lib.TileArgs(cfg=cfg, iic=d["iic"], ntfirst=d["ntfirst"], nstp=d["nstp"], nnew=d["nnew"], nrhs=d["nrhs"], iif=d["iif"],
                       kstp=d["kstp"], krhs=d["krhs"], knew=d["knew"], predictor=d["PREDICTOR"])
    sc = np.concatenate([o.vector(w, N + 1) for w in range(4)])
    w1, w2 = o.vector(4, 2 * nd + 2), o.vector(5, 2 * nd + 2)
    cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
    ca = (_lib.DP * len(names))(*[a.ctypes.data_as(_lib.DP) for a in arrs])
    cm = (C.c_int * len(names))(*mode)
    rc = L.roms_b200_routine_tile(C.byref(ta), _lib.PHASES["t3dmix4"], len(names), cn, ca, cm, sc.ctypes.data_as(_lib.DP), int(o.opt("nfast")),
                                  w1.ctypes.data_as(_lib.DP), w2.ctypes.data_as(_lib.DP), len(w1))
    assert rc == 0
    o.run_phase("t3dmix4")
    for n, a in zip(names, arrs):
        if n in outs:
            assert np.array_equal(a, o.field(n)), n
    bad = cfg_from_oracle(o); bad.mix_geo_ts = 1
    h = C.c_void_p()
    assert L.roms_b200_create(C.byref(bad), C.byref(h)) == 5
    bad = cfg_from_oracle(o); bad.uv_adv = 7
    assert L.roms_b200_create(C.byref(bad), C.byref(h)) == 5


@pytest.mark.parametrize("mode", ["2 uv_adv=3", "1 uv_adv=3 step2d_loop_kernel=0", "0 uv_adv=3 peer=0", "2 uv_adv=1 dj_gradps=2",
                                  "2 uv_adv=2 dj_gradps=3 ts_dif4=1 tnu4=2e13"])
def test_variants_tiling_invariance_across_gpus(mode):
    """The variants on a ring of GPUs (skipped on a one-GPU box): bitwise agreement with the single-tile run, as
    test_gpu_parity.py::test_tiling_invariance_across_gpus demands of the default branches.  uv_adv=3 runs LOOP_2D through the
    per-call kernels k_step2d<XCH, C2> with the exchange fused (modes 2, 1) or stand-alone over NCCL (mode 0, peer=0)."""
    _tiling_invariance(mode)
