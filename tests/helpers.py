"""Shared helpers for the parity tests: build an oracle model and a device Tile holding the same state."""
import numpy as np

import orc
from roms_trunk_mgh_b200 import _lib
from roms_trunk_mgh_b200.ocean import Tile, field_names

APP_OPTS = ["nonlin_eos", "dj_gradps", "curvgrid", "mix_geo_ts", "uv_qdrag", "hadv", "vadv", "ana_vmix", "wvelocity_every_step",
            "bv_frequency", "eos_tderivative", "solar_source", "lmd_nonlocal", "bulk_fluxes", "lmd_mixing", "bvf_mixing", "uv_adv", "ts_dif4", "limit_bstress", "nospl_vvisc", "nospl_vdiff", "qcorrection", "limit_stflx_cooling", "scorrection", "bodyforce", "levsfrc", "levbfrc", "atm_press"]
# (Vtransform is copied separately: the oracle spells it with a capital)


def cfg_from_oracle(o, device=0):
    """roms_b200_config carrying exactly the oracle's switches and parameters."""
    app = int(o.opt("app"))
    cfg = _lib.default_config(app, int(o.opt("Lm")), int(o.opt("Mm")), int(o.opt("N")))
    cfg.NT = int(o.opt("NT"))
    for k in APP_OPTS:
        setattr(cfg, k, int(o.opt(k)))
    cfg.salinity = int(o.opt("salinity"))
    cfg.dt = o.opt("dt"); cfg.ndtfast = int(o.opt("ndtfast"))
    cfg.rho0 = o.opt("rho0"); cfg.g = o.opt("g")
    cfg.R0 = o.opt("R0"); cfg.T0 = o.opt("T0"); cfg.S0 = o.opt("S0"); cfg.Tcoef = o.opt("Tcoef"); cfg.Scoef = o.opt("Scoef")
    cfg.Akt_bak[0] = cfg.Akt_bak[1] = o.opt("Akt_bak"); cfg.Akv_bak = o.opt("Akv_bak")
    cfg.gamma2 = o.opt("gamma2"); cfg.lambda_ = o.opt("lambda"); cfg.hc = o.opt("hc")
    cfg.Tnudg_salt = o.opt("Tnudg_salt")
    cfg.vtransform = int(o.opt("Vtransform"))
    cfg.blk_ZQ = o.opt("blk_ZQ"); cfg.blk_ZT = o.opt("blk_ZT"); cfg.blk_ZW = o.opt("blk_ZW")
    cfg.device = device
    return cfg


ATMOSPHERE = ["Uwind", "Vwind", "Tair", "Pair", "Hair", "rain", "cloud", "srflx"]      # what the host hands bulk_flux (FORCES)


def optional_names(o):
    """Names of the arrays that exist only when a cpp switch of the BENCHMARK set is on (include/roms_b200.h)."""
    NT = int(o.opt("NT"))
    v = []
    if o.opt("uv_qdrag") == 2: v += ["ZoBot"]
    if o.opt("bv_frequency"): v += ["bvf"]
    if o.opt("eos_tderivative"): v += ["alpha", "beta"]
    if o.opt("solar_source"): v += ["srflx", "Jwtype"]
    if o.opt("lmd_nonlocal"): v += [f"ghats_{it}" for it in range(NT)]
    if o.opt("bulk_fluxes"): v += [n for n in ATMOSPHERE if n not in v] + ["lrflx", "lhflx", "shflx"]
    if o.opt("lmd_mixing"): v += ["hsbl", "ksbl"]
    if o.opt("ts_dif4"): v += [f"diff4_{it}" for it in range(NT)]
    if o.opt("atm_press") and "Pair" not in v: v += ["Pair"]
    if o.opt("qcorrection"): v += ["sst", "dqdt"]
    if o.opt("scorrection"): v += ["sss"]
    return v


def copy_state(o, t):
    """Upload every field, vector and index of oracle `o` into device tile `t`."""
    NT = int(o.opt("NT")); N = int(o.opt("N"))
    n2, n3 = field_names(NT)
    for n in n2 + n3 + optional_names(o):
        t.set(n, o.field(n))
    t.set_scoord(o.vector(0, N + 1), o.vector(1, N + 1), o.vector(2, N + 1), o.vector(3, N + 1))
    nd = int(o.opt("ndtfast"))
    t.set_weights(int(o.opt("nfast")), o.vector(4, 2 * nd + 2), o.vector(5, 2 * nd + 2))
    t.set_indices(o.indices())


def make_pair(app, strict=True, spinup=0, **kw):
    o = orc.Oracle(app, **kw)
    o.run_phase("set_data")
    o.run_phase("ini")
    if spinup:
        o.step(spinup)
    t = Tile(cfg_from_oracle(o), strict=strict)
    copy_state(o, t)
    return o, t


def compare(o, t, names, exact=True, rtol=0.0, label=""):
    """Compare fields; returns list of (name, max_abs_diff, max_abs_ref) for mismatching fields."""
    bad = []
    for n in names:
        a = o.field(n); b = t.get(n)
        if exact:
            if not np.array_equal(a, b):
                d = np.abs(a - b)
                bad.append((n, float(np.nanmax(d)), float(np.nanmax(np.abs(a))), int(np.count_nonzero(d > 0))))
        else:
            scale = float(np.max(np.abs(a)))
            d = float(np.max(np.abs(a - b)))
            if not np.isfinite(d) or d > rtol * max(scale, 1e-300):
                bad.append((n, d, scale, -1))
    return bad


def all_names(NT):
    n2, n3 = field_names(NT)
    return n2 + n3


def fill_flux_data(o):
    if o.opt("atm_press") and not o.opt("bulk_fluxes"):      # ATM_PRESS: a pressure field (mb) with gradients in both directions
        a = o.field("Pair"); j = np.arange(a.shape[1])[None, :, None]; i = np.arange(a.shape[2])[None, None, :]
        a[:] = 1013.25 + 6.0 * np.sin(2.0 * np.pi * (i - 2) / (a.shape[2] - 5)) + 0.2 * j
    """sst, dqdt, sss of set_vbc's QCORRECTION / SCORRECTION / SRELAXATION (host data in a real run): analytical, part below -2 degC."""
    if o.opt("qcorrection"):
        a = o.field("sst"); j = np.arange(a.shape[1])[None, :, None]; i = np.arange(a.shape[2])[None, None, :]
        a[:] = 1.0 + 0.1 * j + 0.01 * i
        o.field("dqdt")[:] = -40.0 / (1025.0 * 3985.0) * (1.0 + 0.02 * j)
    if o.opt("scorrection"):
        a = o.field("sss"); j = np.arange(a.shape[1])[None, :, None]; i = np.arange(a.shape[2])[None, None, :]
        a[:] = 34.0 + 0.03 * j + 0.002 * i
    if o.opt("limit_stflx_cooling"):
        o.field("stflux_0")[:] = -1.0e-5
        o.field("stflux_0")[0, ::3, :] = 1.0e-5
        for tl in (1, 2):
            o.field(f"t{tl}_0")[-1, :, ::2] = -2.5
