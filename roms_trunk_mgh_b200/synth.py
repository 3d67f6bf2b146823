"""Synthetic (analytical) inputs for the three supported applications -- the host-side set-up a ROMS run performs before
main3d: s-coordinate (ROMS/Utility/set_scoord.F:393-440), barotropic filter weights (ROMS/Utility/set_weights.F), analytical
grid (ROMS/Functionals/ana_grid.h), metrics (ROMS/Utility/metrics.F:355-534), uniform mixing coefficients
(ROMS/Utility/ini_hmixcoef.F:257-290), initial tracers (ROMS/Functionals/ana_initial.h) and surface stress
(ROMS/Functionals/ana_smflux.h).  Pure numpy; used by bench.py / smoke() / examples to feed the device library.  There are
no files and no RNG: the reference's idealised cases are analytical, so "seeded" means "same formulas".

All 2-D arrays are GLOBAL, shape (Mm+2, Lm+5) == A(-2:Lm+2, 0:Mm+1); `tile_slice` cuts the columns of one xi-tile.
"""
import math

import numpy as np

from . import _lib

APP_UPWELLING, APP_SEAMOUNT, APP_BENCHMARK = 0, 1, 2
PI = 3.14159265358979323846
DEG2RAD = PI / 180.0
ERADIUS = 6371315.0
# roms_<app>.in: THETA_S, THETA_B, TCLINE, TNU2, VISC2, RDRG, RDRG2
APP_PARAMS = {
    APP_UPWELLING: dict(theta_s=3.0, theta_b=0.0, Tcline=25.0, tnu2=0.0, visc2=5.0, rdrg=3.0e-4, rdrg2=3.0e-3),
    APP_SEAMOUNT: dict(theta_s=6.5, theta_b=2.0, Tcline=100.0, tnu2=0.0, visc2=0.0, rdrg=3.0e-4, rdrg2=3.0e-3),
    APP_BENCHMARK: dict(theta_s=0.0, theta_b=0.0, Tcline=400.0, tnu2=500.0, visc2=5000.0, rdrg=3.0e-4, rdrg2=3.0e-3),
}


def set_scoord(N, theta_s, theta_b):
    """Vtransform=2 / Vstretching=4.  Returns sc_r, Cs_r, sc_w, Cs_w indexed by k (N+1 entries; slot 0 of *_r unused)."""
    ds = 1.0 / N

    def C(s):
        Csur = (1.0 - math.cosh(theta_s * s)) / (math.cosh(theta_s) - 1.0) if theta_s > 0.0 else -(s * s)
        return (math.exp(theta_b * Csur) - 1.0) / (1.0 - math.exp(-theta_b)) if theta_b > 0.0 else Csur

    sc_w = np.zeros(N + 1); Cs_w = np.zeros(N + 1); sc_r = np.zeros(N + 1); Cs_r = np.zeros(N + 1)
    for k in range(N - 1, 0, -1):
        sc_w[k] = ds * (k - N); Cs_w[k] = C(sc_w[k])
    sc_w[0] = -1.0; Cs_w[0] = -1.0
    for k in range(1, N + 1):
        sc_r[k] = ds * ((k - N) - 0.5); Cs_r[k] = C(sc_r[k])
    return sc_r, Cs_r, sc_w, Cs_w


def set_weights(ndtfast):
    """POWER_LAW filter (Falpha=2, Fbeta=4, Fgamma=0.284).  Returns nfast, weight(1,:), weight(2,:) (2*ndtfast+2, slot 0 unused)."""
    Falpha, Fbeta, Fgamma = 2.0, 4.0, 0.284
    n = 2 * ndtfast
    w1 = [0.0] * (n + 2); w2 = [0.0] * (n + 2)
    scale = (Falpha + 1.0) * (Falpha + Fbeta + 1.0) / ((Falpha + 2.0) * (Falpha + Fbeta + 2.0) * float(ndtfast))
    gamma = Fgamma * max(0.0, 1.0 - 10.0 / float(ndtfast))
    nfast = 0
    for _ in range(16):
        nfast = 0
        for i in range(1, n + 1):
            cff = scale * float(i)
            w1[i] = cff ** Falpha - cff ** (Falpha + Fbeta) - gamma * cff
            if w1[i] > 0.0:
                nfast = i
            if nfast > 0 and w1[i] < 0.0:
                w1[i] = 0.0
        wsum = 0.0; shift = 0.0
        for i in range(1, nfast + 1):
            wsum += w1[i]; shift += w1[i] * float(i)
        scale = scale * shift / (wsum * float(ndtfast))
    for _ in range(ndtfast):
        wsum = 0.0; shift = 0.0
        for i in range(1, nfast + 1):
            wsum += w1[i]; shift += float(i) * w1[i]
        shift = shift / wsum
        cff = float(ndtfast) - shift
        if cff > 1.0:
            nfast += 1
            for i in range(nfast, 1, -1):
                w1[i] = w1[i - 1]
            w1[1] = 0.0
        elif cff > 0.0:
            wsum = 1.0 - cff
            for i in range(nfast, 1, -1):
                w1[i] = wsum * w1[i] + cff * w1[i - 1]
            w1[1] = wsum * w1[1]
        elif cff < -1.0:
            nfast -= 1
            for i in range(1, nfast + 1):
                w1[i] = w1[i + 1]
            w1[nfast + 1] = 0.0
        elif cff < 0.0:
            wsum = 1.0 + cff
            for i in range(1, nfast):
                w1[i] = wsum * w1[i] - cff * w1[i + 1]
            w1[nfast] = wsum * w1[nfast]
    for j in range(1, nfast + 1):
        cff = w1[j]
        for i in range(1, j + 1):
            w2[i] += cff
    wsum = sum(w1[1:nfast + 1]); cff = sum(w2[1:nfast + 1])
    wsum = 1.0 / wsum; cff = 1.0 / cff
    for i in range(1, nfast + 1):
        w1[i] *= wsum; w2[i] *= cff
    return nfast, np.array(w1), np.array(w2)


class Grid:
    """Index helper for global arrays A(-2:Lm+2, 0:Mm+1)."""

    def __init__(self, Lm, Mm):
        self.Lm, self.Mm = Lm, Mm
        self.LBi, self.UBi, self.LBj, self.UBj = -2, Lm + 2, 0, Mm + 1
        self.ni, self.nj = Lm + 5, Mm + 2
        self.I = np.arange(self.LBi, self.UBi + 1, dtype=np.float64)[None, :]
        self.J = np.arange(self.LBj, self.UBj + 1, dtype=np.float64)[:, None]

    def zeros(self, nk=None):
        return np.zeros((self.nj, self.ni)) if nk is None else np.zeros((nk, self.nj, self.ni))

    def ci(self, i):
        return i - self.LBi

    def exchange(self, a):
        """Periodic images in xi (ROMS/Nonlinear/exchange_2d.F): A(Lm+1:Lm+2)=A(1:2), A(-2:0)=A(Lm-2:Lm)."""
        c, Lm = self.ci, self.Lm
        a[..., c(Lm + 1):c(Lm + 2) + 1] = a[..., c(1):c(2) + 1]
        a[..., c(-2):c(0) + 1] = a[..., c(Lm - 2):c(Lm) + 1]
        return a


def build(app, Lm=0, Mm=0, N=0, **overrides):
    """Returns (cfg, fields, scoord, (nfast, w1, w2)) for a fresh run at rest (time level 1; ini_fields already applied)."""
    cfg = _lib.default_config(app, Lm, Mm, N)
    tnu4 = float(overrides.pop("tnu4", 0.0))                  # roms_*.in TNU4 (m4/s): only with ts_dif4
    for k, v in overrides.items():
        setattr(cfg, k, v)
    Lm, Mm, N, NT = cfg.Lm, cfg.Mm, cfg.N, cfg.NT
    par = APP_PARAMS[app]
    g = Grid(Lm, Mm)
    I, J = g.I, g.J
    F = {}
    # ---- ana_grid.h
    if app == APP_BENCHMARK:
        Xsize, Esize, depth = 360.0, 20.0, 4000.0
    elif app == APP_SEAMOUNT:
        Xsize, Esize, depth, f0, beta = 320.0e3, 320.0e3, 5000.0, 1.0e-4, 0.0
    else:
        Xsize, Esize, depth, f0, beta = 1000.0 * Lm, 1000.0 * Mm, 150.0, -8.26e-5, 0.0
    dx, dy = Xsize / Lm, Esize / Mm
    ones = np.ones((g.nj, g.ni))
    if app == APP_BENCHMARK:
        latr = (-70.0 + dy * (J - 0.5)) * ones
        val1 = Lm / (2.0 * PI * ERADIUS)
        val2 = Mm * 360.0 / (2.0 * PI * ERADIUS * Esize)
        pm = val1 * (1.0 / np.cos((-70.0 + dy * (J - 0.5)) * DEG2RAD)) * ones
        pn = val2 * ones
        fcor = (2.0 * (2.0 * PI * 366.25 / 365.25) / 86400.0) * np.sin(latr * DEG2RAD)
        h = 500.0 + 1750.0 * (1.0 + np.tanh((68.0 + latr) / dy))
    else:
        xr = dx * ((I - 1.0) + 0.5) * ones
        yr = dy * ((J - 1.0) + 0.5) * ones
        pm = (1.0 / dx) * ones
        pn = (1.0 / dy) * ones
        fcor = f0 + beta * (yr - 0.5 * Esize)
        if app == APP_SEAMOUNT:
            v1 = (xr - 0.5 * Xsize) / 40000.0
            v2 = (yr - 0.5 * Esize) / 40000.0
            h = depth - 4500.0 * np.exp(-(v1 * v1 + v2 * v2))
        else:
            jj = np.where(J <= Mm // 2, J, Mm + 1 - J)
            h = np.minimum(depth, 84.5 + 66.526 * np.tanh((jj - 10.0) / 7.0)) * ones
    for a in (pm, pn, fcor, h):
        g.exchange(a)
    F["pm"], F["pn"], F["f"], F["h"] = pm, pn, fcor, h
    dndx, dmde = g.zeros(), g.zeros()
    if cfg.curvgrid:
        # ana_grid.h:761-766 on interior points; wrkX/wrkY = pm/pn evaluated analytically on j-1, j+1
        dmde[1:Mm + 1, :] = 0.5 * ((1.0 / pm[2:Mm + 2, :]) - (1.0 / pm[0:Mm, :]))
        g.exchange(dndx); g.exchange(dmde)
    F["dndx"], F["dmde"] = dndx, dmde
    # ---- metrics.F
    F["om_r"] = 1.0 / pm; F["on_r"] = 1.0 / pn; F["omn"] = 1.0 / (pm * pn); F["fomn"] = fcor * F["omn"]
    F["pnom_r"] = pn / pm; F["pmon_r"] = pm / pn

    def im1(a):
        b = np.empty_like(a); b[:, 1:] = a[:, :-1]; b[:, 0] = a[:, 0]; return b

    def jm1(a):
        b = np.empty_like(a); b[1:, :] = a[:-1, :]; b[0, :] = a[0, :]; return b

    pmu, pnu = im1(pm) + pm, im1(pn) + pn
    F["pmon_u"] = pmu / pnu; F["pnom_u"] = pnu / pmu; F["om_u"] = 2.0 / pmu; F["on_u"] = 2.0 / pnu
    pmv, pnv = jm1(pm) + pm, jm1(pn) + pn
    F["pmon_v"] = pmv / pnv; F["pnom_v"] = pnv / pmv; F["om_v"] = 2.0 / pmv; F["on_v"] = 2.0 / pnv
    pmp = jm1(im1(pm)) + im1(pm) + jm1(pm) + pm
    pnp = jm1(im1(pn)) + im1(pn) + jm1(pn) + pn
    F["pnom_p"] = pnp / pmp; F["pmon_p"] = pmp / pnp; F["om_p"] = 4.0 / pmp; F["on_p"] = 4.0 / pnp
    for n in ("pmon_v", "pnom_v", "om_v", "on_v", "pnom_p", "pmon_p", "om_p", "on_p"):
        F[n][0, :] = 0.0                       # v/psi-type arrays start at j = 1 (metrics.F loops over JstrP:JendT)
    for n in ("om_r", "on_r", "omn", "fomn", "pnom_r", "pmon_r", "pmon_u", "pnom_u", "om_u", "on_u", "pmon_v", "pnom_v", "om_v", "on_v",
              "pnom_p", "pmon_p", "om_p", "on_p"):
        g.exchange(F[n])
    # ---- ini_hmixcoef.F, mod_grid.F drag
    F["visc2_r"] = par["visc2"] * ones; F["visc2_p"] = par["visc2"] * ones
    for it in range(NT):
        F[f"diff2_{it}"] = par["tnu2"] * ones
        if cfg.ts_dif4:
            F[f"diff4_{it}"] = math.sqrt(abs(tnu4)) * ones      # ini_hmixcoef.F:293 with read_phypar.F:6905: diff4 = SQRT(ABS(tnu4))
    F["rdrag"] = (par["rdrg"] if cfg.uv_qdrag == 0 else 0.0) * ones
    F["rdrag2"] = (par["rdrg2"] if cfg.uv_qdrag == 1 else 0.0) * ones
    if cfg.uv_qdrag == 2:
        F["ZoBot"] = 0.02 * ones                               # UV_LOGDRAG: roms_*.in Zob = 0.02 m
    # ---- s-coordinate and depths at rest (set_depth.F:210-246, Zt_avg1 = 0)
    sc_r, Cs_r, sc_w, Cs_w = set_scoord(N, par["theta_s"], par["theta_b"])
    hc = par["Tcline"]
    cfg.hc = hc
    hinv = 1.0 / (hc + h)
    z_w = g.zeros(N + 1); z_r = g.zeros(N); Hz = g.zeros(N)
    z_w[0] = -h
    for k in range(1, N + 1):
        z_w[k] = 0.0 + (0.0 + h) * ((hc * sc_w[k] + Cs_w[k] * h) * hinv)
        z_r[k - 1] = 0.0 + (0.0 + h) * ((hc * sc_r[k] + Cs_r[k] * h) * hinv)
        Hz[k - 1] = z_w[k] - z_w[k - 1]
    F["z_w"], F["z_r"], F["Hz"] = z_w, z_r, Hz
    # ---- ana_initial.h (u = v = zeta = 0)
    if app == APP_BENCHMARK:
        v1 = (44.69 / 39.382) ** 2
        v2 = v1 * (cfg.rho0 * 800.0 / cfg.g) * (5.0e-5 / ((42.689 / 44.69) ** 2))
        T = v2 * np.exp(z_r / 800.0) * (0.6 - 0.4 * np.tanh(z_r / 800.0))
        S = 35.0 * np.ones_like(T)
    elif app == APP_SEAMOUNT:
        T = cfg.T0 + 7.5 * np.exp(z_r / 1000.0)
        S = None
    else:
        T = cfg.T0 + 8.0 * np.exp(z_r / 50.0)
        S = cfg.S0 * np.ones_like(T)
    for A in (T, S):                               # ini_fields -> t3dbc: zero-gradient closed walls (t3dbc_im.F:477-489,611-623)
        if A is not None:
            A[:, 0, :] = A[:, 1, :]
            A[:, Mm + 1, :] = A[:, Mm, :]
    for tl in (1, 2):                              # ini_fields: t(nnew) = t(nstp)
        F[f"t{tl}_0"] = T.copy()
        if NT >= 2:
            F[f"t{tl}_1"] = S.copy()
    # ---- background vertical mixing (mod_mixing.F:1422-1443): k = 1..N-1
    Akv = g.zeros(N + 1); Akv[1:N] = cfg.Akv_bak
    F["Akv"] = Akv
    for it in range(NT):
        a = g.zeros(N + 1); a[1:N] = cfg.Akt_bak[it]
        F[f"Akt_{it}"] = a
    # ---- surface momentum stress at t = 0 (ana_smflux.h); see `sustr_at`
    F["sustr"] = sustr_at(app, g, cfg, 0.0)
    F["svstr"] = g.zeros()
    if cfg.bulk_fluxes:
        F.update(atmosphere_at(g, cfg, 0.0))                   # bulk_flux replaces the analytical stress / heat flux
    if cfg.solar_source:
        F["Jwtype"] = 1.0 * ones                                 # roms_benchmark1.in WTYPE == 1
        F.setdefault("srflx", g.zeros())
    nfast, w1, w2 = set_weights(cfg.ndtfast)
    return cfg, F, (sc_r, Cs_r, sc_w, Cs_w), (nfast, w1, w2)


def sustr_at(app, g, cfg, tdays):
    """Kinematic surface stress sustr(i,j) (m2/s2) at model day tdays.
    UPWELLING: ROMS/Functionals/ana_smflux.h:316-330.  BENCHMARK (reduced physics set): the shipped application takes its
    stress from bulk_flux (out of scope), so a steady analytical zonal stress (0.1/rho0)*sin(pi*(j-0.5)/Mm) is used.
    SEAMOUNT: zero."""
    ones = np.ones((g.nj, g.ni))
    if app == APP_UPWELLING:
        wind = -0.1 * math.sin(PI * tdays / 4.0) / cfg.rho0 if tdays <= 2.0 else -0.1 / cfg.rho0
        return wind * ones
    if app == APP_BENCHMARK:
        return (0.1 / cfg.rho0 * np.sin(PI * (g.J - 0.5) / g.Mm)) * ones
    return 0.0 * ones


ATMOSPHERE = ["Uwind", "Vwind", "Tair", "Pair", "Hair", "rain", "cloud", "srflx"]
# the shipped benchmark.h set on top of the reduced one: BULK_FLUXES, LMD_MIXING (+SKPP, NONLOCAL, RIMIX, CONVEC, RI_SPLINES),
# SOLAR_SOURCE, BV_FREQUENCY, MIX_GEO_TS
FULL_BENCHMARK = dict(bv_frequency=1, eos_tderivative=1, solar_source=1, lmd_nonlocal=1, bulk_fluxes=1, lmd_mixing=1, mix_geo_ts=1)


def atmosphere_at(g, cfg, tdays):
    """The BENCHMARK analytical atmosphere set_data.F hands bulk_flux (global 2-D arrays, periodic images included):
    ana_winds.h (15 m/s Gaussian jet at 60S), ana_tair.h (4 degC), ana_pair.h (1025 mb), ana_humid.h (0.8), ana_rain.h (0),
    ana_cloud.h (0.6) and ana_srflux.h with ALBEDO (Zillman / Laevastu; day of year and hour from the model clock,
    dateclock.F caldate with TIME_REF = 0, DSTART = 0)."""
    Lm, Mm = g.Lm, g.Mm
    ones = np.ones((g.nj, g.ni))
    dx, dy = 360.0 / Lm, 20.0 / Mm
    latr = (-70.0 + dy * (g.J - 0.5)) * ones
    lonr = (dx * (g.I - 0.5)) * ones
    A = dict(Tair=4.0 * ones, Pair=1025.0 * ones, Hair=0.8 * ones, rain=0.0 * ones, cloud=0.6 * ones, Vwind=0.0 * ones)
    cff = 0.2 * (60.0 + latr)
    A["Uwind"] = 15.0 * np.exp(-cff * cff)
    whole = math.floor(tdays); frac = abs(tdays - whole)
    yday = float(1 + int(whole) % 365) + frac; hour = 24.0 * frac
    Dangle = 23.44 * math.cos((172.0 - yday) * 2.0 * PI / 365.2425) * DEG2RAD
    Hangle = (12.0 - hour) * PI / 12.0
    Rsolar = 1353.0 / (cfg.rho0 * 3985.0)                              # Csolar / (rho0*Cp), mod_scalars.F:431-432
    LatRad = latr * DEG2RAD
    zenith = np.sin(LatRad) * math.sin(Dangle) + np.cos(LatRad) * math.cos(Dangle) * np.cos(Hangle - lonr * DEG2RAD)
    e_sat = 10.0 ** ((0.7859 + 0.03477 * A["Tair"]) / (1.0 + 0.00412 * A["Tair"]))
    vap_p = e_sat * A["Hair"]
    zz = np.maximum(zenith, 0.0)
    sr = Rsolar * zz * zz * (1.0 - 0.6 * A["cloud"] ** 3) / ((zz + 2.7) * vap_p * 1.0e-3 + 1.085 * zz + 0.1)
    A["srflx"] = (1.0 - 0.06) * np.where(zenith > 0.0, sr, 0.0)
    for a in A.values():
        g.exchange(a)
    return A


def tile_slice(a, Lm, bounds):
    """Columns LBi..UBi of one tile from a global array (last axis = xi, global origin i = -2)."""
    lo, hi = bounds["LBi"] + 2, bounds["UBi"] + 3
    return np.ascontiguousarray(a[..., lo:hi])


def make_tile(app, Lm=0, Mm=0, N=0, strict=False, NtileI=1, tile=0, device=0, **overrides):
    """Create a device Tile holding a fresh run of `app` (the state main3d sees at its first step).
    Mirrors ROMS/Nonlinear/initial.F:271-574 + main3d.F:269-285: host set-up, upload, then set_depth, set_massflux, omega,
    rho_eos on the device."""
    from .ocean import Tile
    cfg, F, sc, (nfast, w1, w2) = build(app, Lm, Mm, N, **overrides)
    cfg.NtileI, cfg.NtileJ, cfg.tile, cfg.device = NtileI, 1, tile, device
    t = Tile(cfg, strict=strict)
    b = _lib.bounds(cfg.Lm, cfg.Mm, NtileI, 1, tile, distribute=NtileI > 1, strict=strict)
    for name, a in F.items():
        t.set(name, tile_slice(a, cfg.Lm, b))
    t.set_scoord(*sc)
    t.set_weights(nfast, w1, w2)
    idx = dict(iic=1, ntstart=1, ntfirst=1, nstp=1, nnew=2, nrhs=1, iif=1, indx1=1, kstp=1, krhs=1, knew=1, PREDICTOR=0, exit_flag=0,
               time=0.0, tdays=0.0)
    t.set_indices(idx)
    for ph in ("set_depth", "set_massflux", "omega", "rho_eos"):
        t.run_phase(ph)
    t.synth = dict(app=app, grid=Grid(cfg.Lm, cfg.Mm), bounds=b)
    return t
