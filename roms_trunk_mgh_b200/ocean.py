"""Host-side mirror of the reference's main3d driver surface for ONE tile on ONE B200.

`Tile` owns a device-resident copy of the ROMS state (OCEAN, GRID, COUPLING, MIXING, FORCES members used by
ROMS/Nonlinear/main3d.F:307-814) and exposes the routines of that path under their reference names
(set_massflux, rho_eos, ..., step2d, step3d_uv, step3d_t) plus `main3d(nsteps)`.  All numerics run in the CUDA
library behind include/roms_b200.h; this module only moves whole Fortran arrays and integers across the C-ABI.
Arrays are numpy float64 with shape (nk, nj, ni) == Fortran A(LBi:UBi, LBj:UBj, k) in memory order.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DIAG_NAMES, INDEX_NAMES, PHASES

EXIT_FLAG = {0: "NoError", 1: "blow-up", 2: "input error", 5: "configuration error", 8: "fatal algorithm error (CUDA/NCCL)"}


class RomsB200Error(RuntimeError):
    def __init__(self, where, rc):
        super().__init__(f"{where}: exit_flag={rc} ({EXIT_FLAG.get(rc, '?')})")
        self.exit_flag = rc


def _dp(a):
    return a.ctypes.data_as(_lib.DP)


class Tile:
    def __init__(self, cfg, strict=False):
        self.L = _lib.load(strict)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.roms_b200_create(C.byref(cfg), C.byref(self.h))
        if rc:
            self.h = None
            raise RomsB200Error("roms_b200_create", rc)
        ab = (C.c_int * 4)()
        self.L.roms_b200_array_bounds(self.h, ab)
        self.LBi, self.UBi, self.LBj, self.UBj = list(ab)
        self.ni, self.nj = self.UBi - self.LBi + 1, self.UBj - self.LBj + 1
        self.N, self.NT = cfg.N, cfg.NT

    def close(self):
        if getattr(self, "h", None):
            self.L.roms_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, where, rc):
        if rc:
            raise RomsB200Error(where, rc)

    # ---- data movement -------------------------------------------------------------------------------------
    def nk_of(self, name):
        """Number of vertical planes of a field: 1 (2-D), N (rho/u/v levels) or N+1 (W levels 0:N) -- roms_b200_field_levels."""
        if not hasattr(self, "_nk"):
            self._nk = {}
        if name not in self._nk:
            lbk, nk = C.c_int(), C.c_int()
            self._ck(f"field_levels({name})", self.L.roms_b200_field_levels(self.h, name.encode(), C.byref(lbk), C.byref(nk)))
            self._nk[name] = nk.value
        return self._nk[name]

    def set(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self._ck(f"set_field({name})", self.L.roms_b200_set_field(self.h, name.encode(), _dp(a), a.size))

    def get(self, name):
        out = np.empty((self.nk_of(name), self.nj, self.ni))
        self._ck(f"get_field({name})", self.L.roms_b200_get_field(self.h, name.encode(), _dp(out), out.size))
        return out

    def set_scoord(self, sc_r, Cs_r, sc_w, Cs_w):
        for which, v in enumerate((sc_r, Cs_r, sc_w, Cs_w)):
            a = np.ascontiguousarray(v, dtype=np.float64)
            self._ck("set_scoord", self.L.roms_b200_set_scoord(self.h, which, _dp(a), a.size))

    def set_weights(self, nfast, w1, w2):
        a, b = np.ascontiguousarray(w1, dtype=np.float64), np.ascontiguousarray(w2, dtype=np.float64)
        self._ck("set_weights", self.L.roms_b200_set_weights(self.h, int(nfast), _dp(a), _dp(b), a.size))

    def indices(self):
        idx = (C.c_int * 13)()
        tm = (C.c_double * 2)()
        self._ck("get_indices", self.L.roms_b200_get_indices(self.h, idx, tm))
        d = dict(zip(INDEX_NAMES, list(idx)))
        d["time"], d["tdays"] = tm[0], tm[1]
        return d

    def set_indices(self, d):
        idx = (C.c_int * 13)(*[int(d[k]) for k in INDEX_NAMES])
        tm = (C.c_double * 2)(float(d["time"]), float(d["tdays"]))
        self._ck("set_indices", self.L.roms_b200_set_indices(self.h, idx, tm))

    # ---- the path ----------------------------------------------------------------------------------------------
    def run_phase(self, name):
        self._ck(f"run_phase({name})", self.L.roms_b200_run_phase(self.h, PHASES[name]))

    def main3d(self, nsteps=1, sync=True):
        """nsteps baroclinic steps, device resident (ROMS/Nonlinear/main3d.F:189-917)."""
        self._ck("main3d_step", self.L.roms_b200_main3d_step(self.h, int(nsteps)))
        if sync:
            self.sync()

    def set_avg(self, nAVG, ntsAVG=1):
        """AVERAGES (set_avg.F): accumulate time averages over windows of nAVG steps on the device (0: off)."""
        self._ck("set_avg", self.L.roms_b200_set_avg(self.h, int(nAVG), int(ntsAVG)))

    def set_option(self, key, value):
        """roms_b200_set_option: "cuda_graphs", "step2d_exchange", "overlap", "halo_timeout_s"."""
        self._ck(f"set_option({key})", self.L.roms_b200_set_option(self.h, key.encode(), float(value)))

    def sync(self):
        self._ck("sync", self.L.roms_b200_sync(self.h))

    def last_step_ms(self):
        ms = C.c_float()
        self._ck("last_step_ms", self.L.roms_b200_last_step_ms(self.h, C.byref(ms)))
        return ms.value

    def step_forced(self, sustr=None, svstr=None, stflux_temp=None):
        """One step the way main3d sees it from the host: forcing H2D, the step, diag scalars D2H."""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (sustr, svstr, stflux_temp)]
        ptrs = [None if a is None else _dp(a) for a in arrs]
        out = (C.c_double * 12)()
        rc = self.L.roms_b200_step_forced(self.h, ptrs[0], ptrs[1], ptrs[2], self.ni * self.nj, out)
        if rc not in (0, 1):
            raise RomsB200Error("step_forced", rc)
        return dict(zip(DIAG_NAMES, list(out))), rc

    def step_fields(self, fields):
        """step_forced with any set of 2-D forcing arrays by name (roms_b200_step_fields), e.g. the atmosphere bulk_flux reads."""
        names = list(fields)
        arrs = [np.ascontiguousarray(fields[n], dtype=np.float64) for n in names]
        cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
        cp = (_lib.DP * len(names))(*[_dp(a) for a in arrs])
        out = (C.c_double * 12)()
        rc = self.L.roms_b200_step_fields(self.h, len(names), cn, cp, self.ni * self.nj, out)
        if rc not in (0, 1):
            raise RomsB200Error("step_fields", rc)
        return dict(zip(DIAG_NAMES, list(out))), rc

    def register_host(self, *arrays):
        """Pin caller-owned forcing arrays (they must outlive the Tile or be released with unregister_host): step_forced
        then copies from them directly.  Mirrors what a Fortran host does once for FORCES(ng)%sustr, %svstr, %stflux."""
        for a in arrays:
            assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
            self._ck("register_host", self.L.roms_b200_register_host(self.h, a.ctypes.data, a.nbytes))
            self._pinned = getattr(self, "_pinned", []) + [a]          # keep the arrays alive

    def diag(self):
        out = (C.c_double * 12)()
        rc = self.L.roms_b200_diag(self.h, out)
        if rc not in (0, 1):
            raise RomsB200Error("diag", rc)
        return dict(zip(DIAG_NAMES, list(out)))

    def profile(self, on=True):
        self._ck("profile_enable", self.L.roms_b200_profile_enable(self.h, int(on)))

    def profile_get(self):
        ms = (C.c_double * 32)()
        n = C.c_longlong()
        self._ck("profile_get", self.L.roms_b200_profile_get(self.h, ms, C.byref(n)))
        inv = {v: k for k, v in PHASES.items()}
        return {inv.get(i, str(i)): ms[i] for i in range(32) if ms[i] > 0.0}, n.value

    def launch_count(self):
        return int(self.L.roms_b200_launch_count(self.h))

    # reference routine names (ROMS/Nonlinear/main3d.F USE ... ONLY list, :94-124)
    def set_massflux(self): self.run_phase("set_massflux")
    def rho_eos(self): self.run_phase("rho_eos")
    def set_vbc(self): self.run_phase("set_vbc")
    def ana_vmix(self): self.run_phase("ana_vmix")
    def bulk_flux(self): self.run_phase("bulk_flux")
    def lmd_vmix(self): self.run_phase("lmd_vmix")
    def bvf_mix(self): self.run_phase("bvf_mix")
    def omega(self): self.run_phase("omega")
    def wvelocity(self): self.run_phase("wvelocity")
    def set_zeta(self): self.run_phase("set_zeta")
    def pre_step3d(self): self.run_phase("pre_step3d")
    def prsgrd(self): self.run_phase("prsgrd")
    def t3dmix2(self): self.run_phase("t3dmix")
    def rhs3d_tile(self): self.run_phase("rhs3d")
    def uv3dmix2(self): self.run_phase("uv3dmix")
    def step2d(self): self.run_phase("step2d")
    def set_depth(self): self.run_phase("set_depth")
    def step3d_uv(self): self.run_phase("step3d_uv")
    def step3d_t(self): self.run_phase("step3d_t")

    def rhs3d(self):
        """The rhs3d driver (ROMS/Nonlinear/rhs3d.F:25-167): pre_step3d, prsgrd, t3dmix2, rhs3d_tile, uv3dmix2."""
        for ph in ("pre_step3d", "prsgrd", "t3dmix", "rhs3d", "uv3dmix"):
            self.run_phase(ph)


def field_names(NT):
    """All transferable field names (2-D list, 3-D list)."""
    n2 = ["h", "f", "pm", "pn", "om_r", "on_r", "om_u", "on_u", "om_v", "on_v", "om_p", "on_p", "omn", "fomn", "pmon_r", "pnom_r",
          "pmon_u", "pnom_u", "pmon_v", "pnom_v", "pmon_p", "pnom_p", "dndx", "dmde", "rdrag", "rdrag2", "visc2_r", "visc2_p",
          "zeta1", "zeta2", "zeta3", "ubar1", "ubar2", "ubar3", "vbar1", "vbar2", "vbar3", "rzeta1", "rzeta2", "rubar1", "rubar2",
          "rvbar1", "rvbar2", "Zt_avg1", "DU_avg1", "DU_avg2", "DV_avg1", "DV_avg2", "rufrc", "rvfrc", "rhoA", "rhoS", "sustr", "svstr",
          "bustr", "bvstr"]
    n3 = ["u1", "u2", "v1", "v2", "ru1", "ru2", "rv1", "rv2", "rho", "pden", "Hz", "z_r", "z_w", "Huon", "Hvom", "W", "wvel", "Akv"]
    for it in range(NT):
        n2 += [f"diff2_{it}", f"stflx_{it}", f"btflx_{it}", f"stflux_{it}", f"btflux_{it}"]
        n3 += [f"t1_{it}", f"t2_{it}", f"t3_{it}", f"Akt_{it}"]
    return n2, n3
