"""TEST INFRASTRUCTURE: ctypes front end of tests/emu (host emulation of the barrier-free CUDA kernels)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DP = C.POINTER(C.c_double)
EMULATED = {"set_massflux": 1, "rho_eos": 2, "set_vbc": 3, "ana_vmix": 4, "omega": 5, "wvelocity": 6, "set_zeta": 7, "pre_step3d": 8,
            "prsgrd": 9, "t3dmix": 10, "rhs3d": 11, "uv3dmix": 12, "step2d": 13, "set_depth": 14, "step3d_uv": 15, "omega2": 16, "step3d_t": 17,
            "diag": 18, "bulk_flux": 23, "lmd_vmix": 24, "bvf_mix": 25, "t3dmix4": 26}
IOPT = ["Lm", "Mm", "N", "NT", "nonlin_eos", "curvgrid", "uv_qdrag", "salinity", "hadv", "vadv", "itemp", "isalt", "bv_frequency",
        "eos_tderivative", "solar_source", "lmd_nonlocal", "bulk_fluxes", "lmd_mixing", "uv_adv", "ts_dif4", "dj_gradps", "mix_geo_ts",
        "ana_vmix", "ndtfast", "limit_bstress", "NtileI_", "tile_", "nospl_vvisc", "nospl_vdiff", "qcorrection", "limit_stflx_cooling", "scorrection", "bodyforce", "levsfrc", "levbfrc", "Vtransform", "atm_press"]
DOPT = ["dt", "g", "rho0", "R0", "T0", "S0", "Tcoef", "Scoef", "gamma2", "lambda", "hc", "Akv_bak", "Akt_bak", "Akt_bak", "blk_ZQ", "blk_ZT", "blk_ZW", "Tnudg_salt"]
_L = None


def lib():
    global _L
    if _L is None:
        subprocess.run(["make", "-s", "-C", HERE], check=True)
        _L = C.CDLL(os.path.join(HERE, "_build", "libemu.so"))
        _L.emu_create.restype = C.c_void_p
        _L.emu_create.argtypes = [C.POINTER(C.c_int), DP]
        _L.emu_destroy.argtypes = [C.c_void_p]
        _L.emu_xfer.argtypes = [C.c_void_p, C.c_char_p, DP, C.c_int]
        _L.emu_levels.argtypes = [C.c_void_p, C.c_char_p]
        _L.emu_extent.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        _L.emu_diag.argtypes = [C.c_void_p, DP]
        _L.emu_scoord.argtypes = [C.c_void_p, C.c_int, DP, C.c_int]
        _L.emu_indices.argtypes = [C.c_void_p] + [C.c_int] * 4
        _L.emu_indices2d.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_double] * 3
        _L.emu_run.argtypes = [C.c_void_p, C.c_int]
    return _L


class EmuTile:
    """The kernels of one tile run thread by thread on the host, configured from an oracle instance."""

    def __init__(self, o, NtileI=1, tile=0):
        """NtileI > 1: xi-tile `tile` of a ring (arrays Istr-3 .. Iend+2 like the device's; set / get take GLOBAL arrays and move
        this tile's columns, so a set() from the oracle is a perfect halo exchange)."""
        self.L = lib()
        io = (C.c_int * len(IOPT))(*[(NtileI if k == "NtileI_" else tile if k == "tile_" else int(o.opt(k))) for k in IOPT])
        dv = (C.c_double * len(DOPT))(*[float(o.opt(k)) for k in DOPT])
        self.h = C.c_void_p(self.L.emu_create(io, dv))
        self.N, self.Lm, self.Mm = int(o.opt("N")), int(o.opt("Lm")), int(o.opt("Mm"))
        ext = (C.c_int * 2)()
        self.L.emu_extent(self.h, ext)
        self.LBi, self.ni = ext[0], ext[1]
        self.c0 = self.LBi + 2                                  # first column of the tile's arrays inside a global (-2 .. Lm+2) array
        nd = int(o.opt("ndtfast"))
        self.nfast, self.w1, self.w2 = int(o.opt("nfast")), o.vector(4, 2 * nd + 2), o.vector(5, 2 * nd + 2)
        for w in range(4):
            v = np.ascontiguousarray(o.vector(w, self.N + 1))
            self.L.emu_scoord(self.h, w, v.ctypes.data_as(DP), v.size)

    def close(self):
        if self.h:
            self.L.emu_destroy(self.h); self.h = None

    def set(self, name, arr):
        assert arr.shape == (self.L.emu_levels(self.h, name.encode()), self.Mm + 2, self.Lm + 5), (name, arr.shape)
        a = np.ascontiguousarray(arr[:, :, self.c0:self.c0 + self.ni], dtype=np.float64)
        assert self.L.emu_xfer(self.h, name.encode(), a.ctypes.data_as(DP), 1) == 0, name

    def get(self, name):
        out = np.empty((self.L.emu_levels(self.h, name.encode()), self.Mm + 2, self.ni))
        assert self.L.emu_xfer(self.h, name.encode(), out.ctypes.data_as(DP), 0) == 0, name
        return out                                              # this tile's columns LBi .. LBi + ni - 1 (the whole array for one tile)

    def set_indices(self, d):
        istart = 0 if d["iic"] == d["ntfirst"] else (1 if d["iic"] == d["ntfirst"] + 1 else 2)
        self.L.emu_indices(self.h, d["nstp"], d["nnew"], d["nrhs"], istart)
        iif = d["iif"]
        w = lambda a, i: float(a[i]) if 0 <= i < len(a) else 0.0
        self.L.emu_indices2d(self.h, iif, d["kstp"], d["krhs"], d["knew"], d["PREDICTOR"], self.nfast, w(self.w1, iif - 1), w(self.w2, iif), w(self.w2, iif + 1))

    def diag(self):
        """The scalars of roms_b200_diag (diag.F), through the three device kernels and the host-side finish."""
        self.run_phase("diag")
        out = (C.c_double * 12)()
        self.L.emu_diag(self.h, out)
        return dict(zip(["avgke", "avgpe", "avgkp", "volume", "max_speed", "maxCu", "maxCv", "maxCw", "ubarmax", "vbarmax", "umax", "vmax"], list(out)))

    def run_phase(self, name):
        rc = self.L.emu_run(self.h, EMULATED[name])
        assert rc == 0, (name, rc)
