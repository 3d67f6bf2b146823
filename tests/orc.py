"""ctypes binding to the CPU oracle (oracle/, TEST INFRASTRUCTURE ONLY).

Builds oracle/_build/liboracle.so on demand with the recipe in oracle/Makefile.  Nothing under
roms_trunk_mgh_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

APP_UPWELLING, APP_SEAMOUNT, APP_BENCHMARK = 0, 1, 2
PHASES = dict(set_massflux=1, rho_eos=2, set_vbc=3, ana_vmix=4, omega=5, wvelocity=6, set_zeta=7, pre_step3d=8, prsgrd=9,
              t3dmix=10, rhs3d=11, uv3dmix=12, step2d=13, set_depth=14, step3d_uv=15, omega2=16, step3d_t=17, diag=18,
              set_data=19, step2d_loop=20, ini=21, set_avg=22, bulk_flux=23, lmd_vmix=24, bvf_mix=25, t3dmix4=26)
INDEX_NAMES = ["iic", "ntstart", "ntfirst", "nstp", "nnew", "nrhs", "iif", "indx1", "kstp", "krhs", "knew", "PREDICTOR", "exit_flag"]
BOUND_NAMES = ["tile", "Itile", "Jtile", "LBi", "UBi", "LBj", "UBj", "IminS", "ImaxS", "JminS", "JmaxS", "Istr", "IstrB", "IstrP",
               "IstrR", "IstrT", "IstrM", "IstrU", "Iend", "IendB", "IendP", "IendR", "IendT", "Jstr", "JstrB", "JstrP", "JstrR",
               "JstrT", "JstrM", "JstrV", "Jend", "JendB", "JendP", "JendR", "JendT", "Istrm3", "Istrm2", "Istrm1", "IstrUm2",
               "IstrUm1", "Iendp1", "Iendp2", "Iendp2i", "Iendp3", "Jstrm3", "Jstrm2", "Jstrm1", "JstrVm2", "JstrVm1", "Jendp1",
               "Jendp2", "Jendp2i", "Jendp3", "Western_Edge", "Eastern_Edge", "Southern_Edge", "Northern_Edge"]

_libs = {}


def cpu_tag():
    """Short hash of this host's CPU feature flags: names the -march=native builds, so that a library built on another
    host (the build container vs. the GPU box) is rebuilt here instead of dying with an illegal instruction."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as fh:
            flags = next((ln for ln in fh if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return "_" + hashlib.sha1(flags.encode()).hexdigest()[:8]


def build(kind="parity"):
    tag = cpu_tag()
    target = {"parity": "_build/liboracle.so", "fast": f"_build/liboracle_fast{tag}.so", "fastmath": f"_build/liboracle_fastmath{tag}.so",
              "chk": "_build/liboracle_chk.so"}[kind]
    subprocess.run(["make", "-s", f"-j{min(os.cpu_count() or 1, 12)}", "-C", ORACLE_DIR, f"TAG={tag}", target], check=True)
    return os.path.join(ORACLE_DIR, target)


def lib(kind="parity"):
    if kind in _libs:
        return _libs[kind]
    L = C.CDLL(build(kind))
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int] * 6
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.orc_get_option.restype = C.c_double
    L.orc_get_option.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_init.argtypes = [C.c_void_p]
    L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_run_phase.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_get_indices.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.orc_set_indices.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.orc_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int)]
    L.orc_vector.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int]
    L.orc_diag.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.orc_bounds.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)]
    L.orc_eos_point.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)]
    L.orc_physics_point.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_set_weights.restype = C.c_int
    L.orc_set_weights.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_timed_steps.restype = C.c_double
    L.orc_timed_steps.argtypes = [C.c_void_p, C.c_int, C.c_int]
    _libs[kind] = L
    return L


class Oracle:
    """One oracle model instance.  Fields are exposed as zero-copy numpy views indexed [k, j, i] (C order of the
    Fortran (i,j,k) layout); use .origin(name) for the Fortran lower bounds."""

    def __init__(self, app, Lm=0, Mm=0, N=0, NtileI=1, NtileJ=1, kind="parity", **options):
        self.L = lib(kind)
        self.h = C.c_void_p(self.L.orc_create(app, Lm, Mm, N, NtileI, NtileJ))
        for k, v in options.items():
            if self.L.orc_set_option(self.h, k.encode(), float(v)) != 0:
                raise KeyError(k)
        self.L.orc_init(self.h)
        self._dims = {}

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def opt(self, key):
        v = self.L.orc_get_option(self.h, key.encode())
        if v < -1e299:
            raise KeyError(key)
        return v

    def field(self, name):
        p = C.POINTER(C.c_double)()
        dims = (C.c_int * 6)()
        if self.L.orc_field(self.h, name.encode(), C.byref(p), dims) != 0:
            raise KeyError(name)
        LBi, ni, LBj, nj, LBk, nk = list(dims)
        self._dims[name] = (LBi, LBj, LBk)
        return np.ctypeslib.as_array(p, shape=(nk, nj, ni))

    def origin(self, name):
        if name not in self._dims:
            self.field(name)
        return self._dims[name]

    def vector(self, which, n):
        out = np.zeros(n)
        self.L.orc_vector(self.h, which, out.ctypes.data_as(C.POINTER(C.c_double)), n)
        return out

    def step(self, n=1, nthreads=1):
        self.L.orc_step(self.h, n, nthreads)

    def run_phase(self, name, nthreads=1):
        self.L.orc_run_phase(self.h, PHASES[name], nthreads)

    def indices(self):
        idx = (C.c_int * 13)()
        tm = (C.c_double * 2)()
        self.L.orc_get_indices(self.h, idx, tm)
        d = dict(zip(INDEX_NAMES, list(idx)))
        d["time"], d["tdays"] = tm[0], tm[1]
        return d

    def set_indices(self, d):
        idx = (C.c_int * 13)(*[int(d[k]) for k in INDEX_NAMES])
        tm = (C.c_double * 2)(d["time"], d["tdays"])
        self.L.orc_set_indices(self.h, idx, tm)

    def diag(self):
        out = (C.c_double * 12)()
        self.L.orc_diag(self.h, out)
        return dict(zip(["avgke", "avgpe", "avgkp", "volume", "max_speed", "maxCu", "maxCv", "maxCw", "ubarmax", "vbarmax", "umax", "vmax"], list(out)))

    def timed_steps(self, n, nthreads):
        return self.L.orc_timed_steps(self.h, n, nthreads)


def bounds(Lm, Mm, NtileI, NtileJ, tile, distribute=False, kind="parity"):
    out = (C.c_int * 57)()
    lib(kind).orc_bounds(Lm, Mm, NtileI, NtileJ, tile, int(distribute), out)
    return dict(zip(BOUND_NAMES, list(out)))


def eos_point(T, S, z, kind="parity"):
    out = (C.c_double * 3)()
    lib(kind).orc_eos_point(T, S, z, out)
    return out[0], out[1], out[2]


def physics_point(which, *args, kind="parity"):
    """which = 0: (bulk_psiu, bulk_psit)(Z/L); 1: lmd_swfrac(depth Z >= 0, Jerlov type); 2: (wm, ws)(Ustar, sigma, Bflux)."""
    a = (C.c_double * 4)(*[float(x) for x in args] + [0.0] * (4 - len(args)))
    out = (C.c_double * 2)()
    lib(kind).orc_physics_point(which, a, out)
    return out[0], out[1]


def set_weights(ndtfast, dt=1.0, kind="parity"):
    chk = (C.c_double * 5)()
    w1 = (C.c_double * (2 * ndtfast + 2))()
    w2 = (C.c_double * (2 * ndtfast + 2))()
    nfast = lib(kind).orc_set_weights(ndtfast, dt, chk, w1, w2)
    return nfast, list(chk), np.array(w1), np.array(w2)


def state_field_names(NT):
    n2 = ["zeta1", "zeta2", "zeta3", "ubar1", "ubar2", "ubar3", "vbar1", "vbar2", "vbar3", "rzeta1", "rzeta2", "rubar1", "rubar2",
          "rvbar1", "rvbar2", "Zt_avg1", "DU_avg1", "DU_avg2", "DV_avg1", "DV_avg2", "rufrc", "rvfrc", "rhoA", "rhoS", "sustr", "svstr",
          "bustr", "bvstr"]
    n3 = ["u1", "u2", "v1", "v2", "ru1", "ru2", "rv1", "rv2", "rho", "pden", "Hz", "z_r", "z_w", "Huon", "Hvom", "W", "wvel", "Akv"]
    for it in range(NT):
        n2 += [f"stflx_{it}", f"btflx_{it}", f"stflux_{it}", f"btflux_{it}"]
        n3 += [f"t1_{it}", f"t2_{it}", f"t3_{it}", f"Akt_{it}"]
    return n2, n3
