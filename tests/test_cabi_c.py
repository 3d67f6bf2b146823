"""The drop-in boundary seen from compiled hosts: tests/c_driver.c (plain C99) and fortran/mod_b200.F (ISO_C_BINDING shim).

CPU: the C driver compiles against include/roms_b200.h with -std=c99 -Wall -Werror and links against the library; the
struct layouts the three hosts assume agree (C sizeof/offsetof == ctypes == member order of the Fortran BIND(C) types); the
shim declares an INTERFACE for every symbol it uses and the driver patch applies to the files it names.
GPU: the C driver plays main3d (routine by routine through roms_b200_routine_tile, and in the resident form) and must leave
the state bit-identical to the oracle's."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import orc
from roms_trunk_mgh_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "roms_trunk_mgh_b200", "lib")


def build_driver(strict=True):
    _lib.load(strict)                                   # fails loudly if the library has not been built
    exe = os.path.join(LIBDIR, "c_driver_strict" if strict else "c_driver")
    lib = "roms_b200_strict" if strict else "roms_b200"
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "c_driver.c"), "-o", exe,
           "-L", LIBDIR, "-l" + lib, "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_driver_compiles_and_struct_layouts_agree():
    exe = build_driver(True)
    out = subprocess.run([exe, "layout"], capture_output=True, text=True, check=True).stdout
    lay = dict(ln.split() for ln in out.strip().splitlines())
    assert int(lay["sizeof.config"]) == C.sizeof(_lib.Config) and int(lay["sizeof.tile"]) == C.sizeof(_lib.TileArgs)
    for name, _ in _lib.Config._fields_:
        cname = "lambda" if name == "lambda_" else name
        assert int(lay["config." + cname]) == getattr(_lib.Config, name).offset, name
    for name, _ in _lib.TileArgs._fields_:
        assert int(lay["tile." + name]) == getattr(_lib.TileArgs, name).offset, name
    # the Fortran BIND(C) types list the same members in the same order (the companion C processor then lays them out alike)
    src = open(os.path.join(ROOT, "fortran", "mod_b200.F")).read()

    def members(tname):
        body = re.search(r"TYPE, BIND\(C\) :: %s\n(.*?)END TYPE" % tname, src, re.S).group(1)
        out = []
        for ln in body.splitlines():
            if "::" in ln:
                out += [m.split("(")[0].strip() for m in re.split(r",\s*(?![^()]*\))", ln.split("::")[1])]
        return out
    assert members("roms_b200_config") == [("lambda" if n == "lambda_" else n) for n, _ in _lib.Config._fields_]
    assert members("roms_b200_tile_t") == [n for n, _ in _lib.TileArgs._fields_]


def test_fortran_shim_is_consistent_with_the_header():
    """Every roms_b200_* function the shim calls has an INTERFACE block whose BIND(C) name is declared in the header, and every
    phase constant equals the header's enum value."""
    src = open(os.path.join(ROOT, "fortran", "mod_b200.F")).read()
    hdr = open(os.path.join(ROOT, "include", "roms_b200.h")).read()
    bound = set(re.findall(r"BIND\(C, NAME='(\w+)'\)", src))
    for sym in bound - {"strlen"}:
        assert re.search(r"\b%s\s*\(" % sym, hdr), sym
        assert sym in _lib.EXPORTS, sym
    used = set(re.findall(r"\b(roms_b200_\w+)\s*\(", src))
    assert used <= bound | {"roms_b200_config", "roms_b200_tile_t"}, used - bound
    for name, val in re.findall(r"parameter :: B200_(\w+)\s*=\s*(\d+)", src):
        m = re.search(r"ROMS_B200_%s = (\d+)" % name, hdr)
        assert m and int(m.group(1)) == int(val), name
    # the by-name vocabulary of b200_loc covers every field name the routines ask for
    cases = set(re.findall(r"CASE \('(\w+)'\)", src))
    L = _lib.load(True)
    for ph in list(range(1, 18)) + [23, 24, 25, 26]:
        spec = L.roms_b200_routine_args(ph)
        if not spec:
            continue
        for part in spec.decode().split(";"):
            for n in part.split(":")[1].split(","):
                base = n.lstrip("?").replace("_*", "")
                base = base if base in cases else base.rstrip("123")
                assert base in cases, (ph, n)


def test_driver_patch_names_existing_call_sites():
    """fortran/patches/b200_drivers.patch only adds lines, one CALL b200_routine per driver of the chain."""
    p = open(os.path.join(ROOT, "fortran", "patches", "b200_drivers.patch")).read()
    files = re.findall(r"^\+\+\+ b/(\S+)", p, re.M)
    assert len(files) == 22 and len(set(files)) == 22                 # the 18 drivers of the chain + bulk_flux.F, lmd_vmix.F, t3dmix4_s.h, prsgrd40.h
    assert len(re.findall(r"^\+\s+CALL b200_routine \(ng, tile, B200_\w+\)", p, re.M)) == 22
    assert not re.findall(r"^-(?!--)", p, re.M)                        # nothing of the reference is removed
    ref = "/root/reference"
    if os.path.isdir(ref):                                             # not on the GPU box
        import shutil
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            for f in files:
                os.makedirs(os.path.join(td, os.path.dirname(f)), exist_ok=True)
                shutil.copy(os.path.join(ref, f), os.path.join(td, f))
            r = subprocess.run(["patch", "-p1", "-i", os.path.join(ROOT, "fortran", "patches", "b200_drivers.patch")], cwd=td,
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr


# ---- GPU: the C driver against the oracle --------------------------------------------------------------------------
def write_state(path, o, cfg):
    from helpers import optional_names
    from roms_trunk_mgh_b200.ocean import field_names
    NT, N, nd = int(o.opt("NT")), int(o.opt("N")), int(o.opt("ndtfast"))
    n2, n3 = field_names(NT)
    recs = [(n, np.ascontiguousarray(o.field(n))) for n in n2 + n3 + optional_names(o)]
    raw = bytes(cfg) + b"\0" * (-C.sizeof(cfg) % 8)
    recs.append(("@config", np.frombuffer(raw, dtype=np.float64).reshape(1, 1, -1)))
    d = o.indices()
    recs.append(("@indices", np.array([float(d[k]) for k in orc.INDEX_NAMES] + [d["time"], d["tdays"]]).reshape(1, 1, -1)))
    for w, nm in enumerate(("@sc_r", "@Cs_r", "@sc_w", "@Cs_w")):
        recs.append((nm, o.vector(w, N + 1).reshape(1, 1, -1)))
    recs.append(("@w1", o.vector(4, 2 * nd + 2).reshape(1, 1, -1)))
    recs.append(("@w2", o.vector(5, 2 * nd + 2).reshape(1, 1, -1)))
    recs.append(("@nfast", np.array([o.opt("nfast")]).reshape(1, 1, 1)))
    with open(path, "wb") as fh:
        fh.write(b"RB2S" + struct.pack("<i", len(recs)))
        for n, a in recs:
            a = np.ascontiguousarray(a, dtype=np.float64)
            fh.write(n.encode().ljust(24, b"\0") + struct.pack("<3i", *a.shape))
            fh.write(a.tobytes())


def read_state(path):
    out = {}
    with open(path, "rb") as fh:
        assert fh.read(4) == b"RB2S"
        (n,) = struct.unpack("<i", fh.read(4))
        for _ in range(n):
            name = fh.read(24).rstrip(b"\0").decode()
            shp = struct.unpack("<3i", fh.read(12))
            out[name] = np.frombuffer(fh.read(8 * shp[0] * shp[1] * shp[2]), dtype=np.float64).reshape(shp)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["routine", "resident"])
@pytest.mark.parametrize("case", ["benchmark", "seamount", "benchmark_full"])
def test_c_driver_plays_main3d_bit_exact(tmp_path, mode, case):
    """benchmark_full: the shipped benchmark.h cpp set -- the C host also calls bulk_flux and lmd_vmix by routine; one step (the
    host would refresh the shortwave flux between steps), fields within 1e-9 (device pow / exp / log / atan), ksbl exactly."""
    from helpers import cfg_from_oracle
    full = dict(Lm=48, Mm=32, N=30, bv_frequency=1, eos_tderivative=1, solar_source=1, lmd_nonlocal=1, bulk_fluxes=1, lmd_mixing=1, mix_geo_ts=1)
    app, kw = {"benchmark": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10)), "seamount": (orc.APP_SEAMOUNT, {}),
               "benchmark_full": (orc.APP_BENCHMARK, full)}[case]
    exe = build_driver(True)
    o = orc.Oracle(app, **kw)
    o.run_phase("set_data"); o.run_phase("ini")
    o.step(5 if case == "benchmark_full" else 1)          # start from a state with live momentum (AB2 branch next)
    cfg = cfg_from_oracle(o)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_state(fin, o, cfg)
    nsteps = 1 if case == "benchmark_full" else 3
    if case == "benchmark_full":                          # the set_data products of the step the C host is about to play
        d = o.indices(); d["tdays"] = d["time"] / 86400.0; o.set_indices(d)
        o.run_phase("set_data")
        write_state(fin, o, cfg)
    r = subprocess.run([exe, fin, fout, mode, str(nsteps)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    o.step(nsteps)
    got = read_state(fout)
    d = o.indices()
    for k, v in zip(orc.INDEX_NAMES, got["@indices"].ravel()[:13]):
        if k not in ("nstp", "nnew", "nrhs") or mode == "routine":
            assert int(v) == d[k], (k, v, d[k])              # the time-index state machine, run by the C host
    if case == "benchmark_full":
        def rel(a, b):
            return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300)
        assert np.array_equal(got["ksbl"], o.field("ksbl"))
        bad = [(n, rel(got[n], o.field(n))) for n in got if not n.startswith("@") and n not in ("ksbl", "W", "wvel") and rel(got[n], o.field(n)) > 1e-9]
    else:
        bad = [n for n in got if not n.startswith("@") and not np.array_equal(got[n], o.field(n))]
    assert not bad, bad
