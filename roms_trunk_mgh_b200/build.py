"""Build the sm_100a shared library in-tree (roms_trunk_mgh_b200/lib/).

Two variants of the same sources:
  libroms_b200.so         production: -O3, FMA contraction on
  libroms_b200_strict.so  parity aid: -fmad=false, so every product is rounded like the CPU oracle (which is compiled
                          -ffp-contract=off); used by the bit-exact index/branch tests, never by bench.py
nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "lib", "obj")
SOURCES = ["api.cu", "api_tile.cu", "api_nccl.cu", "k_glue.cu", "k_pre.cu", "k_rhs.cu", "k_step3d.cu", "k_step2d.cu", "k_step2d_loop.cu", "k_diag.cu", "k_mixgeo.cu", "k_physics.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _compile(src, obj, extra, log):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [src, os.path.join(HERE, "..", "include", "roms_b200.h")]
    if not _newer(deps, obj):
        return
    cmd = [_nvcc()] + ARCH + COMMON + extra + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed for " + src)


def build(variants=("prod", "strict"), verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    out = {}
    for var in variants:
        extra = ["-fmad=false"] if var == "strict" else []
        name = "libroms_b200_strict.so" if var == "strict" else "libroms_b200.so"
        target = os.path.join(LIBDIR, name)
        objs = []
        jobs = []
        for s in SOURCES:
            src = os.path.join(CSRC, s)
            obj = os.path.join(OBJDIR, s.replace(".cu", "") + "_" + var + ".o")
            objs.append(obj)
            jobs.append((src, obj, extra, obj + ".log"))
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
            list(ex.map(lambda a: _compile(*a), jobs))
        if _newer(objs, target):
            cmd = [_nvcc()] + ARCH + ["-shared", "-o", target] + objs + ["-lcudart", "-ldl"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("link failed")
        out[var] = target
        if verbose:
            print("built", target)
    return out


if __name__ == "__main__":
    build(verbose=True)
