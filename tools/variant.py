"""Kernel-tuning aid: build an alternative production library with extra -D flags for one source file.

    python tools/variant.py NAME k_step3d.cu[,k_pre.cu,...] -DTS=128 -DCH=6     ->  roms_trunk_mgh_b200/lib/var/libroms_b200_NAME.so
    ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_NAME.so python bench.py ...
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roms_trunk_mgh_b200 import build as B  # noqa: E402


def main():
    name, srcs, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
    B.build(variants=("prod",))
    vdir = os.path.join(B.LIBDIR, "var")
    os.makedirs(vdir, exist_ok=True)
    vobj = {}
    for src in srcs:
        obj = os.path.join(vdir, src.replace(".cu", "") + "_" + name + ".o")
        cmd = [B._nvcc()] + B.ARCH + B.COMMON + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stdout + r.stderr)
        vobj[src] = obj
    objs = [vobj.get(s, os.path.join(B.OBJDIR, s.replace(".cu", "") + "_prod.o")) for s in B.SOURCES]
    target = os.path.join(vdir, f"libroms_b200_{name}.so")
    subprocess.run([B._nvcc()] + B.ARCH + ["-shared", "-o", target] + objs + ["-lcudart", "-ldl"], check=True)
    print("built", target)


if __name__ == "__main__":
    main()
