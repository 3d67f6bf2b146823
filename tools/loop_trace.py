"""Tuning aid: stage-boundary time stamps of one CTA of the persistent barotropic-loop kernel (library built with -DLK_TRACE).
    python tools/variant.py trace k_step2d_loop.cu,api.cu -DLK_TRACE
    ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_trace.so python tools/loop_trace.py 256 256 30
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roms_trunk_mgh_b200 import synth
Lm, Mm, N = (int(x) for x in sys.argv[1:4])
t = synth.make_tile(synth.APP_BENCHMARK, Lm, Mm, N)
t.main3d(6)
out = (C.c_ulonglong * 64)()
t.L.roms_b200_debug_loop_trace.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
assert t.L.roms_b200_debug_loop_trace(t.h, out) == 0
a = np.array(list(out), dtype=np.int64).reshape(8, 8)
names = ["start", "synced", "stage0 loaded", "stage1 done", "stage2 done+stage3 loads", "stage3 done", "barrier", "flag published"]
print("call: ns since the call's start ->", names[1:])
for r in range(8):
    print(20 + r, [int(a[r, k] - a[r, 0]) for k in range(1, 8)], "next call starts +", int(a[r + 1, 0] - a[r, 0]) if r < 7 else "")
