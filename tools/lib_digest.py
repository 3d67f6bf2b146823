"""Kernel-tuning aid: digest of the model state after a few steps with the production library selected by ROMS_B200_LIB
(tools/variant.py).  A variant that only changes how operands are fetched must print the digest of the default library.

    ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_X.so python tools/lib_digest.py [Lm Mm N steps]
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roms_trunk_mgh_b200 import synth  # noqa: E402
from roms_trunk_mgh_b200.ocean import field_names  # noqa: E402

a = [int(x) for x in sys.argv[1:5]] + [96, 40, 30, 6][len(sys.argv) - 1:]
t = synth.make_tile(synth.APP_BENCHMARK, a[0], a[1], a[2])
t.main3d(a[3])
n2, n3 = field_names(t.NT)
h = hashlib.sha256()
for n in n2 + n3:
    h.update(np.ascontiguousarray(t.get(n)).tobytes())
print("DIGEST", os.path.basename(os.environ.get("ROMS_B200_LIB", "default")), h.hexdigest()[:16], "maxspeed", t.diag()["max_speed"])
