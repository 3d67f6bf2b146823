"""Run by test_gpu_parity.py in a subprocess with kernel-selection environment variables set (they are read once per
process): steps a strict-build tile and the oracle side by side and demands bit-exact fields; prints a digest of the state.

    ROMS_B200_STEP2D=march python tests/gpu_variant_check.py benchmark30 6
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc  # noqa: E402
from helpers import all_names, compare, make_pair  # noqa: E402

CASES = {"seamount": (orc.APP_SEAMOUNT, {}), "benchmark30": (orc.APP_BENCHMARK, dict(Lm=96, Mm=40, N=30)),
         "ragged": (orc.APP_BENCHMARK, dict(Lm=37, Mm=19, N=7)), "wide": (orc.APP_BENCHMARK, dict(Lm=256, Mm=24, N=6))}


def main():
    case, nsteps = sys.argv[1], int(sys.argv[2])
    app, kw = CASES[case]
    o, t = make_pair(app, strict=True, **kw)
    o.step(nsteps)
    t.main3d(nsteps)                      # steps >= 3 replay the captured CUDA graph unless ROMS_B200_NO_GRAPH=1
    names = all_names(int(o.opt("NT")))
    bad = compare(o, t, names, exact=True)
    h = hashlib.sha256()
    for n in names:
        h.update(np.ascontiguousarray(t.get(n)).tobytes())
    print("DIGEST", h.hexdigest())
    if bad:
        print("MISMATCH", bad[:4])
        sys.exit(1)
    print("VARIANT_OK", case, nsteps, {k: os.environ[k] for k in os.environ if k.startswith("ROMS_B200_")})


if __name__ == "__main__":
    main()
