"""Run by test_gpu_parity.py in a subprocess: steps a strict-build tile and the oracle side by side with the given
roms_b200_set_option switches and demands bit-exact fields; prints a digest of the state.

    python tests/gpu_variant_check.py benchmark30 6 cuda_graphs=0
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
    opts = dict(a.split("=") for a in sys.argv[3:])
    for k, v in opts.items():
        t.set_option(k, float(v))
    o.step(nsteps)
    t.main3d(nsteps)                      # steps >= 3 replay the captured CUDA graph unless cuda_graphs=0
    names = all_names(int(o.opt("NT")))
    bad = compare(o, t, names, exact=True)
    h = hashlib.sha256()
    for n in names:
        h.update(np.ascontiguousarray(t.get(n)).tobytes())
    print("DIGEST", h.hexdigest())
    if bad:
        print("MISMATCH", bad[:4])
        sys.exit(1)
    print("VARIANT_OK", case, nsteps, opts)


if __name__ == "__main__":
    main()
