"""Regenerates the fixtures in tests/golden/.

Two kinds of numbers live here:
  * reference-quoted known answers (the only ones the reference tree holds for this path):
      - EOS check values, ROMS/Nonlinear/rho_eos.F:21-29 (T=3 C, S=35.5, Z=-5000 m)
      - set_weights integrals "values must be 1, 1, approx 1/2, 1, 1", ROMS/Utility/set_weights.F FORMAT 40
    These are typed in by hand from the cited lines, never computed.
  * oracle-generated regression vectors (marked "oracle_generated"): they pin the C++ restatement against accidental
    edits; they are NOT reference outputs (the Fortran reference cannot be built here).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

kat = {
    "rho_eos_check_values": {"source": "ROMS/Nonlinear/rho_eos.F:21-29", "T": 3.0, "S": 35.5, "Z": -5000.0,
                             "den": 1050.3639165364, "den1": 1028.2845117925, "bulk": 23786.056026320},
    "set_weights_statement": {"source": "ROMS/Utility/set_weights.F FORMAT 40", "values": [1, 1, 0.5, 1, 1]},
}
_, chk, _, _ = orc.set_weights(30)
kat["set_weights_ndtfast30_integrals"] = [round(c, 12) for c in chk]
kat["set_weights_ndtfast30_integrals_note"] = "oracle_generated (12 decimals, the precision FORMAT 40 prints)"
o = orc.Oracle(orc.APP_UPWELLING)
o.run_phase("set_data"); o.run_phase("ini")
d = o.diag()
kat["upwelling_step0"] = {"note": "oracle_generated", "avgpe": d["avgpe"], "volume": d["volume"]}
json.dump(kat, open(os.path.join(HERE, "reference_kat.json"), "w"), indent=1)
o = orc.Oracle(orc.APP_UPWELLING)
o.step(10)
np.savez_compressed(os.path.join(HERE, "upwelling_10steps.npz"), **{n: o.field(n).copy() for n in ("zeta1", "u1", "v1", "t1_0")})
# SEAMOUNT (A4/A4 tracers, MIX_GEO_TS, QDRAG, no-slip) and a small BENCHMARK (nonlinear EOS, CURVGRID, U3/C4, MIX_S_TS): the
# state the GPU parity tests must reproduce (tests/test_gpu_parity.py::test_cuda_path_against_committed_golden_vectors)
GOLD_CASES = {"seamount_6steps": (orc.APP_SEAMOUNT, {}, 6), "benchmark_64x32x10_6steps": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10), 6)}
for name, (app, kw, nsteps) in GOLD_CASES.items():
    o = orc.Oracle(app, **kw)
    o.run_phase("set_data"); o.run_phase("ini")
    o.step(nsteps)
    NT = int(o.opt("NT"))
    names = ["zeta1", "ubar1", "vbar1", "u1", "v1", "rho", "W"] + [f"t1_{it}" for it in range(NT)]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{n: o.field(n).copy() for n in names})
# the shipped BENCHMARK cpp set minus MIX_GEO_TS (bulk_flux + lmd_vmix + the in-chain terms) on a small channel: 16 steps, so that
# the surface boundary layer has started to deepen (ksbl takes two values)
FULL = dict(bv_frequency=1, eos_tderivative=1, solar_source=1, lmd_nonlocal=1, bulk_fluxes=1, lmd_mixing=1)
o = orc.Oracle(orc.APP_BENCHMARK, Lm=32, Mm=24, N=30, **FULL)
o.run_phase("set_data"); o.run_phase("ini")
o.step(16)
names = ["zeta1", "ubar1", "u1", "t1_0", "Akv", "Akt_1", "hsbl", "ksbl", "ghats_0", "sustr", "stflux_0", "lhflx", "shflx", "lrflx"]
np.savez_compressed(os.path.join(HERE, "benchmark_fullphysics_32x24x30_16steps.npz"), **{n: o.field(n).copy() for n in names})
print("golden fixtures written")
