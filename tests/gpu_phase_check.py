"""Developer tool (run on a GPU box): walk one baroclinic step phase by phase on the oracle and on the device and report,
per phase, which fields differ.  `python tests/gpu_phase_check.py [strict|prod] [app ...]`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc  # noqa: E402
from helpers import all_names, compare, copy_state, make_pair  # noqa: E402

STEP_PHASES = ["set_massflux", "rho_eos", "set_vbc", "ana_vmix", "omega", "wvelocity", "set_zeta", "pre_step3d", "prsgrd", "t3dmix",
               "rhs3d", "uv3dmix", "step2d_loop", "set_depth", "step3d_uv", "omega2", "step3d_t"]


def begin_step(o, t):
    d = o.indices()
    d["nstp"] = 1 + ((d["iic"] - d["ntstart"]) % 2)
    d["nnew"] = 3 - d["nstp"]
    d["nrhs"] = d["nstp"]
    d["tdays"] = d["time"] / 86400.0
    o.set_indices(d)
    o.run_phase("set_data")
    t.set("sustr", o.field("sustr")); t.set("svstr", o.field("svstr"))
    t.set_indices(o.indices())


def end_step(o, t):
    for m in (o, t):
        d = m.indices()
        d["iic"] += 1
        d["time"] += o.opt("dt")
        m.set_indices(d)


def walk(app, strict, spinup, **kw):
    o, t = make_pair(app, strict=strict, spinup=spinup, **kw)
    names = all_names(int(o.opt("NT")))
    exact = strict
    nbad = 0
    begin_step(o, t)
    for ph in STEP_PHASES:
        o.run_phase(ph)
        t.run_phase(ph)
        if ph == "step2d_loop":
            assert o.indices() == t.indices(), (o.indices(), t.indices())
        bad = compare(o, t, names, exact=exact, rtol=1e-11)
        if bad:
            nbad += 1
            print(f"  [{ph}] MISMATCH:", [(n, f"{d:.3e}", f"{s:.3e}", c) for n, d, s, c in bad])
            copy_state(o, t)
        else:
            print(f"  [{ph}] ok")
    end_step(o, t)
    return nbad


def multi(app, strict, nsteps, **kw):
    o, t = make_pair(app, strict=strict, spinup=0, **kw)
    names = all_names(int(o.opt("NT")))
    worst = 0.0
    for s in range(nsteps):
        o.step(1)
        t.set("sustr", o.field("sustr")); t.set("svstr", o.field("svstr"))
        t.main3d(1)
    bad = compare(o, t, names, exact=strict, rtol=1e-10)
    for n in ["zeta1", "u1", "v1", "t1_0"]:
        a = o.field(n); b = t.get(n)
        worst = max(worst, float(np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-300)))
    print(f"  multi-step({nsteps}) mismatches: {[(n, f'{d:.3e}', f'{s:.3e}', c) for n, d, s, c in bad]}  worst rel = {worst:.3e}")
    print("  diag oracle:", {k: f"{v:.10e}" for k, v in o.diag().items()})
    print("  diag device:", {k: f"{v:.10e}" for k, v in t.diag().items()})
    return len(bad)


if __name__ == "__main__":
    strict = (len(sys.argv) < 2) or sys.argv[1] == "strict"
    apps = sys.argv[2:] or ["seamount", "benchmark", "upwelling"]
    table = {"upwelling": (orc.APP_UPWELLING, {}), "seamount": (orc.APP_SEAMOUNT, {}),
             "benchmark": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10)),
             "benchmark30": (orc.APP_BENCHMARK, dict(Lm=96, Mm=40, N=30)),
             "benchmark_geo": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, mix_geo_ts=1)),
             "benchmark_p31": (orc.APP_BENCHMARK, dict(Lm=64, Mm=32, N=10, dj_gradps=0, nonlin_eos=0))}
    total = 0
    for a in apps:
        app, kw = table[a]
        for spin in (0, 1, 3):
            print(f"== {a} strict={strict} spinup={spin}")
            total += walk(app, strict, spin, **kw)
        print(f"== {a} multi-step")
        total += multi(app, strict, 6, **kw)
    print("TOTAL MISMATCHING PHASES:", total)
    sys.exit(1 if total else 0)
