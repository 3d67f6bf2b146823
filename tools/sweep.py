"""Kernel-tuning aid: run bench.py once per variant library (tools/variant.py) and print the chosen phase times.

    python tools/sweep.py step3d_uv,step3d_t base a b c
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
phases = sys.argv[1].split(",")
for name in sys.argv[2:]:
    env = dict(os.environ)
    if name != "base":
        env["ROMS_B200_LIB"] = os.path.join(ROOT, "roms_trunk_mgh_b200", "lib", "var", f"libroms_b200_{name}.so")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu", "--no-extras"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(name, f"step {d['ms_per_step']:.3f} ms |", " ".join(f"{k}={d['phase_ms'][k]:.3f}" for k in phases), flush=True)
    except Exception as e:  # noqa: BLE001
        print(name, "FAILED", e, r.stderr[-400:], flush=True)
