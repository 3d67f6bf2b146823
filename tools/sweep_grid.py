"""Kernel-tuning aid: like tools/sweep.py, on a chosen grid.   python tools/sweep_grid.py b3tile8 rhs3d,uv3dmix base a b"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
grid, phases = sys.argv[1], sys.argv[2].split(",")
for name in sys.argv[3:]:
    env = dict(os.environ)
    if name != "base":
        env["ROMS_B200_LIB"] = os.path.join(ROOT, "roms_trunk_mgh_b200", "lib", "var", f"libroms_b200_{name}.so")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--grid", grid, "--steps", "20", "--warmup", "3", "--no-cpu", "--no-extras"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(name, f"step {d['ms_per_step']:.4f} ms |", " ".join(f"{k}={d['phase_ms'][k]*1e3:.1f}us" for k in phases), flush=True)
    except Exception as e:  # noqa: BLE001
        print(name, "FAILED", e, r.stderr[-400:], flush=True)
