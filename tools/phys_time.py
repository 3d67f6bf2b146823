#!/usr/bin/env python
"""Tuning aid: per-phase device times of the full BENCHMARK cpp set (synth.FULL_BENCHMARK) on one GPU.
usage: python tools/phys_time.py [grid]   (grid: benchmark3 | benchmark1 | b3tile8)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roms_trunk_mgh_b200 import synth  # noqa: E402

GR = {"benchmark1": (512, 64, 30), "benchmark3": (2048, 256, 30), "b3tile8": (256, 256, 30)}
g = GR[sys.argv[1] if len(sys.argv) > 1 else "benchmark3"]
t = synth.make_tile(synth.APP_BENCHMARK, *g, **synth.FULL_BENCHMARK)
t.main3d(10)
t.main3d(6); ms = t.last_step_ms() / 6
t.profile(2); t.main3d(2); t.profile(2); t.main3d(4)
pf, _ = t.profile_get()
print(json.dumps({"ms_per_step": ms, "phase_ms": {k: round(v / 4, 4) for k, v in sorted(pf.items(), key=lambda x: -x[1])}}))
