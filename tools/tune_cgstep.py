#!/usr/bin/env python
"""Kernel-shape sweep of the single-kernel CG iteration (csrc/cgstep.cu) on one GPU:

    python tools/tune_cgstep.py [L] [variant ...]

For every variant (1000*rows per item [0 = static partition] + 100*consumer warps + 10*stages + blocks per SM) one CGNE solve at L^2 is timed per launch with
CUDA events (glb_prof_*); prints one JSON object per variant, and the two-kernel loop for comparison."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg  # noqa: E402

glb = _load_pkg()
ctx = glb.Context(device=0)
# first argument: L (square lattice) or XxY
_a = sys.argv[1] if len(sys.argv) > 1 else "4096"
X, Y = (int(_a.split("x")[0]), int(_a.split("x")[1])) if "x" in _a else (int(_a), int(_a))
L = X
variants = [int(v) for v in sys.argv[2:]] or [433, 508433, 32433]
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
links, b_h = ctx.synthetic_inputs(X, Y)
N = ctx.staggered(links, X, Y, 0.1, glb.STAG_NORMAL)
Dd = ctx.staggered(links, X, Y, 0.1, glb.STAG_DAGGER)
V = X * Y
b = ctx.vector(V).upload(b_h)
bp = ctx.vector(V)
Dd.apply(bp, b)
x = ctx.vector(V)


def one(label):
    for _ in range(2):
        x.zero()
        ctx.solve("CG", N, x, bp, max_iter=5000, eps=1e-10)
    ctx.prof_enable(True)
    x.zero()
    info = ctx.solve("CG", N, x, bp, max_iter=5000, eps=1e-10)
    ctx.prof_enable(False)
    it = info["iter"]
    ts = ctx.prof_read(7)
    rec = {"variant": label, "lattice": [X, Y], "iterations": it}
    if len(ts) == 1:     # persistent kernel: one launch ran all it + 1 steps
        ms = ts[0] / (it + 1)
        rec.update(launches=1, cg_step_ms=ms, GBps=160.0 * V / ms / 1e6, frac=160.0 * V / ms / 1e6 / PEAK,
                   ms_per_iteration=ms)
    elif ts:
        ts = ts[:it + 1]
        ms = float(np.mean(ts[2:]))
        rec.update(launches=len(ts), cg_step_ms=ms, GBps=160.0 * V / ms / 1e6, frac=160.0 * V / ms / 1e6 / PEAK,
                   ms_min=float(np.min(ts[2:])), ms_max=float(np.max(ts[2:])), ms_per_iteration=ms)
    else:
        f, u = ctx.prof_read(1)[:it - 1], ctx.prof_read(3)[:it]
        rec.update(fused_ms=float(np.mean(f)), update_ms=float(np.mean(u)), ms_per_iteration=float(np.mean(f) + np.mean(u)))
    print(json.dumps(rec), flush=True)


for v in variants:
    ctx.cg_step_mode(True, v)
    try:
        one(v)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"variant": v, "error": str(e)}), flush=True)
ctx.cg_step_mode(False)
one("two-kernel loop")
