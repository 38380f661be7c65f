#!/usr/bin/env python
"""One tuning sample under the environment variables of the caller (kernels read GLB_* once per process):
   python tools/tune_variant.py normal [L]   -> D^dag D apply ms + CGNE solve ms at L^2
   python tools/tune_variant.py coarse       -> apply_stencil_2d nc=8 (512^2), two-link, nc=16 (256^2)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from __graft_entry__ import _load_pkg  # noqa: E402
import torch  # noqa: E402

glb = _load_pkg()
ctx = glb.Context(device=0)
stream = torch.cuda.ExternalStream(ctx.stream())
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("GLB_"))


def time_loop(fn, reps):
    for _ in range(3):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


what = sys.argv[1]
if what == "normal":
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    rows = [L - 2, L - 1] + list(range(L)) + [0, 1]
    U = bench.gauge_rows(L, rows)
    b = bench.rhs_rows(L, rows[2:-2])
    V = L * L
    opN = ctx.staggered_local(U, L, L, 0.1, glb.STAG_NORMAL)
    opDd = ctx.staggered_local(U, L, L, 0.1, glb.STAG_DAGGER)
    x, y, bp = ctx.vector(V).upload(b), ctx.vector(V), ctx.vector(V)
    opDd.apply(bp, x)
    ms_apply = time_loop(lambda: opN.apply(y, x), 40)
    sol = ctx.vector(V)

    def solve():
        sol.zero()
        return ctx.solve("CG", opN, sol, bp, max_iter=5000, eps=1e-10)
    for _ in range(2):
        info = solve()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        info = solve()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    ms_solve = e0.elapsed_time(e1) / 5
    print("[%s] L=%d  DdagD apply %.4f ms (%.0f GB/s)  CGNE solve %.2f ms  %d it  %.4f ms/it (%.0f GB/s moved at 192 B/site)"
          % (tag, L, ms_apply, 64 * V / ms_apply / 1e6, ms_solve, info["iter"], ms_solve / info["iter"],
             192 * V / (ms_solve / info["iter"]) / 1e6), flush=True)
elif what == "coarse":
    rg = np.random.default_rng(3)
    for (X, nc, two) in ((512, 8, False), (512, 8, True), (256, 16, False), (2048, 1, False), (512, 4, False)):
        V = X * X
        rc = lambda n: (rg.standard_normal(n) + 1j * rg.standard_normal(n))
        cl, hp = rc(V * nc * nc), rc(4 * V * nc * nc)
        tl = rc(8 * V * nc * nc) if two else None
        op = ctx.stencil2d(cl, hp, tl, X, X, nc, shift=0.1)
        xv, yv = ctx.vector(V * nc).upload(rc(V * nc)), ctx.vector(V * nc)
        ms = time_loop(lambda: op.apply(yv, xv), 30)
        bps = ((13 if two else 5) * nc * nc + 2 * nc) * 16
        print("[%s] stencil X=%d nc=%d two=%d  %.4f ms  %.0f GB/s (%.1f %% of 6547)"
              % (tag, X, nc, two, ms, bps * V / ms / 1e6, bps * V / ms / 1e6 / 65.472), flush=True)
        op.destroy()
        del cl, hp, tl
