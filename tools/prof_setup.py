#!/usr/bin/env python
"""Workload for profiling the multigrid set-up kernels and the preconditioned stencil paths at config-5 size
(2048^2 fine lattice, 4x4 blocks, 8 null vectors; 512^2 x 8 coarse level): one device set-up with short smoothing
solves, then the composite operators and the prepare / reconstruct passes on both levels.

    ncu --set full -k regex:"mg_|coarse_part|coarse_sign" ... python tools/prof_setup.py [L]
Also prints CUDA-event timings of every step (not under a profiler: run it plainly for those)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import mg_setup  # noqa: E402
from __graft_entry__ import _load_pkg  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    block, nc = 4, 8
    glb = _load_pkg()
    ctx = glb.Context(device=0)
    rows = list(range(L))
    U = bench.gauge_rows(L, rows)
    mass = 0.1
    cl0, hp0, _ = mg_setup.staggered_stencil(U, L, L, 0.0)
    fine = ctx.stencil2d(cl0, hp0, None, L, L, 1, shift=mass)
    mg = ctx.multigrid_setup(fine, L, L, [block], [nc], seed=1337, max_iter=20)
    print(json.dumps(dict(kind="setup", L=L, **mg.setup_seconds())), flush=True)
    Lc = L // block
    clc, hpc, shc = mg.level_stencil(1)
    coarse = ctx.stencil2d(clc, hpc, None, Lc, Lc, nc, shift=shc[0])

    def timed(label, fn, nbytes, reps=20):
        fn()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.sync()
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps(dict(kind="pass", what=label, ms=dt * 1e3, algorithmic_GB=nbytes / 1e9, GBps=nbytes / dt / 1e9)),
              flush=True)

    for name, op, X, n_c in (("fine nc=1 2048^2", fine, L, 1), ("coarse nc=8 512^2", coarse, Lc, nc)):
        n = X * X * n_c
        rg = np.random.default_rng(1)
        a, b, out = ctx.vector(n).upload(rg.standard_normal(n) + 0j), ctx.vector(n).upload(rg.standard_normal(n) + 0j), ctx.vector(n)
        V = X * X
        vec = n_c * 16 * V
        tb = 0 if n_c == 1 else 1
        # matrix bytes ONE partial apply reads: e/o -- the 4 hopping rows of the live half of the sites; t/b -- half of
        # the columns of the clover + 4 hopping rows of the live half of the colours
        mat = (4 * n_c * n_c * 16 * V / 2) if tb == 0 else (5 * n_c * n_c * 16 * V / 4)
        # one partial apply touches half of the output rows: half of the matrices, the whole input (neighbours), the
        # whole output (the dead half is zeroed / copied)
        timed(name + ": prec_prepare (one pass)", lambda: op.prec_prepare(tb, out, a), mat + 3 * vec)
        timed(name + ": prec_reconstruct (one pass)", lambda: op.prec_reconstruct(tb, out, a, b), mat + 4 * vec)
        for view, passes in (("M2MDEODOE" if tb == 0 else "M2MDTBDBT", 2), ("NORMAL_EO" if tb == 0 else "NORMAL_TB", 4),
                             ("DAGGER_EO" if tb == 0 else "DAGGER_TB", 1)):
            v = op.view(view)
            full = (5 * n_c * n_c * 16 * V + 2 * vec) if view.startswith("DAGGER") else None
            nbytes = (full + 4 * vec) if full else passes * (mat + 2 * vec) + vec
            timed(name + ": " + view, lambda: v.apply(out, a), nbytes)
            v.destroy()


if __name__ == "__main__":
    main()
