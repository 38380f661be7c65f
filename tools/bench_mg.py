#!/usr/bin/env python
"""Config 5 end to end (one GPU): the aa_mg outer solve -- minv_vector_gcr_var_precond_restart(64) preconditioned
by the two-level V cycle (GCR smoother 6+6, coarse GCR(64) to 1e-2; multigrid/aa_mg/input_params.cpp:621-800) -- on
an L x L staggered lattice (default 2048, mass 1e-2, blocksize 4, 4 null vectors x even/odd = 8 coarse colours),
next to the unpreconditioned solvers at the same mass.  One JSON object per line.

    python tools/bench_mg.py [L] [mass]

The hierarchy is set up with tools/mg_setup.py: null vectors by the DEVICE BiCGStab, block orthonormalisation and
the Galerkin coarse stencil in numpy on the host (set-up on the device is SURVEY 8f-2, not built yet; its time is
reported separately and is not part of the solve)."""
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
    mass = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
    block, nraw = 4, 4
    glb = _load_pkg()
    ctx = glb.Context(device=0)
    V = L * L

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    rows = list(range(L))
    U = bench.gauge_rows(L, rows)
    bh = bench.rhs_rows(L, rows)
    D = ctx.staggered(U, L, L, mass, 0)

    # ---- set-up (host + device solver; reported, not part of the solve)
    t0 = time.perf_counter()
    raw = mg_setup.null_vectors_device(ctx, D, V, nraw)
    t_null = time.perf_counter() - t0
    t0 = time.perf_counter()
    vecs = mg_setup.block_orthonormalize(mg_setup.split_even_odd(raw, L, L), L, L, block, block)
    cl0, hp0, sh0 = mg_setup.staggered_stencil(U, L, L, mass)
    clc, hpc = mg_setup.coarse_stencil(vecs, hp0, sh0, L, L, block, block)
    t_host = time.perf_counter() - t0
    nc = len(vecs)
    Lc = L // block
    fine = ctx.stencil2d(cl0, hp0, None, L, L, 1, shift=sh0)
    coarse = ctx.stencil2d(clc, hpc, None, Lc, Lc, nc)
    tr = ctx.mg_transfer(L, L, 1, block, block, vecs)
    emit(kind="mg_setup", L=L, mass=mass, null_vectors=nc, coarse_lattice=[Lc, Lc, nc],
         seconds_null_vectors_device_bicgstab=t_null, seconds_host_numpy_orthonormalize_and_galerkin=t_host)
    del raw, vecs, clc, hpc

    b = ctx.vector(V).upload(bh)
    x = ctx.vector(V)
    bnorm = float(np.sqrt(ctx.norm2sq(b)))
    chk = ctx.vector(V)

    def true_rel(xv):
        D.apply(chk, xv)
        return float(np.sqrt(ctx.diffnorm2sq(chk, b))) / bnorm

    # ---- the outer solve of config 5 on the stencil operator (what the reference's driver applies) and on the
    # native staggered kernel as the fine operator
    for label, fine_op in (("fine level = nc=1 stencil (as the reference)", fine), ("fine level = native staggered kernel", D)):
        mg = ctx.multigrid([fine_op, coarse], [tr])
        mg.set()   # GCR smoother 6+6, coarse GCR(64) to 1e-2, V cycle
        for rep in range(2):
            x.zero()
            ctx.sync()
            l0 = ctx.launches()
            t0 = time.perf_counter()
            info = mg.vpgcr(x, b, max_iter=100000, eps=5e-7, restart_freq=64)
            dt = time.perf_counter() - t0
        emit(kind="solve", L=L, mass=mass, solver="config 5: VPGCR(64) + two-level V cycle, tol 5e-7; " + label,
             seconds=dt, iterations=info["iter"], outer_ops=info["ops_count"], success=info["success"],
             true_rel_residual=true_rel(x), kernel_launches=ctx.launches() - l0, dslash_counts=mg.counts())
        mg.destroy()

    # ---- the same system without the preconditioner
    for name, solver, kw in (("minv_vector_gcr_restart(64) on D", "GCR_RESTART", dict(restart_freq=64)),
                             ("minv_vector_bicgstab on D", "BICGSTAB", dict())):
        for rep in range(2):
            x.zero()
            ctx.sync()
            t0 = time.perf_counter()
            info = ctx.solve(solver, D, x, b, max_iter=200000, eps=5e-7, **kw)
            dt = time.perf_counter() - t0
        emit(kind="solve", L=L, mass=mass, solver="unpreconditioned " + name + ", tol 5e-7", seconds=dt,
             iterations=info["iter"], ops=info["ops_count"], success=info["success"], true_rel_residual=true_rel(x))
    N = ctx.staggered(U, L, L, mass, glb.STAG_NORMAL)
    Dd = ctx.staggered(U, L, L, mass, glb.STAG_DAGGER)
    bp = ctx.vector(V)
    Dd.apply(bp, b)
    for rep in range(2):
        x.zero()
        ctx.sync()
        t0 = time.perf_counter()
        info = ctx.solve("CG", N, x, bp, max_iter=200000, eps=5e-7)
        dt = time.perf_counter() - t0
    emit(kind="solve", L=L, mass=mass, solver="unpreconditioned CGNE (minv_vector_cg on D^dag D), tol 5e-7 on the normal system",
         seconds=dt, iterations=info["iter"], ops=info["ops_count"], success=info["success"], true_rel_residual=true_rel(x))


if __name__ == "__main__":
    main()
