#!/usr/bin/env python
"""Config 5 end to end (one GPU): the aa_mg outer solve -- minv_vector_gcr_var_precond_restart(64) preconditioned
by the two-level V cycle (GCR smoother 6+6, coarse GCR(64) to 1e-2; multigrid/aa_mg/input_params.cpp:621-800) -- on
an L x L staggered lattice (default 2048, mass 1e-2, blocksize 4, 4 null vectors x even/odd = 8 coarse colours),
next to the unpreconditioned solvers at the same mass.  One JSON object per line.

    python tools/bench_mg.py [L] [mass] [--numpy-setup]

The hierarchy is set up ON THE DEVICE (glbx_mg_setup = null_generate_random_smooth_dev with the
driver's defaults -- BiCGStab to 5e-5, at most 500 iterations, null mass 1e-2, BLOCK_EO --, block_orthonormalize_dev,
generate_coarse_from_fine_stencil_dev); its wall-clock time is reported per phase.  --numpy-setup additionally times
the host restatement of the same steps (tools/mg_setup.py: one CPU core, vectorised numpy) on the device's own raw
null vectors for comparison."""
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
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    L = int(pos[0]) if len(pos) > 0 else 2048
    mass = float(pos[1]) if len(pos) > 1 else 0.01
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

    # ---- set-up on the device (reported, not part of the solve)
    cl0, hp0, _ = mg_setup.staggered_stencil(U, L, L, 0.0)        # get_square_staggered_u1_stencil, mass in the shift
    fine = ctx.stencil2d(cl0, hp0, None, L, L, 1, shift=mass)
    nc = 2 * nraw
    Lc = L // block
    for rep in range(2):                                          # second run: memory pool and kernels warm
        ctx.sync()
        t0 = time.perf_counter()
        mgs = ctx.multigrid_setup(fine, L, L, [block], [nc], seed=1337)
        ctx.sync()
        t_dev = time.perf_counter() - t0
        secs = mgs.setup_seconds()
        emit(kind="mg_setup", where="device (glbx_mg_setup)", run=rep, L=L, mass=mass, null_vectors=nc,
             coarse_lattice=[Lc, Lc, nc], seconds_total=t_dev, seconds_null_vectors=secs["null_vectors"],
             seconds_block_orthonormalize=secs["block_orthonormalize"], seconds_transfer_and_galerkin=secs["galerkin"],
             null_vector_applies=mgs.counts()["nullvectors"][0])
        if rep == 0:
            mgs.destroy()
    if "--numpy-setup" in sys.argv:
        raw = [mgs.null_vector(0, v) + mgs.null_vector(0, v + nraw) for v in range(nraw)]   # stand-ins of the same shape
        t0 = time.perf_counter()
        vecs = mg_setup.block_orthonormalize(mg_setup.split_even_odd(raw, L, L), L, L, block, block)
        clc, hpc = mg_setup.coarse_stencil(vecs, hp0, mass, L, L, block, block)
        emit(kind="mg_setup", where="host numpy restatement (one core): block orthonormalisation + Galerkin product only",
             L=L, seconds=time.perf_counter() - t0)
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
    mgs.set()   # GCR smoother 6+6, coarse GCR(64) to 1e-2, V cycle
    for rep in range(2):
        x.zero()
        ctx.sync()
        l0 = ctx.launches()
        t0 = time.perf_counter()
        info = mgs.vpgcr(x, b, max_iter=100000, eps=5e-7, restart_freq=64)
        dt = time.perf_counter() - t0
    emit(kind="solve", L=L, mass=mass, solver="config 5: VPGCR(64) + two-level V cycle, tol 5e-7; hierarchy set up on the "
         "device; fine level = nc=1 stencil (as the reference)", seconds=dt, iterations=info["iter"],
         outer_ops=info["ops_count"], success=info["success"], true_rel_residual=true_rel(x),
         kernel_launches=ctx.launches() - l0, dslash_counts=mgs.counts())
    # ---- where the solve spends its time: the same solve once more with a CUDA-event pair around every classified
    # launch (glb_prof_*); GB/s on the algorithmic bytes of each launch, share of the wall-clock time of THIS solve
    x.zero()
    ctx.sync()
    ctx.prof_enable(True)
    t0 = time.perf_counter()
    info_p = mgs.vpgcr(x, b, max_iter=100000, eps=5e-7, restart_freq=64)
    ctx.sync()
    dt_p = time.perf_counter() - t0
    summ = ctx.prof_summary()
    ctx.prof_enable(False)
    peak = 6543.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    table, covered = [], 0.0
    for name, (n, ms, by) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
        covered += ms
        table.append({"kernel": name, "launches": n, "ms_total": round(ms, 3), "us_per_launch": round(1e3 * ms / n, 2),
                      "GBps": round(by / ms / 1e6, 1) if ms > 0 and by > 0 else None,
                      "frac_of_hbm_peak": round(by / ms / 1e6 / peak, 3) if ms > 0 and by > 0 else None,
                      "share_of_solve": round(ms / (1e3 * dt_p), 3)})
    emit(kind="solve_profile", L=L, mass=mass, seconds_with_events=dt_p, iterations=info_p["iter"],
         kernels=table, share_in_classified_kernels=round(covered / (1e3 * dt_p), 3),
         note="rest = stream synchronisations of the host-scalar shells, event overhead, unclassified small kernels "
              "(copies, memsets, vector subtraction inside glb_sub counts as BLAS-1)")
    mgs.destroy()

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
