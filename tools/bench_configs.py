#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configurations (one GPU): lattice sweep of the
stencil applies, every solver of the path at 4096^2, the coarse-stencil apply of config 5 and the
latency-bound config 1.  Prints one JSON object per line; `python tools/bench_configs.py > out.jsonl`.

Algorithmic bytes follow SURVEY section 8 (d-bytes).  Timing: CUDA events on the library's stream for
pure kernel loops, host wall clock around complete solver calls (they synchronise internally).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from __graft_entry__ import _load_pkg  # noqa: E402

PEAK = 6547.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    import torch
    glb = _load_pkg()
    ctx = glb.Context(device=0)
    stream = torch.cuda.ExternalStream(ctx.stream())
    quick = "--quick" in sys.argv

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

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    # ---- 1. lattice sweep of the stencil applies
    for L in ([256, 1024, 4096] if quick else [256, 512, 1024, 2048, 4096, 8192]):
        rows = [L - 2, L - 1] + list(range(L)) + [0, 1]
        U = bench.gauge_rows(L, rows)
        b = bench.rhs_rows(L, rows[2:-2])
        x, y = ctx.vector(L * L).upload(b), ctx.vector(L * L)
        for flags, name, bps in ((0, "D (square_staggered_u1)", 64), (glb.STAG_DAGGER, "D^dag", 64),
                                 (glb.STAG_GAMMA5, "gamma5 D", 64), (glb.STAG_NORMAL, "D^dag D one pass", 64)):
            op = ctx.staggered_local(U, L, L, 0.1, flags)
            ms = time_loop(lambda: op.apply(y, x), 200 if L <= 1024 else 40)
            emit(kind="apply", L=L, op=name, ms=ms, bytes_per_site=bps, GBps=bps * L * L / ms / 1e6,
                 frac_of_measured_hbm_peak=bps * L * L / ms / 1e6 / PEAK,
                 note="working set fits L2 (latency/launch-bound)" if L <= 1024 else "")
            op.destroy()
        op = ctx.laplace_u1(U[4 * L:4 * L + 2 * L * L], L, L, 0.1)
        ms = time_loop(lambda: op.apply(y, x), 40)
        emit(kind="apply", L=L, op="square_laplace_u1", ms=ms, bytes_per_site=64, GBps=64 * L * L / ms / 1e6,
             frac_of_measured_hbm_peak=64 * L * L / ms / 1e6 / PEAK)
        op.destroy()
        op = ctx.laplace(L, L, 1, 4.01 + 1j, np.complex128)
        ms = time_loop(lambda: op.apply(y, x), 40)
        emit(kind="apply", L=L, op="square_laplacian (imag_laplace.cpp)", ms=ms, bytes_per_site=32,
             GBps=32 * L * L / ms / 1e6, frac_of_measured_hbm_peak=32 * L * L / ms / 1e6 / PEAK)
        op.destroy()
        del x, y

    # ---- 2. every solver of the path at 4096^2 (config 3) and 1024^2
    for L in ([1024] if quick else [1024, 4096]):
        V = L * L
        rows = [L - 2, L - 1] + list(range(L)) + [0, 1]
        U = bench.gauge_rows(L, rows)
        b_h = bench.rhs_rows(L, rows[2:-2])
        D = ctx.staggered_local(U, L, L, 0.1, 0)
        N = ctx.staggered_local(U, L, L, 0.1, glb.STAG_NORMAL)
        Dd = ctx.staggered_local(U, L, L, 0.1, glb.STAG_DAGGER)
        b = ctx.vector(V).upload(b_h)
        bp = ctx.vector(V)
        Dd.apply(bp, b)
        x = ctx.vector(V)
        # bytes per iteration, SURVEY 8 d-bytes (fused minimum)
        cases = [("CGNE: minv_vector_cg on D^dag D", "CG", N, bp, dict(eps=1e-10), lambda it: 272.0 * it),
                 ("minv_vector_cr on D^dag D", "CR", N, bp, dict(eps=1e-10), lambda it: (272.0 + 64.0) * it),
                 ("minv_vector_bicgstab on D", "BICGSTAB", D, b, dict(eps=1e-10), lambda it: 336.0 * it),
                 ("minv_vector_bicgstab_l(l=4) on D", "BICGSTAB_L", D, b, dict(eps=1e-10, l=4), None),
                 ("minv_vector_gcr_restart(20) on D", "GCR_RESTART", D, b, dict(eps=1e-8, restart_freq=20),
                  lambda it: (288.0 + 48.0 * 9.5) * it),
                 ("minv_vector_gmres_restart(20) on D", "GMRES_RESTART", D, b, dict(eps=1e-8, restart_freq=20),
                  lambda it: 1250.0 * it)]
        for name, solver, op, rhs, kw, bytes_fn in cases:
            x.zero()
            ctx.solve(solver, op, x, rhs, max_iter=100000, **kw)  # warm-up (pool, caches)
            x.zero()
            ctx.sync()
            t0 = time.perf_counter()
            info = ctx.solve(solver, op, x, rhs, max_iter=100000, **kw)
            dt = time.perf_counter() - t0
            rec = dict(kind="solve", L=L, solver=name, seconds=dt, iterations=info["iter"], ops=info["ops_count"],
                       success=info["success"], iterations_per_s=info["iter"] / dt,
                       true_rel_residual=float(np.sqrt(info["resSq"]) / np.sqrt(ctx.norm2sq(rhs))))
            if bytes_fn:
                gb = bytes_fn(info["iter"]) * V / 1e9
                rec.update(algorithmic_GB=gb, GBps=gb / dt, frac_of_measured_hbm_peak=gb / dt / PEAK)
            emit(**rec)
        shifts = [0.0, 0.01, 0.05, 0.25]
        xs = [ctx.vector(V) for _ in shifts]
        ctx.solve_cg_m(N, xs, bp, shifts, resid_freq_check=10, max_iter=100000, eps=1e-10)
        ctx.sync()
        t0 = time.perf_counter()
        info, _ = ctx.solve_cg_m(N, xs, bp, shifts, resid_freq_check=10, max_iter=100000, eps=1e-10)
        dt = time.perf_counter() - t0
        gb = (240.0 + 80.0 * len(shifts)) * info["iter"] * V / 1e9
        emit(kind="solve", L=L, solver="minv_vector_cg_m on D^dag D, shifts {0,.01,.05,.25}", seconds=dt,
             iterations=info["iter"], ops=info["ops_count"], success=info["success"], iterations_per_s=info["iter"] / dt,
             algorithmic_GB=gb, GBps=gb / dt, frac_of_measured_hbm_peak=gb / dt / PEAK,
             rel_residuals=[float(np.sqrt(r) / np.sqrt(ctx.norm2sq(bp))) for r in info["resSqmrhs"]])
        for o in (D, N, Dd):
            o.destroy()
        del xs, x, b, bp

    # ---- 3. config 5 pieces: coarse stencil apply (512^2, nc = 8) and the nc = 1 fine stencil at 2048^2
    rg = np.random.default_rng(0)
    for X, nc, two in ([(256, 8, False)] if quick else [(512, 8, False), (512, 8, True), (256, 16, False), (2048, 1, False),
                                                        (512, 4, False)]):
        Vc = X * X
        m = Vc * nc * nc
        cl = (rg.standard_normal(m) + 1j * rg.standard_normal(m))
        hp = (rg.standard_normal(4 * m) + 1j * rg.standard_normal(4 * m))
        tl = (rg.standard_normal(8 * m) + 1j * rg.standard_normal(8 * m)) if two else None
        op = ctx.stencil2d(cl, hp, tl, X, X, nc, shift=0.1)
        v = ctx.vector(Vc * nc).upload(rg.standard_normal(Vc * nc) + 0j)
        w = ctx.vector(Vc * nc)
        ms = time_loop(lambda: op.apply(w, v), 40)
        bps = ((13 if two else 5) * nc * nc + 2 * nc) * 16
        emit(kind="apply", L=X, op="apply_stencil_2d nc=%d%s" % (nc, " two-link" if two else ""), ms=ms,
             bytes_per_site=bps, GBps=bps * Vc / ms / 1e6, frac_of_measured_hbm_peak=bps * Vc / ms / 1e6 / PEAK)
        op.destroy()
        del v, w, cl, hp, tl

    # ---- 3b. config 5, gated solves: unpreconditioned minv_vector_gcr_restart(..., 64, ...) on the fine level
    # (nc = 1 stencil of the 2048^2 staggered operator, built as get_square_staggered_u1_stencil does,
    # operators_stencil.cpp:14-63; tol 5e-7 = the outer precision of input_params.cpp:638) and on a
    # 512^2 x 8 coarse level (synthetic diagonally dominant stencil -- the MG set-up that would produce
    # the real one is SURVEY 8f-2; tol 1e-2 = the coarse precision, input_params.cpp:790)
    if not quick:
        L = 2048
        V = L * L
        rows = list(range(L))
        U = bench.gauge_rows(L, rows).reshape(L, L, 2)
        eta = (1.0 - 2.0 * (np.arange(L) % 2))[None, :]
        hop = np.empty((4, L, L), dtype=np.complex128)
        hop[0] = -0.5 * U[:, :, 0]
        hop[1] = -0.5 * eta * U[:, :, 1]
        hop[2] = 0.5 * np.conj(np.roll(U[:, :, 0], 1, axis=1))
        hop[3] = 0.5 * eta * np.conj(np.roll(U[:, :, 1], 1, axis=0))
        op = ctx.stencil2d(np.zeros(V, dtype=np.complex128), hop.reshape(-1), None, L, L, 1, shift=0.1)
        Dref = ctx.staggered(U.reshape(-1), L, L, 0.1, 0)
        bh = bench.rhs_rows(L, rows)
        b, x, chk = ctx.vector(V).upload(bh), ctx.vector(V), ctx.vector(V)
        op.apply(x, b)
        Dref.apply(chk, b)
        xs_, cs_ = x.download(), chk.download()   # staggered_stencil.cpp:232: function operator vs stencil operator
        same = float(np.linalg.norm(xs_ - cs_) / np.linalg.norm(cs_))
        del xs_, cs_
        for _ in range(2):
            x.zero()
            ctx.sync()
            t0 = time.perf_counter()
            info = ctx.solve("GCR_RESTART", op, x, b, max_iter=100000, eps=5e-7, restart_freq=64)
            dt = time.perf_counter() - t0
        gb = (288.0 + 48.0 * 31.5 + (112.0 - 64.0)) * info["iter"] * V / 1e9   # d-bytes GCR, stencil apply 112 B/site
        emit(kind="solve", L=L, solver="config 5 fine level: minv_vector_gcr_restart(64) on the nc=1 stencil, tol 5e-7",
             seconds=dt, iterations=info["iter"], ops=info["ops_count"], success=info["success"],
             iterations_per_s=info["iter"] / dt, stencil_vs_function_operator_rel_diff=same,
             true_rel_residual=float(np.sqrt(info["resSq"]) / np.sqrt(ctx.norm2sq(b))),
             algorithmic_GB=gb, GBps=gb / dt, frac_of_measured_hbm_peak=gb / dt / PEAK)
        op.destroy()
        Dref.destroy()
        del b, x, chk, hop, U
        Xc, nc = 512, 8
        Vc = Xc * Xc
        m = Vc * nc * nc
        cl = 0.1 * (rg.standard_normal(m) + 1j * rg.standard_normal(m))
        cl.reshape(Vc, nc, nc)[:, np.arange(nc), np.arange(nc)] += 1.0
        hp = 0.2 * (rg.standard_normal(4 * m) + 1j * rg.standard_normal(4 * m)) / np.sqrt(nc)
        op = ctx.stencil2d(cl, hp, None, Xc, Xc, nc)
        b = ctx.vector(Vc * nc).upload(rg.standard_normal(Vc * nc) + 1j * rg.standard_normal(Vc * nc))
        x = ctx.vector(Vc * nc)
        for _ in range(2):
            x.zero()
            ctx.sync()
            t0 = time.perf_counter()
            info = ctx.solve("GCR_RESTART", op, x, b, max_iter=1024, eps=1e-2, restart_freq=64)
            dt = time.perf_counter() - t0
        emit(kind="solve", L=Xc, solver="config 5 coarse level: minv_vector_gcr_restart(64) on a 512^2 x 8 stencil, tol 1e-2",
             seconds=dt, iterations=info["iter"], ops=info["ops_count"], success=info["success"],
             iterations_per_s=info["iter"] / dt,
             true_rel_residual=float(np.sqrt(info["resSq"]) / np.sqrt(ctx.norm2sq(b))))
        op.destroy()
        del b, x, cl, hp

    # ---- 4. config 1: real 64^2 Laplace, CG to 1e-10 (launch/latency-bound: 32 KiB vectors)
    N = 64
    b = np.zeros(N * N)
    b[N // 2 + (N // 2) * N] = 1.0
    op = ctx.laplace(N, N, 1, 4.01, np.float64)
    bd = ctx.vector(N * N, np.float64).upload(b)
    xd = ctx.vector(N * N, np.float64)
    for force in (False, True):
        ctx.force_host_scalars(force)
        xd.zero()
        ctx.solve("CG", op, xd, bd, max_iter=4000, eps=1e-10)
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(20):
            xd.zero()
            info = ctx.solve("CG", op, xd, bd, max_iter=4000, eps=1e-10)
        dt = (time.perf_counter() - t0) / 20
        emit(kind="solve", L=N, solver="config 1: minv_vector_cg, real 64^2 Laplace (%s)" % (
            "host-scalar shell" if force else "device-resident loop"), seconds=dt, iterations=info["iter"],
            iterations_per_s=info["iter"] / dt, us_per_iteration=1e6 * dt / info["iter"])
    ctx.force_host_scalars(False)


if __name__ == "__main__":
    main()
