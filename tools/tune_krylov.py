#!/usr/bin/env python
"""Device-resident BiCGStab / CR loops (csrc/krylov.cu) on one GPU: CUDA-graph replay against direct launches and the
host-scalar shell, per lattice size.

    python tools/tune_krylov.py [L ...]          (default 256 512 1024 2048 4096)

One JSON object per line: us per iteration and GB/s on the bytes the kernels move."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg  # noqa: E402


def main():
    glb = _load_pkg()
    ctx = glb.Context(device=0)
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024, 2048, 4096]
    try:
        PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        PEAK = 6650.0


    def timed(fn, reps=3):
        fn()
        ctx.sync()
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            ctx.sync()
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        return r, best


    for L in sizes:
        V = L * L
        links, b_h = ctx.synthetic_inputs(L, L)
        D = ctx.staggered(links, L, L, 0.1, 0)
        N = ctx.staggered(links, L, L, 0.1, glb.STAG_NORMAL)
        Dd = ctx.staggered(links, L, L, 0.1, glb.STAG_DAGGER)
        b = ctx.vector(V).upload(b_h)
        bp = ctx.vector(V)
        Dd.apply(bp, b)
        x = ctx.vector(V)

        def solve(solver, op, rhs):
            x.zero()
            return ctx.solve(solver, op, x, rhs, max_iter=100000, eps=1e-10)

        for solver, op, rhs, variants in (
                ("BICGSTAB", D, b, [("graph", 0, True, False), ("direct", 0, False, False),
                                    ("host-scalar shell", 0, True, True)]),
                ("CR", N, bp, [("graph", 0, True, False), ("direct", 0, False, False), ("host-scalar shell", 0, True, True)])):
            for name, fuse, graph, shell in variants:
                pg = ctx.krylov_graph_mode(graph)
                ctx.force_host_scalars(shell)
                try:
                    info, dt = timed(lambda: solve(solver, op, rhs))
                finally:
                    ctx.force_host_scalars(False)
                    ctx.krylov_graph_mode(pg)
                bytes_it = 352.0 if solver == "BICGSTAB" else 272.0
                if shell:
                    bytes_it = 352.0 if solver == "BICGSTAB" else 304.0
                us = 1e6 * dt / max(info["iter"], 1)
                print(json.dumps({"L": L, "solver": solver, "variant": name, "iterations": info["iter"], "us_per_iteration": round(us, 2),
                                  "moved_GBps": round(bytes_it * V / us / 1e3, 1), "frac_of_hbm_peak": round(bytes_it * V / us / 1e3 / PEAK, 3)}),
                      flush=True)
        for o in (D, N, Dd):
            o.destroy()
        del x, b, bp


if __name__ == "__main__":
    main()
