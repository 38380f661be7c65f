"""Device-resident BiCGStab / CR loops (csrc/krylov.cu) against the host-scalar shells and the CPU oracle.

The device loops run the shells' own vector kernels with the scalars formed on the GPU, so on the same inputs they
must reproduce the shells' iterates (same iteration count, same solution) and, through them, the reference's counts
(generic_bicgstab.cpp:258-308, generic_cr.cpp:246-286).  Checked on the staggered operator and D^dag D (complex), the
real Laplacian, and the coarse stencil; with batches replayed as a CUDA graph and launched directly.
"""
import numpy as np
import pytest

from conftest import rel_err, synthetic
from test_solvers_gpu import iters_ok, run_dev, true_rel_residual

pytestmark = pytest.mark.gpu


def both_ways(ctx, op, solver, b, x0=None, **kw):
    ctx.force_host_scalars(True)
    try:
        xs, shell = run_dev(ctx, op, solver, b, x0=x0, **kw)
    finally:
        ctx.force_host_scalars(False)
    xd, dev = run_dev(ctx, op, solver, b, x0=x0, **kw)
    return xs, shell, xd, dev


@pytest.mark.parametrize("L", [64, 256])
def test_staggered_device_loops_equal_shells(ctx, glb, orc, L):
    U, b = synthetic(orc, L)
    D = ctx.staggered(U, L, L, 0.1, 0)
    Nrm = ctx.staggered(U, L, L, 0.1, glb.STAG_NORMAL)
    oD = orc.op("STAG_U1", L, L, mass=0.1, links=U)
    oN = orc.op("STAG_NORMAL_U1", L, L, mass=0.1, links=U)
    assert ctx.krylov_supported("BICGSTAB", D) and ctx.krylov_supported("CR", Nrm)
    for solver, op, oop in (("BICGSTAB", D, oD), ("CR", Nrm, oN), ("BICGSTAB", Nrm, oN), ("CR", D, oD)):
        if solver == "CR" and op is D:
            continue  # CR needs a Hermitian operator
        xs, shell, xd, dev = both_ways(ctx, op, solver, b, max_iter=100000, eps=1e-10)
        assert dev["iter"] == shell["iter"] and dev["ops_count"] == shell["ops_count"], (solver, dev, shell)
        assert dev["success"] and shell["success"] and dev["name"] == shell["name"]
        assert rel_err(xd, xs) < 1e-12, (solver, rel_err(xd, xs))
        assert true_rel_residual(oop, xd, b) < 1e-10 * 1.0000001
        assert abs(np.sqrt(dev["resSq"]) / np.linalg.norm(b) - true_rel_residual(oop, xd, b)) < 1e-12
        want = orc.solve(solver, oop, b, max_iter=100000, eps=1e-10)[1]
        assert iters_ok(orc, solver, oop, b, dev["iter"], want["iter"], max_iter=100000, eps=1e-10), (solver, dev, want)
        print("%s %s L=%d: device loop %d iterations (shell %d, oracle %d), bit-identical to the shell: %s"
              % (solver, "D" if op is D else "D^dag D", L, dev["iter"], shell["iter"], want["iter"],
                 bool(np.array_equal(xd, xs))))


def test_graph_and_direct_batches_agree(ctx, glb, orc):
    L = 64
    U, b = synthetic(orc, L)
    D = ctx.staggered(U, L, L, 0.1, 0)
    outs = {}
    for graph in (True, False):
        prev = ctx.krylov_graph_mode(graph)
        try:
            x = ctx.vector(D.local_size, D.dtype).zero()
            bd = ctx.vector(D.local_size, D.dtype).upload(b)
            rep = ctx.krylov_device("BICGSTAB", D, x, bd, max_iter=100000, eps=1e-10, want_history=True)
            outs[graph] = (x.download(), rep)
        finally:
            ctx.krylov_graph_mode(prev)
    (xg, rg), (xn, rn) = outs[True], outs[False]
    assert rg["iterations"] == rn["iterations"] and rg["ops"] == rn["ops"]
    assert np.array_equal(xg, xn)
    assert np.array_equal(rg["history"], rn["history"])
    assert not rn["used_graph"]
    assert rg["used_graph"], "batches were not replayed as a CUDA graph (capture failed?)"
    # the recurrence residual of the last iteration is the one that passed the stopping test
    assert np.sqrt(rg["history"][-1]) < 1e-10 * rg["bnorm"] and np.sqrt(rg["history"][-2]) >= 1e-10 * rg["bnorm"]
    assert rg["ops"] == 2 * rg["iterations"] + 1


def test_iteration_cap_and_initial_guess(ctx, glb, orc):
    """k == max_iter-1 ends the loop with success = false (real CR and both BiCGStab overloads; the complex CR never
    reports failure, generic_cr.cpp:288), and the initial guess is used"""
    L = 64
    U, b = synthetic(orc, L)
    D = ctx.staggered(U, L, L, 0.1, 0)
    Nrm = ctx.staggered(U, L, L, 0.1, glb.STAG_NORMAL)
    oD = orc.op("STAG_U1", L, L, mass=0.1, links=U)
    oN = orc.op("STAG_NORMAL_U1", L, L, mass=0.1, links=U)
    for cap in (1, 7, 8, 9, 17):  # around the batch size of the enqueue loop
        for solver, op, oop in (("BICGSTAB", D, oD), ("CR", Nrm, oN)):
            xs, shell, xd, dev = both_ways(ctx, op, solver, b, max_iter=cap, eps=1e-10)
            want = orc.solve(solver, oop, b, max_iter=cap, eps=1e-10)[1]
            assert dev["iter"] == shell["iter"] == want["iter"] == cap, (solver, cap, dev, shell, want)
            assert dev["success"] == shell["success"] == want["success"], (solver, cap)
            assert dev["ops_count"] == shell["ops_count"] == want["ops_count"], (solver, cap)
            assert rel_err(xd, xs) < 1e-12
    x0 = orc.rng(7).gaussian(L * L)
    xs, shell, xd, dev = both_ways(ctx, D, "BICGSTAB", b, x0=x0, max_iter=100000, eps=1e-10)
    assert dev["iter"] == shell["iter"] and rel_err(xd, xs) < 1e-12 and true_rel_residual(oD, xd, b) < 1.0000001e-10


def test_real_laplace_and_coarse_stencil(ctx, glb, orc):
    """the real overloads (generic_bicgstab.cpp:22, generic_cr.cpp:28) and a stencil2d operator"""
    N = 128
    b = np.zeros(N * N)
    b[N // 2 + (N // 2) * N] = 1.0
    lap = ctx.laplace(N, N, 1, 4 + 0.01, np.float64)
    olap = orc.op("LAPLACE_REAL", N, N, mass=0.01)
    for solver in ("CR", "BICGSTAB"):
        xs, shell, xd, dev = both_ways(ctx, lap, solver, b, x0=b.copy(), max_iter=4000, eps=1e-8)
        assert dev["iter"] == shell["iter"] and dev["success"] == shell["success"], (solver, dev, shell)
        assert rel_err(xd, xs) < 1e-12
        assert true_rel_residual(olap, xd, b) < 1e-8 * 1.0000001
        # real CR reports failure at the cap (generic_cr.cpp:117)
        xs, shell, xd, dev = both_ways(ctx, lap, solver, b, x0=b.copy(), max_iter=5, eps=1e-8)
        assert dev["iter"] == shell["iter"] == 5 and dev["success"] is False and shell["success"] is False
    # a diagonally dominant nc = 8 stencil (the ring kernel with the fused epilogue)
    X, Y, nc = 48, 40, 8
    V = X * Y
    rg = np.random.default_rng(5)
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    cl = 0.05 * rc(V * nc * nc)
    cl.reshape(V, nc, nc)[:, np.arange(nc), np.arange(nc)] += 4.0
    hp = 0.05 * rc(4 * V * nc * nc)
    bb = rc(V * nc)
    st = ctx.stencil2d(cl, hp, None, X, Y, nc, shift=0.25)
    ost = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, two_link=None, shift=0.25)
    xs, shell, xd, dev = both_ways(ctx, st, "BICGSTAB", bb, max_iter=1000, eps=1e-10)
    assert dev["iter"] == shell["iter"] and dev["success"] and rel_err(xd, xs) < 1e-12
    assert true_rel_residual(ost, xd, bb) < 1e-10 * 1.0000001
