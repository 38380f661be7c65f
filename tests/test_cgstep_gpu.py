"""GPU parity of the single-kernel CG iteration (csrc/cgstep.cu) vs the reference's minv_vector_cg
(generic_cg.cpp:278-377) on D^dag D = square_staggered_normal_u1 (operators.cpp:444).

The kernel forms r, x, p, Ap and the four inner products of one CG iteration in one pass and predicts the
numerator of beta one step ahead; everything else is the reference's arithmetic.  Checked here, through the C ABI:
  * after m = 1..4 iterations the iterate equals the reference's to rounding (max_iter = m forces the exit);
  * complete solves take the reference's iteration count (bar: +-2 %, in practice equal), reach the same true
    residual (recomputed with the ORACLE's operator), and the prediction never deviates by more than 1e-10;
  * lattice shapes that exercise the x seam (tile wider than the lattice, several wraps), partial strips, row blocks
    of uneven height and the two-kernel loop (glb_cg_step_mode(0)) as cross-check.
"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
MASS = 0.1


def inputs(orc, X, Y, seed=1337):
    r = orc.rng(seed)
    U = r.gauss_gauge_u1(X, Y, 6.0)
    b = r.gaussian(X * Y)
    return U, b


def close_iters(got, want):
    return abs(got - want) <= max(1, int(round(0.02 * want)))


def dev_cg(ctx, glb, U, X, Y, rhs, max_iter, eps=1e-10, x0=None):
    op = ctx.staggered(U, X, Y, MASS, glb.STAG_NORMAL)
    x = ctx.vector(X * Y)
    x.upload(x0) if x0 is not None else x.zero()
    b = ctx.vector(X * Y).upload(rhs)
    rep = ctx.cg_device(op, x, b, max_iter=max_iter, eps=eps, want_history=True)
    return x.download(), rep


@pytest.mark.parametrize("X,Y", [(4, 4), (16, 16), (64, 64), (130, 34), (256, 48), (2, 40), (114, 9)])
def test_first_iterations_match_reference(ctx, glb, orc, X, Y):
    U, b0 = inputs(orc, X, Y)
    oN = orc.op("STAG_NORMAL_U1", X, Y, mass=MASS, links=U)
    rhs = orc.op("STAG_DAGGER_U1", X, Y, mass=MASS, links=U).apply(b0)
    x0 = orc.rng(7).gaussian(X * Y)
    assert ctx.cg_step_mode(True)
    for m in (1, 2, 3, 4):
        want, winfo = orc.solve("CG", oN, rhs, x0=x0, max_iter=m, eps=1e-30)
        got, rep = dev_cg(ctx, glb, U, X, Y, rhs, m, eps=1e-30, x0=x0)
        assert rep["iterations"] == winfo["iter"] == m and rep["hit_max_iter"]
        assert rel_err(got, want) < 1e-13, (X, Y, m, rel_err(got, want))


@pytest.mark.parametrize("L", [64, 256, 1024])
def test_solve_counts_and_residual(ctx, glb, orc, golden, L):
    U, b0 = inputs(orc, L, L)
    oN = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    rhs = orc.op("STAG_DAGGER_U1", L, L, mass=MASS, links=U).apply(b0)
    if str(L) in golden["synthetic_beta6_m0.1"]:
        want_it = golden["synthetic_beta6_m0.1"][str(L)]["CGNE"]["iter"]
    else:
        want_it = orc.solve("CG", oN, rhs, max_iter=10000, eps=1e-10)[1]["iter"]
    ctx.cg_step_mode(True)
    x1, rep1 = dev_cg(ctx, glb, U, L, L, rhs, 10000)
    err1 = ctx.cg_last_pred_err()
    ctx.cg_step_mode(False)
    try:
        x2, rep2 = dev_cg(ctx, glb, U, L, L, rhs, 10000)
    finally:
        ctx.cg_step_mode(True)
    assert close_iters(rep1["iterations"], want_it), (rep1["iterations"], want_it)
    assert close_iters(rep2["iterations"], want_it), (rep2["iterations"], want_it)
    assert 0.0 < err1 < 1e-10, err1
    for x in (x1, x2):
        rr = float(np.linalg.norm(oN.apply(x) - rhs) / np.linalg.norm(rhs))
        assert rr < 1.05e-10, rr
    assert rel_err(x1, x2) < 1e-8          # same solution from both loops (condition number ~ 1e2)
    # the recurrence residuals of the two loops follow each other to rounding for the first iterations
    h1, h2 = rep1["history"], rep2["history"]
    k = min(20, len(h1), len(h2))
    assert np.allclose(h1[:k], h2[:k], rtol=1e-9)


def test_reference_call_uses_single_kernel_loop(ctx, glb, orc):
    """minv_vector_cg(host vectors, square_staggered_normal_u1): the drop-in entry point ends in the same loop"""
    L = 64
    U, b0 = inputs(orc, L, L)
    oN = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    rhs = orc.op("STAG_DAGGER_U1", L, L, mass=MASS, links=U).apply(b0)
    _, want = orc.solve("CG", oN, rhs, max_iter=10000, eps=1e-10)
    x = np.zeros(L * L, dtype=np.complex128)
    n0 = ctx.launches()
    info = ctx.host_solve("CG", ctx._desc("STAG_NORMAL_U1", L, L, mass=MASS, links=U), x, rhs, max_iter=10000, eps=1e-10)
    launched = ctx.launches() - n0
    assert info["success"] and close_iters(info["iter"], want["iter"]) and info["ops_count"] == want["ops_count"]
    assert launched < 1.5 * info["iter"] + 40, launched   # ~ one kernel per iteration, not two or three
    assert float(np.linalg.norm(oN.apply(x) - rhs) / np.linalg.norm(rhs)) < 1.05e-10


def test_max_iter_and_zero_rhs(ctx, glb, orc):
    L = 32
    U, b0 = inputs(orc, L, L)
    oN = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    rhs = oN.apply(b0)
    want, winfo = orc.solve("CG", oN, rhs, max_iter=7, eps=1e-10)
    got, rep = dev_cg(ctx, glb, U, L, L, rhs, 7)
    assert rep["iterations"] == 7 and rep["hit_max_iter"] and winfo["iter"] == 7 and not winfo["success"]
    assert rel_err(got, want) < 1e-12
