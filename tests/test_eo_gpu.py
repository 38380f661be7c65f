"""Even/odd preconditioned staggered path on the GPU (operators.cpp:456-616) through the C ABI:
the pieces are bit-identical to the oracle, and prepare -> CG on m^2 - D_eo D_oe -> reconstruct solves D x = b
with the oracle's iteration count."""
import numpy as np
import pytest

from conftest import rel_err, synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [8, 64, 130])
def test_even_odd_pieces_bit_exact(ctx, glb, orc, L):
    U, b = synthetic(orc, L)
    w = orc.rng(11).gaussian(L * L)
    m = 0.13
    for flag, kind in ((glb.STAG_DEO, "STAG_DEO_U1"), (glb.STAG_DOE, "STAG_DOE_U1"), (glb.STAG_M2MDEODOE, "STAG_M2MDEODOE_U1")):
        assert np.array_equal(ctx.staggered(U, L, L, m, flag).apply_host(b), orc.op(kind, L, L, mass=m, links=U).apply(b))
    D, oD = ctx.staggered(U, L, L, m, 0), orc.op("STAG_U1", L, L, mass=m, links=U)
    db, dw = ctx.vector(L * L).upload(b), ctx.vector(L * L).upload(w)
    out = ctx.vector(L * L)
    D.eoprec_prepare(out, db)
    assert np.array_equal(out.download(), oD.eoprec_prepare(b))
    D.eoprec_reconstruct(out, db, dw)
    assert np.array_equal(out.download(), oD.eoprec_reconstruct(b, w))
    # the reference-named host entry points
    d = ctx._desc("STAG_U1", L, L, mass=m, links=U)
    assert np.array_equal(ctx.host_eoprec_prepare(d, b), oD.eoprec_prepare(b))
    assert np.array_equal(ctx.host_eoprec_reconstruct(d, b, w), oD.eoprec_reconstruct(b, w))
    # fused reductions of the composite operator
    M = ctx.staggered(U, L, L, m, glb.STAG_M2MDEODOE)
    dot, nrm = M.apply_dot(out, db, dw, want_norm=True)
    y = out.download()
    assert abs(dot - np.vdot(w, y)) <= 1e-12 * abs(np.vdot(w, y)) and abs(nrm - np.vdot(y, y).real) <= 1e-12 * nrm


@pytest.mark.parametrize("L,m", [(64, 0.1), (256, 0.05)])
def test_even_odd_preconditioned_solve(ctx, glb, orc, L, m):
    U, b = synthetic(orc, L)
    D, M = ctx.staggered(U, L, L, m, 0), ctx.staggered(U, L, L, m, glb.STAG_M2MDEODOE)
    oD, oM = orc.op("STAG_U1", L, L, mass=m, links=U), orc.op("STAG_M2MDEODOE_U1", L, L, mass=m, links=U)
    db, be, xe, x = ctx.vector(L * L).upload(b), ctx.vector(L * L), ctx.vector(L * L), ctx.vector(L * L)
    D.eoprec_prepare(be, db)
    xe.zero()
    got = ctx.solve("CG", M, xe, be, max_iter=20000, eps=1e-10)        # device-resident CG loop on the e/o system
    D.eoprec_reconstruct(x, xe, db)
    _, want = orc.solve("CG", oM, oD.eoprec_prepare(b), max_iter=20000, eps=1e-10)
    assert got["success"] and abs(got["iter"] - want["iter"]) <= max(1, int(round(0.02 * want["iter"])))
    xs = x.download()
    assert np.linalg.norm(oD.apply(xs) - b) / np.linalg.norm(b) < 1e-8
    # against CGNE on the full lattice: D^dag D is block diagonal in parity, so the e/o system takes no more iterations
    N = ctx.staggered(U, L, L, m, glb.STAG_NORMAL)
    bp, xn = ctx.vector(L * L), ctx.vector(L * L)
    ctx.staggered(U, L, L, m, glb.STAG_DAGGER).apply(bp, db)
    xn.zero()
    plain = ctx.solve("CG", N, xn, bp, max_iter=20000, eps=1e-10)
    assert got["iter"] <= plain["iter"] + 1
    assert rel_err(xs, xn.download()) < 1e-6


@pytest.mark.parametrize("X,Y,nc,two", [(6, 8, 4, False), (8, 6, 2, True), (5, 7, 3, True), (64, 64, 8, False), (32, 48, 8, True)])
def test_partial_stencil_applies_bit_exact(ctx, glb, orc, X, Y, nc, two):
    """glb_op_apply_part and the reference-named host entry points apply_stencil_2d_{eo,oe,tb,bt}"""
    rg = np.random.default_rng(X * 100 + nc)
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    V = X * Y
    cl, hp, tl, v = rc(V * nc * nc), rc(4 * V * nc * nc), (rc(8 * V * nc * nc) if two else None), rc(V * nc)
    kw = dict(shift=0.3, eo_shift=0.1j, dof_shift=0.2)
    oop = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, two_link=tl, **kw)
    op = ctx.stencil2d(cl, hp, tl, X, Y, nc, **kw)
    dv, out = ctx.vector(V * nc).upload(v), ctx.vector(V * nc)
    d = ctx._desc("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, two_link=tl, **kw)
    for part in ("EO", "OE", "TB", "BT"):
        want = oop.apply_part(part, v)
        op.apply_part(part, out, dv)
        assert np.array_equal(out.download(), want)
        assert np.array_equal(ctx.host_stencil_part(d, part, v), want)
