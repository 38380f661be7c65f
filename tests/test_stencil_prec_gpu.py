"""The preconditioned stencil paths (rest of SURVEY 8f-3) on the GPU through the C ABI: the composite operators the
reference builds on a stencil_2d -- m^2 - D_eo D_oe, m^2 - D_tb D_bt, the non-Galerkin normal operators, the daggered
operators (operators_stencil.cpp:196-214, mg_complex.cpp:1228-1372) -- and the prepare / reconstruct steps around a
preconditioned solve (operators_stencil.cpp:179-236, mg_complex.cpp:1211-1272).  Every one of them runs the
reference's per-element expressions without FMA contraction, so the results are compared for bit equality against
the reference-compiled checker; the solve by its iteration count and true residual.  (tests/
test_gpu_logic_on_mock_cpu.py runs this module's Python side against the CPU mock of the C ABI.)"""
import numpy as np
import pytest

import oracle_py
from conftest import rel_err, synthetic

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif("ref" not in oracle_py.available(),
                                 reason="the composite stencil operators are only in oracle/_ref/libref_oracle.so")]

EO_VIEWS = ["M2MDEODOE", "NORMAL_EO", "DAGGER_EO"]
TB_VIEWS = ["M2MDTBDBT", "NORMAL_TB", "DAGGER_TB"]


def _rand_stencil(X, Y, nc, seed, shift):
    rg = np.random.default_rng(seed)
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    V = X * Y
    return rc(V * nc * nc), rc(4 * V * nc * nc), rc(V * nc), rc(V * nc), dict(shift=shift)


@pytest.mark.parametrize("X,Y,nc,shift", [(8, 8, 1, 0.25), (6, 10, 4, 0.3), (16, 12, 8, 0.1 + 0.05j), (64, 64, 8, 0.02),
                                          (5, 7, 3, 0.4)])
def test_composite_operators_and_prec_steps_bit_exact(ctx, glb, X, Y, nc, shift):
    orc = oracle_py.load("ref")
    cl, hp, v, w, kw = _rand_stencil(X, Y, nc, X * 31 + nc, shift)
    n = X * Y * nc
    base = ctx.stencil2d(cl, hp, None, X, Y, nc, **kw)
    dv, dw, out = ctx.vector(n).upload(v), ctx.vector(n).upload(w), ctx.vector(n)
    views = EO_VIEWS + (TB_VIEWS if nc % 2 == 0 else ["DAGGER_TB"])
    for view in views:
        want = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, view=view, **kw).apply(v)
        op = base.view(view)
        op.apply(out, dv)
        assert np.array_equal(out.download(), want), view
        # the reference-named host callback
        d = ctx._desc("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, view=view, **kw)
        assert np.array_equal(ctx.host_apply(d, v), want), view
        # reductions next to a composite apply (what CR / GCR ask of a native operator)
        dot, nrm = op.apply_dot(out, dv, dw, want_norm=True)
        assert abs(dot - np.vdot(w, want)) <= 1e-12 * max(abs(np.vdot(w, want)), 1e-300)
        assert abs(nrm - np.vdot(want, want).real) <= 1e-12 * nrm
        op.destroy()
    plain = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, **kw)
    d = ctx._desc("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, **kw)
    for tb in ([0, 1] if nc % 2 == 0 else [0]):
        want_p = oracle_py.ref_stencil_prec(orc, plain, tb, v)
        want_r = oracle_py.ref_stencil_prec(orc, plain, tb, v, w)
        base.prec_prepare(tb, out, dv)
        assert np.array_equal(out.download(), want_p)
        base.prec_reconstruct(tb, out, dv, dw)
        assert np.array_equal(out.download(), want_r)
        assert np.array_equal(ctx.host_stencil_prec(d, tb, v), want_p)
        assert np.array_equal(ctx.host_stencil_prec(d, tb, v, w), want_r)


def test_view_follows_the_shift_of_its_base(ctx, glb):
    """the reference's set-up changes stencil_2d::shift in place; a view reads the base's shift at apply time"""
    orc = oracle_py.load("ref")
    X, Y, nc = 8, 8, 2
    cl, hp, v, _, _ = _rand_stencil(X, Y, nc, 5, 0.0)
    base = ctx.stencil2d(cl, hp, None, X, Y, nc, shift=0.5)
    view = base.view("M2MDTBDBT")
    dv, out = ctx.vector(v.size).upload(v), ctx.vector(v.size)
    for m in (0.5, 0.125):
        base.set_shifts(shift=m)
        assert view.get_shifts()[0] == complex(m)
        view.apply(out, dv)
        want = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, shift=m, view="M2MDTBDBT").apply(v)
        assert np.array_equal(out.download(), want)


@pytest.mark.parametrize("L,m,solver", [(32, 0.1, "CG"), (64, 0.05, "CG"), (32, 0.1, "BICGSTAB")])
def test_even_odd_preconditioned_solve_through_the_stencil(ctx, glb, L, m, solver):
    """the reference's sequence (null_gen.cpp:262-273): prepare, solve m^2 - D_eo D_oe on the even sites with the
    callback apply_square_staggered_m2mdeodoe_stencil, reconstruct -- host vectors, the reference's own calls"""
    orc = oracle_py.load("ref")
    U, b = synthetic(orc, L)
    sten = orc.op("STENCIL_FROM_STAG", L, L, mass=m, links=U)
    sten_m = orc.op("STENCIL_FROM_STAG", L, L, mass=m, links=U, view="M2MDEODOE")
    be_want = oracle_py.ref_stencil_prec(orc, sten, 0, b)
    xe_want, want = orc.solve(solver, sten_m, be_want, max_iter=20000, eps=1e-10)
    x_want = oracle_py.ref_stencil_prec(orc, sten, 0, xe_want, b)
    d = ctx._desc("STENCIL_FROM_STAG", L, L, mass=m, links=U)
    dm = ctx._desc("STENCIL_FROM_STAG", L, L, mass=m, links=U, view="M2MDEODOE")
    be = ctx.host_stencil_prec(d, 0, b)
    assert np.array_equal(be, be_want)
    xe = np.zeros_like(b)
    got = ctx.host_solve(solver, dm, xe, be, max_iter=20000, eps=1e-10)
    x = ctx.host_stencil_prec(d, 0, xe, b)
    assert got["success"] and want["success"] and got["name"] == want["name"]
    if solver == "CG":
        assert abs(got["iter"] - want["iter"]) <= max(1, int(round(0.02 * want["iter"])))
    D = orc.op("STAG_U1", L, L, mass=m, links=U)
    assert np.linalg.norm(D.apply(x) - b) / np.linalg.norm(b) < 1e-8
    assert rel_err(x, x_want) < 1e-6
    i = np.arange(L * L)
    odd = ((i % L + i // L) % 2) == 1
    assert np.all(xe[odd] == 0)        # the preconditioned system lives on the even sites only
