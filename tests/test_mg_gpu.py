"""Multigrid preconditioner on the GPU: grid transfers, one V cycle and the VPGCR outer solve
against the REFERENCE's own multigrid (oracle/_ref, built from /root/reference here and shipped prebuilt to the
GPU box) on the same hierarchy.  Transfers are bit-exact; the cycle contains Krylov smoothers whose inner
products sum in a different order on the device, so it is compared to 1e-9 and the solve by iteration count
(+-2 %, in practice equal) and true residual."""
import numpy as np
import pytest

import oracle_py
from conftest import rel_err
from mg_common import build_reference_mg, quiet_stdout

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif("ref" not in oracle_py.available(),
                                 reason="the reference's multigrid is only in oracle/_ref/libref_oracle.so")]


def device_hierarchy(ctx, mg, U=None, mass=None, native_fine=False):
    """upload the reference-built hierarchy: stencil operators per level + the transfer"""
    ops = []
    for lvl in (0, 1):
        X, Y, nc = mg.dims(lvl)
        cl, hp, sh = mg.stencil(lvl)
        if lvl == 0 and native_fine:
            ops.append(ctx.staggered(U, X, Y, mass, 0))      # the function operator instead of its stencil
        else:
            ops.append(ctx.stencil2d(cl, hp, None, X, Y, nc, shift=sh[0], eo_shift=sh[1], dof_shift=sh[2]))
    Xf, Yf, _ = mg.dims(0)
    Xc, Yc, nv = mg.dims(1)
    tr = ctx.mg_transfer(Xf, Yf, 1, Xf // Xc, Yf // Yc, [mg.null(0, v) for v in range(nv)])
    return ops, tr


@pytest.mark.parametrize("L,nvec", [(16, 2), (64, 4)])
def test_transfers_bit_exact(ctx, glb, L, nvec):
    orc = oracle_py.load("ref")
    mg, U, b = build_reference_mg(orc, L=L, nvec=nvec)
    ops, tr = device_hierarchy(ctx, mg)
    rg = np.random.default_rng(L)
    c = rg.standard_normal(mg.size(1)) + 1j * rg.standard_normal(mg.size(1))
    f = rg.standard_normal(mg.size(0)) + 1j * rg.standard_normal(mg.size(0))
    dc, df = ctx.vector(mg.size(1)).upload(c), ctx.vector(mg.size(0)).upload(f)
    oc, of = ctx.vector(mg.size(1)), ctx.vector(mg.size(0))
    tr.prolong(of, dc)
    tr.restrict(oc, df)
    assert np.array_equal(of.download(), mg.prolong(0, c))
    assert np.array_equal(oc.download(), mg.restrict(0, f))
    # the uploaded coarse stencil is the Galerkin operator: coarse apply == restrict(fine apply(prolong))
    t1, t2 = ctx.vector(mg.size(0)), ctx.vector(mg.size(1))
    ops[0].apply(t1, of)
    tr.restrict(t2, t1)
    ops[1].apply(oc, dc)
    assert rel_err(oc.download(), t2.download()) < 1e-13


@pytest.mark.parametrize("L,nvec,cfg", [(16, 2, dict()), (32, 4, dict()),
                                         (32, 4, dict(smooth="BICGSTAB", n_pre=3, n_post=2, inner="CG")),
                                         (32, 2, dict(n_pre=0, n_post=4, inner="BICGSTAB")),
                                         (32, 2, dict(n_pre=2, n_post=0, inner="CR", n_restart=16)),
                                         (32, 4, dict(smooth="MINRES", n_pre=4, n_post=4))])
def test_vcycle_matches_reference(ctx, glb, L, nvec, cfg):
    """one cycle with the coarse system solved to 1e-11: a coarse solve stopped at the reference's default 1e-2
    may take one iteration more or less on the device (reduction order), which changes the cycle's output at the
    1e-2 level -- that case is covered by the outer-solve test below, where only iteration counts matter"""
    orc = oracle_py.load("ref")
    mg, U, b = build_reference_mg(orc, L=L, nvec=nvec)
    ops, tr = device_hierarchy(ctx, mg)
    dmg = ctx.multigrid(ops, [tr])
    full = dict(smooth="GCR", n_pre=6, n_post=6, inner="GCR", n_max=4096, n_restart=64, rel_res=1e-11)
    full.update(cfg)
    mg.set_precond(**full)
    dmg.set(**full)
    with quiet_stdout():
        want = mg.vcycle(b)
    out, rhs = ctx.vector(b.size), ctx.vector(b.size).upload(b)
    dmg.vcycle(out, rhs)
    # Krylov smoothers amplify the reduction-order rounding (BiCGStab most): measured 1e-12 .. 4e-9
    assert rel_err(out.download(), want) < 1e-6
    cnt = dmg.counts()
    assert cnt["presmooth"][0] == (full["n_pre"] + 2 if full["n_pre"] else 0) or full["smooth"] != "GCR"
    assert cnt["krylov"][1] > 0


@pytest.mark.parametrize("L,nvec,restart,native", [(32, 4, 64, False), (64, 4, 64, False), (64, 4, 8, False),
                                                   (64, 4, 64, True)])
def test_vpgcr_mg_solve(ctx, glb, L, nvec, restart, native):
    """config 5's outer solve (VPGCR(64), tol 5e-7, GCR smoother 6+6, coarse GCR(64) to 1e-2) at test size"""
    orc = oracle_py.load("ref")
    mass = 0.01
    mg, U, b = build_reference_mg(orc, L=L, mass=mass, nvec=nvec)
    ops, tr = device_hierarchy(ctx, mg, U=U, mass=mass, native_fine=native)
    dmg = ctx.multigrid(ops, [tr])
    mg.set_precond()
    dmg.set()
    with quiet_stdout():
        xo, want = mg.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=restart)
    x, rhs = ctx.vector(b.size), ctx.vector(b.size).upload(b)
    x.zero()   # pool memory is not cleared: the initial guess is the caller's business, as in the reference
    got = dmg.vpgcr(x, rhs, max_iter=1000, eps=5e-7, restart_freq=restart)
    assert got["success"] and want["success"]
    assert abs(got["iter"] - want["iter"]) <= max(1, int(0.02 * want["iter"]))
    D = orc.op("STAG_U1", L, L, mass=mass, links=U)
    xs = x.download()
    assert np.linalg.norm(b - D.apply(xs)) / np.linalg.norm(b) < 5e-7 * 1.0001
    assert rel_err(xs, xo) < 1e-4
    # against the unpreconditioned solver the reference's own tests compare with: an order of magnitude fewer applies
    x0 = ctx.vector(b.size)
    x0.zero()
    plain = ctx.solve("GCR_RESTART", ops[0], x0, rhs, max_iter=100000, eps=5e-7, restart_freq=64)
    assert plain["iter"] > 5 * got["iter"]
