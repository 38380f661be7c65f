"""Multigrid set-up on the GPU: the device kernels glb_mg_block_orthonormalize / glb_mg_partition /
glb_mg_galerkin and the complete set-up sequence (host/null_gen_dev.cpp, glbx_mg_setup) against the REFERENCE's own
set-up (oracle/_ref: block_orthonormalize, generate_coarse_from_fine_stencil, null_generate_random_smooth driven as
in aa_mg_square_staggered_u1.cpp:716-1143).

What can be compared how: block orthonormalisation and the partition run the reference's statements in its order
inside one thread -> compared element for element (gate 1e-14; bit equality is reported); the Galerkin product is
summed directly where the reference probes with 9 applies per colour -> 1e-13; the null vectors come out of
Krylov solves whose inner products sum in a different order on the device and are then pushed through a block
Gram-Schmidt of locally almost parallel vectors, which amplifies rounding by many orders of magnitude
(tests/test_mg_setup_mock_cpu.py) -> they are compared directly only after few smoothing iterations, and the fully
smoothed hierarchy by its properties (block orthonormality, coarse operator = P^dag A P) and by the preconditioned
solve it gives (iteration count next to the reference's)."""
import os
import sys

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, rel_err
from mg_common import quiet_stdout

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif("ref" not in oracle_py.available(),
                                 reason="the reference's multigrid is only in oracle/_ref/libref_oracle.so")]


def _raw_vectors(orc, L, n, dof=1, seed=3, split="site"):
    """n gaussian vectors on an L x L x dof lattice, split BLOCK_EO style into 2n (even part first)"""
    rng = orc.rng(seed)
    raw = [rng.gaussian(L * L * dof) for _ in range(n)]
    idx = np.arange(L * L * dof)
    if split == "site":
        site = idx // dof
        even = ((site % L + site // L) % 2) == 0
    else:
        even = (idx % dof) < dof // 2
    return [np.where(even, v, 0) for v in raw] + [np.where(~even, v, 0) for v in raw]


@pytest.mark.parametrize("L,block,nraw", [(16, 4, 2), (32, 4, 4), (24, 2, 1), (64, 8, 3)])
def test_block_orthonormalize_and_galerkin_level0(ctx, glb, L, block, nraw):
    orc = oracle_py.load("ref")
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    mass = 0.03
    vecs = _raw_vectors(orc, L, nraw)
    nv = len(vecs)
    for ignore in (False, True):
        ref = oracle_py.RefMg(orc, L, L, U, mass, [block], [nv], [vecs], ignore_shifts=ignore)
        dv = [ctx.vector(L * L).upload(v) for v in vecs]
        ctx.mg_block_orthonormalize(L, L, 1, block, block, dv)
        exact = True
        for v in range(nv):
            got, want = dv[v].download(), ref.null(0, v)
            assert rel_err(got, want) < 1e-14
            exact = exact and np.array_equal(got, want)
        print("block_orthonormalize L=%d block=%d nvec=%d bit-identical: %s" % (L, block, nv, exact))
        cl0, hp0, sh0 = ref.stencil(0)
        fine = ctx.stencil2d(cl0, hp0, None, L, L, 1, shift=sh0[0], eo_shift=sh0[1], dof_shift=sh0[2])
        tr = glb.MgTransfer(ctx, L, L, 1, block, block, dv)          # from the device-resident vectors
        coarse = tr.galerkin(fine, ignore_shifts=ignore)
        Lc = L // block
        cl, hp = coarse.stencil_download(Lc, Lc, nv)
        clr, hpr, shr = ref.stencil(1)
        assert rel_err(cl, clr) < 1e-13 and rel_err(hp, hpr) < 1e-13
        assert coarse.get_shifts() == (0j, 0j, 0j)
        # the transfer built from device vectors is the reference's prolongator
        rg = np.random.default_rng(L)
        c = rg.standard_normal(ref.size(1)) + 1j * rg.standard_normal(ref.size(1))
        of = ctx.vector(L * L)
        tr.prolong(of, ctx.vector(ref.size(1)).upload(c))
        assert rel_err(of.download(), ref.prolong(0, c)) < 1e-14
        # and the coarse operator applies like the reference's coarse stencil (shift as the driver sets it)
        if ignore:
            coarse.set_shifts(shift=shr[0])
        out = ctx.vector(ref.size(1))
        coarse.apply(out, ctx.vector(ref.size(1)).upload(c))
        assert rel_err(out.download(), ref.apply_level(1, c)) < 1e-13


def test_galerkin_below_the_top_level(ctx, glb):
    """fine level with 4 dofs per site (clover + hopping matrices) -> 2 x 2 blocks, 4 coarse colours"""
    orc = oracle_py.load("ref")
    L, mass = 32, 0.02
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    v0 = _raw_vectors(orc, L, 2, seed=5)
    v1 = _raw_vectors(orc, L // 4, 2, dof=4, seed=6, split="colour")
    ref = oracle_py.RefMg(orc, L, L, U, mass, [4, 2], [4, 4], [v0, v1])
    X1, Y1, d1 = ref.dims(1)
    cl1, hp1, sh1 = ref.stencil(1)
    fine = ctx.stencil2d(cl1, hp1, None, X1, Y1, d1, shift=sh1[0], eo_shift=sh1[1], dof_shift=sh1[2])
    dv = [ctx.vector(X1 * Y1 * d1).upload(v) for v in v1]
    ctx.mg_block_orthonormalize(X1, Y1, d1, 2, 2, dv)
    for v in range(4):
        assert rel_err(dv[v].download(), ref.null(1, v)) < 1e-14
    tr = glb.MgTransfer(ctx, X1, Y1, d1, 2, 2, dv)
    coarse = tr.galerkin(fine)
    X2, Y2, d2 = ref.dims(2)
    cl, hp = coarse.stencil_download(X2, Y2, d2)
    clr, hpr, _ = ref.stencil(2)
    assert rel_err(cl, clr) < 1e-13 and rel_err(hp, hpr) < 1e-13


@pytest.mark.parametrize("X,Y,dof,period", [(16, 16, 1, 0), (12, 20, 1, 0), (8, 8, 4, 4), (8, 8, 4, 8), (6, 10, 6, 6)])
def test_partition_exact(ctx, X, Y, dof, period):
    rg = np.random.default_rng(X + dof)
    n = X * Y * dof
    v = rg.standard_normal(n) + 1j * rg.standard_normal(n)
    tgt = rg.standard_normal(n) + 1j * rg.standard_normal(n)       # only the "odd" elements may be overwritten
    idx = np.arange(n)
    if period:
        odd = (idx % period) >= period // 2
    else:
        site = idx // dof
        odd = ((site % X + site // X) % 2) == 1
    a, b = ctx.vector(n).upload(v), ctx.vector(n).upload(tgt)
    ctx.mg_partition(X, Y, dof, period, a, b)
    assert np.array_equal(a.download(), np.where(odd, 0, v))
    assert np.array_equal(b.download(), np.where(odd, v, tgt))


@pytest.mark.parametrize("X,Y,dof,period", [(16, 16, 1, 0), (12, 20, 2, 0), (8, 8, 8, 8), (8, 8, 4, 8), (6, 10, 6, 6)])
def test_partition_corner_exact(ctx, X, Y, dof, period):
    """BLOCK_CORNER: the three moves of null_gen.cpp:74-88 (sites) / :132-152 (colour quarters, integer bounds)"""
    rg = np.random.default_rng(X + dof)
    n = X * Y * dof
    v = rg.standard_normal(n) + 1j * rg.standard_normal(n)
    idx = np.arange(n)
    if period:
        c, p = idx % period, period
        cls = np.where((c >= p // 4) & (c < 2 * p // 4), 1, np.where((c >= 2 * p // 4) & (c < 3 * p // 4), 2,
                                                                    np.where(c >= 3 * p // 4, 3, 0)))
    else:
        site = idx // dof
        xo, yo = (site % X) % 2, (site // X) % 2
        cls = np.where((xo == 1) & (yo == 1), 1, np.where((xo == 1) & (yo == 0), 2, np.where((xo == 0) & (yo == 1), 3, 0)))
    src = ctx.vector(n).upload(v)
    for k in (1, 2, 3):
        tgt = rg.standard_normal(n) + 1j * rg.standard_normal(n)
        dst = ctx.vector(n).upload(tgt)
        ctx.mg_partition_corner(X, Y, dof, period, k, src, dst)
        assert np.array_equal(dst.download(), np.where(cls == k, v, tgt))
    assert np.array_equal(src.download(), np.where(cls == 0, v, 0))


def test_setup_argument_errors(ctx, glb):
    """the set-up entry points refuse what they cannot do, with a message, and leave nothing behind"""
    if ctx.cu.glb_device(ctx.h) < 0:
        pytest.skip("argument checking is the CUDA library's; the CPU mock of the C ABI has none")
    orc = oracle_py.load("ref")
    L = 16
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    fine, _ = _fine_stencil(ctx, orc, U, L, 0.05)
    D = ctx.staggered(U, L, L, 0.05, 0)
    vecs = [ctx.vector(L * L).upload(v) for v in _raw_vectors(orc, L, 1)]
    with pytest.raises(glb.GlbError):                      # block size does not divide the lattice
        glb.MgTransfer(ctx, L, L, 1, 3, 3, vecs)
    tr = glb.MgTransfer(ctx, L, L, 1, 4, 4, vecs)
    with pytest.raises(glb.GlbError):                      # the Galerkin kernel needs a stencil2d operator
        tr.galerkin(D)
    tr5 = glb.MgTransfer(ctx, 20, 20, 1, 4, 4, [ctx.vector(400).upload(np.ones(400)), ctx.vector(400).upload(np.ones(400))])
    cl, hp, _ = __import__("mg_setup").staggered_stencil(orc.rng(3).gauss_gauge_u1(20, 20, 6.0), 20, 20, 0.0)
    with pytest.raises(glb.GlbError):                      # 5 x 5 coarse sites: the reference's even/odd probing needs an even extent
        tr5.galerkin(ctx.stencil2d(cl, hp, None, 20, 20, 1, shift=0.1))
    with pytest.raises(glb.GlbError):                      # views exist on stencil2d operators only ...
        D.view("M2MDEODOE")
    with pytest.raises(glb.GlbError):                      # ... and not on other views
        fine.view("NORMAL_EO").view("DAGGER_EO")
    odd = ctx.stencil2d(np.zeros(64 * 9, complex), np.zeros(4 * 64 * 9, complex), None, 8, 8, 3, shift=0.1)
    with pytest.raises(glb.GlbError):                      # top/bottom needs an even number of colours
        odd.view("M2MDTBDBT")
    with pytest.raises(glb.GlbError):
        odd.prec_prepare(1, ctx.vector(192), ctx.vector(192))
    with pytest.raises(glb.GlbError):                      # BLOCK_TOPO without the gauge links is refused (stderr says why)
        ctx.multigrid_setup(fine, L, L, [4], [4], bstrat=3, max_iter=2)
    assert fine.get_shifts()[0] == complex(0.05)           # and the caller's shift is back
    with pytest.raises(glb.GlbError):
        ctx.mg_partition_corner(8, 8, 1, 0, 4, vecs[0], vecs[1])


def _fine_stencil(ctx, orc, U, L, mass):
    import mg_setup
    cl0, hp0, sh0 = mg_setup.staggered_stencil(U, L, L, 0.0)
    return ctx.stencil2d(cl0, hp0, None, L, L, 1, shift=mass), hp0


def test_setup_few_smoothing_iterations_matches_reference(ctx, glb):
    """3 BiCGStab iterations per vector: the rounding amplification of the block Gram-Schmidt is still moderate, so the
    device set-up can be held directly against the reference's from the same std::mt19937 seed (measured on the B200:
    2e-7 -- the device's inner products sum in a different order, 1e-16, and the Gram-Schmidt of the locally similar
    vectors does the rest; on the CPU mock, where the order is the reference's, the vectors are bit-identical)"""
    orc = oracle_py.load("ref")
    L, mass, blocks, nvecs = 32, 0.05, [4], [4]
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    kw = dict(seed=17, max_iter=3)
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, mass, blocks, nvecs, **kw)
    fine, _ = _fine_stencil(ctx, orc, U, L, mass)
    mg = ctx.multigrid_setup(fine, L, L, blocks, nvecs, **kw)
    for v in range(nvecs[0]):
        assert rel_err(mg.null_vector(0, v), ref.null(0, v)) < 1e-5
    cl, hp, sh = mg.level_stencil(1)
    clr, hpr, shr = ref.stencil(1)
    assert rel_err(cl, clr) < 1e-5 and rel_err(hp, hpr) < 1e-5
    assert sh == (complex(mass), 0j, 0j) and fine.get_shifts()[0] == complex(mass)
    assert mg.counts()["nullvectors"] == ref.null_counts()
    mg.destroy()


def test_setup_handle_outlives_its_fine_operator(ctx, glb):
    """the set-up handle owns the coarse levels, the caller owns level 0: destroying them in either order is fine"""
    orc = oracle_py.load("ref")
    L = 16
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    fine, _ = _fine_stencil(ctx, orc, U, L, 0.05)
    mg = ctx.multigrid_setup(fine, L, L, [4], [4], seed=1, max_iter=2)
    fine.destroy()
    mg.destroy()


@pytest.mark.parametrize("L,blocks,nvecs,opts", [(32, [4], [4], dict()), (64, [4], [8], dict()),
                                                 (64, [4], [8], dict(do_ortho_eo=True)),
                                                 (64, [4, 2], [4, 4], dict()),
                                                 # preconditioned null-vector solves (null_gen.cpp:259-313): the e/o
                                                 # (t/b below the top level) system, and the normal equations
                                                 (64, [4], [8], dict(null_prec=1, null_gen="CG")),
                                                 (64, [4, 2], [4, 4], dict(null_prec=1)),
                                                 (64, [4], [8], dict(null_prec=2, null_gen="CG", tol=1e-3)),
                                                 # BLOCK_CORNER (four parts per smoothed vector), free-field vectors
                                                 (64, [4], [8], dict(bstrat=2)),
                                                 (32, [4], [2], dict(do_free=True)),
                                                 # BLOCK_TOPO: chiral projectors from the symmetric shifts of the links
                                                 (64, [4], [8], dict(bstrat=3))])
def test_setup_defaults_hierarchy_and_solve(ctx, glb, L, blocks, nvecs, opts):
    """the driver's defaults (BiCGStab to 5e-5, at most 500 iterations, null mass 1e-2, BLOCK_EO): structural
    properties of the device-built hierarchy and the outer solve VPGCR(64) + V cycle next to the reference's own
    set-up + solve from the same seed"""
    import mg_setup
    orc = oracle_py.load("ref")
    mass = 0.02
    rng = orc.rng(7)
    U = rng.gauss_gauge_u1(L, L, 6.0)
    b = rng.gaussian(L * L)
    kw = dict(seed=23, **opts)
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, mass, blocks, nvecs, **kw)
    fine, hp0 = _fine_stencil(ctx, orc, U, L, mass)
    mg = ctx.multigrid_setup(fine, L, L, blocks, nvecs, links=(U if opts.get("bstrat") == 3 else None), **kw)
    secs = mg.setup_seconds()
    assert secs["total"] > 0
    # block orthonormality of the top-level vectors: <v_i, v_j> = delta_ij inside every block
    vecs = np.array([mg.null_vector(0, v) for v in range(nvecs[0])])
    B = mg_setup._to_blocks(vecs, L, L, blocks[0], blocks[0])              # [Yc, Xc, nv, block dofs]
    gram = np.einsum("yxik,yxjk->yxij", np.conj(B), B)
    assert np.max(np.abs(gram - np.eye(nvecs[0]))) < 1e-12
    # BLOCK_EO: the first half lives on even sites, the second half on odd sites
    idx = np.arange(L * L)
    even = ((idx % L + idx // L) % 2) == 0
    if opts.get("bstrat", 1) == 1:
        for v in range(nvecs[0]):
            assert np.all(vecs[v][~even if v < nvecs[0] // 2 else even] == 0)
    # the level-1 operator is P^dag A P of the device's own vectors (numpy restatement checked against the
    # reference in tests/test_mg_setup_cpu.py), with the mass in the shift
    cl, hp, sh = mg.level_stencil(1)
    clc, hpc = mg_setup.coarse_stencil(list(vecs), hp0, 0.0, L, L, blocks[0], blocks[0])
    assert rel_err(cl, clc) < 1e-13 and rel_err(hp, hpc) < 1e-13
    assert sh == (complex(mass), 0j, 0j)
    # smoothing work spent: the same order as the reference's (iteration counts of the 5e-5 solves may differ by a few)
    got_n, want_n = mg.counts()["nullvectors"], ref.null_counts()
    assert abs(got_n[0] - want_n[0]) <= 0.25 * want_n[0] + 8
    # the outer solve
    ref.set_precond()
    mg.set()
    with quiet_stdout():
        xo, want = ref.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
    x, rhs = ctx.vector(b.size), ctx.vector(b.size).upload(b)
    x.zero()
    got = mg.vpgcr(x, rhs, max_iter=1000, eps=5e-7, restart_freq=64)
    assert got["success"] and want["success"]
    assert abs(got["iter"] - want["iter"]) <= max(2, 0.25 * want["iter"]), (got, want)
    D = orc.op("STAG_U1", L, L, mass=mass, links=U)
    xs = x.download()
    assert np.linalg.norm(b - D.apply(xs)) / np.linalg.norm(b) < 5e-7 * 1.0001
    assert rel_err(xs, xo) < 1e-4
    print("setup L=%d blocks=%s nvecs=%s: device %.3f s (null %.3f, ortho %.4f, galerkin %.4f); outer iterations %d "
          "(reference %d)" % (L, blocks, nvecs, secs["total"], secs["null_vectors"], secs["block_orthonormalize"],
                              secs["galerkin"], got["iter"], want["iter"]))
    mg.destroy()
