"""Multigrid set-up on "device" vectors (host/null_gen_dev.cpp, glbx_mg_setup) without a GPU -- SURVEY 8f-2.

Linked against the host-memory mock of the C ABI (tests/mock: serial reductions, reference loop order), the set-up
sequence null_generate_random_smooth_dev -> block_orthonormalize_dev -> generate_coarse_from_fine_stencil_dev must
reproduce the REFERENCE's own set-up (null_gen.cpp:193, mg_complex.cpp:259, :827 driven as in
aa_mg_square_staggered_u1.cpp:716-1143; oracle/ref_mg_shim.cpp refmg_setup) from the same std::mt19937 seed:
the top-level null vectors BIT FOR BIT (same random numbers, same solver, same partition / normalise /
orthogonalise statements), the Galerkin coarse stencil to rounding (the reference assembles it by probing, we sum
P^dag A P directly), and with it the preconditioned solve.  Needs the reference-compiled checker (oracle/_ref)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, load_pkg, rel_err
from mg_common import quiet_stdout

MOCK = os.path.join(ROOT, "tests", "mock", "libglb200_inverters_mock.so")

pytestmark = pytest.mark.skipif("ref" not in oracle_py.available(),
                                reason="the reference's multigrid is only in oracle/_ref/libref_oracle.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    glb = load_pkg()
    L = C.CDLL(MOCK, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    pd = C.POINTER(cd)
    L.glb_op_create_stencil2d.argtypes = [vp, vp, vp, vp, ci, ci, ci, pd, pd, pd, C.POINTER(vp)]
    L.glb_op_stencil_download.argtypes = [vp, vp, vp]
    L.glb_op_get_shifts.argtypes = [vp, pd, pd, pd]
    L.glbx_mg_setup.restype = vp
    L.glbx_mg_setup.argtypes = [vp, ci, ci, ci, C.POINTER(ci), C.POINTER(ci), ci, cd, ci, pd, C.POINTER(ci), ci, ci, ci, ci,
                                C.c_uint, ci, ci, ci, vp]
    L.glbx_mg_level_op.restype = vp
    L.glbx_mg_level_op.argtypes = [vp, ci]
    L.glbx_mg_null_vector.restype = vp
    L.glbx_mg_null_vector.argtypes = [vp, ci, ci]
    L.glbx_mg_destroy.argtypes = [vp]
    L.glbx_mg_set.argtypes = [vp, ci, ci, ci, ci, ci, ci, cd, ci, ci]
    L.glbx_mg_vpgcr.argtypes = [vp, vp, vp, ci, cd, ci, ci, C.POINTER(glb.Result)]
    L.glbx_mg_counts.argtypes = [vp, C.POINTER(ci)]
    L.glbx_force_host_scalars.argtypes = [ci]
    L.glbx_force_host_scalars(1)
    L._glb = glb
    return L


def _dims(X, Y, blocks, nvecs, level):
    dof = 1
    for l in range(level):
        X, Y, dof = X // blocks[l], Y // blocks[l], nvecs[l]
    return X, Y, dof


def _mock_setup(lib, ref, X, Y, blocks, nvecs, **kw):
    """our set-up on the mock, started from the reference's own level-0 stencil (hopping only, mass in the shift)"""
    cl, hp, sh = ref.stencil(0)
    keep = [cl, hp]
    c2 = [(C.c_double * 2)(complex(z).real, complex(z).imag) for z in sh]
    fine = C.c_void_p()
    assert lib.glb_op_create_stencil2d(None, _p(cl), _p(hp), None, X, Y, 1, c2[0], c2[1], c2[2], C.byref(fine)) == 0
    n = len(blocks)
    h = lib.glbx_mg_setup(fine, X, Y, n, (C.c_int * n)(*blocks), (C.c_int * n)(*nvecs), kw.get("bstrat", 1),
                          kw.get("null_mass", 1e-2), lib._glb.Multigrid.SMOOTH[kw.get("null_gen", "BICGSTAB")],
                          (C.c_double * n)(*[kw.get("tol", 5e-5)] * n), (C.c_int * n)(*[kw.get("max_iter", 500)] * n),
                          kw.get("restart_freq", 0), kw.get("bicgstab_l", -1), int(kw.get("do_ortho_eo", False)),
                          int(kw.get("do_global_ortho_conj", False)), kw.get("seed", 1337), 0, kw.get("null_prec", 0),
                          int(kw.get("do_free", False)), _p(ref.links) if kw.get("bstrat", 1) == 3 else None)
    assert h, "glbx_mg_setup failed"
    return h, fine, keep


def _null(lib, h, level, v, size):
    p = lib.glbx_mg_null_vector(h, level, v)
    assert p
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(2 * size,)).view(np.complex128).copy()


def _stencil(lib, h, level, X, Y, nc):
    op = lib.glbx_mg_level_op(h, level)
    cl = np.empty(X * Y * nc * nc, dtype=np.complex128)
    hp = np.empty(4 * X * Y * nc * nc, dtype=np.complex128)
    assert lib.glb_op_stencil_download(op, _p(cl), _p(hp)) == 0
    a = [(C.c_double * 2)() for _ in range(3)]
    assert lib.glb_op_get_shifts(op, a[0], a[1], a[2]) == 0
    return cl, hp, [complex(v[0], v[1]) for v in a]


def _null_counts(lib, h, n_refine):
    n = n_refine + 1
    buf = (C.c_int * (5 * n))()
    lib.glbx_mg_counts(h, buf)
    return [buf[4 * n + l] for l in range(n)]


CASES = [
    dict(L=16, blocks=[4], nvecs=[4], kw=dict(seed=11)),                                    # the driver's defaults, scaled down
    dict(L=16, blocks=[4], nvecs=[8], kw=dict(seed=5, do_ortho_eo=True)),                    # --null-ortho-eo yes
    dict(L=16, blocks=[2], nvecs=[4], kw=dict(seed=3, do_global_ortho_conj=True, null_gen="GCR", max_iter=40)),
    dict(L=24, blocks=[4], nvecs=[3], kw=dict(seed=9, bstrat=0, null_gen="CG", tol=1e-3)),   # BLOCK_NONE, CG smoothing
    dict(L=16, blocks=[4], nvecs=[4], kw=dict(seed=2, null_gen="BICGSTAB_L", bicgstab_l=2, restart_freq=0)),
    # preconditioned null-vector solves (null_gen.cpp:259-313): even/odd system through the stencil (CG: it is
    # Hermitian), and the normal equations through epsilon D epsilon
    dict(L=16, blocks=[4], nvecs=[4], kw=dict(seed=6, null_prec=1, null_gen="CG")),
    dict(L=16, blocks=[4], nvecs=[4], kw=dict(seed=6, null_prec=1, do_ortho_eo=True)),
    dict(L=16, blocks=[2], nvecs=[4], kw=dict(seed=8, null_prec=2, null_gen="CG", tol=1e-3)),
    # BLOCK_CORNER: four vectors per smoothed one (null_gen.cpp:74-88); free-field vectors (null_gen.cpp:162-191)
    dict(L=16, blocks=[4], nvecs=[8], kw=dict(seed=12, bstrat=2, max_iter=60)),
    dict(L=16, blocks=[4], nvecs=[4], kw=dict(seed=12, bstrat=2, do_ortho_eo=True, null_gen="GCR", max_iter=30)),
    dict(L=16, blocks=[4], nvecs=[2], kw=dict(do_free=True)),
    dict(L=16, blocks=[4], nvecs=[4], kw=dict(do_free=True, bstrat=2)),
    # BLOCK_TOPO: the chiral projectors (1 +- Gamma_5)/2 from the symmetric shifts (null_gen.cpp:36-71)
    dict(L=16, blocks=[4], nvecs=[4], kw=dict(seed=14, bstrat=3, max_iter=60)),
    dict(L=16, blocks=[4], nvecs=[2], kw=dict(do_free=True, bstrat=3)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "L%d_b%s_n%s_%s" % (c["L"], c["blocks"], c["nvecs"], c["kw"].get("null_gen", "BICGSTAB") + "_prec%d" % c["kw"].get("null_prec", 0)))
def test_two_level_setup_matches_reference(lib, case):
    L, blocks, nvecs, kw = case["L"], case["blocks"], case["nvecs"], case["kw"]
    orc = oracle_py.load("ref")
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    mass = 0.05
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, mass, blocks, nvecs, **kw)
    h, fine, keep = _mock_setup(lib, ref, L, L, blocks, nvecs, **kw)
    try:
        # null vectors after block_orthonormalize: the same bits
        for v in range(nvecs[0]):
            assert np.array_equal(_null(lib, h, 0, v, L * L), ref.null(0, v)), "null vector %d" % v
        # operator applications spent on the null vectors (dslash_tracker::nullvectors)
        assert _null_counts(lib, h, 1) == ref.null_counts()
        # Galerkin coarse stencil: probing (reference) vs direct sum (ours)
        Xc, Yc, nc = ref.dims(1)
        assert (Xc, Yc, nc) == _dims(L, L, blocks, nvecs, 1)
        cl, hp, sh = _stencil(lib, h, 1, Xc, Yc, nc)
        clr, hpr, shr = ref.stencil(1)
        assert rel_err(cl, clr) < 1e-13 and rel_err(hp, hpr) < 1e-13
        assert sh == [complex(s) for s in shr] == [complex(mass), 0j, 0j]     # the shift copied down, mass restored
        _, _, sh0 = _stencil(lib, h, 0, L, L, 1)
        assert sh0 == [complex(mass), 0j, 0j]
    finally:
        lib.glbx_mg_destroy(h)


def test_three_level_setup_and_solve(lib):
    """Two refinements.  Below the top level the operator is our Galerkin product, equal to the reference's only to
    rounding, and the reference's procedure amplifies rounding enormously there: the smoothed vectors of a small
    coarse lattice are locally almost parallel, so the block Gram-Schmidt divides by tiny numbers (measured here: a
    2e-16 relative change of the level-1 clover moves the level-1 null vectors by 5e-8 after 4 smoothing iterations
    and by 40 % after 20).  So the level-1 logic (colour partition, solver hand-over, counts) is pinned with 4
    smoothing iterations, and the fully smoothed hierarchy is judged by what it is for: the preconditioned solve."""
    L, blocks, nvecs, mass = 32, [4, 2], [4, 4], 0.02
    orc = oracle_py.load("ref")
    rng = orc.rng(7)
    U = rng.gauss_gauge_u1(L, L, 6.0)
    b = rng.gaussian(L * L)
    glb = lib._glb

    # (a) few smoothing iterations below the top level: everything agrees
    kw = dict(seed=21, max_iter=[500, 4])
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, mass, blocks, nvecs, **kw)
    cl, hp, sh = ref.stencil(0)
    c2 = [(C.c_double * 2)(complex(z).real, complex(z).imag) for z in sh]
    fine = C.c_void_p()
    assert lib.glb_op_create_stencil2d(None, _p(cl), _p(hp), None, L, L, 1, c2[0], c2[1], c2[2], C.byref(fine)) == 0
    two = C.c_int * 2
    h = lib.glbx_mg_setup(fine, L, L, 2, two(*blocks), two(*nvecs), 1, 1e-2, glb.Multigrid.SMOOTH["BICGSTAB"],
                          (C.c_double * 2)(5e-5, 5e-5), two(500, 4), 0, -1, 0, 0, 21, 0, 0, 0, None)
    assert h
    try:
        for v in range(nvecs[0]):
            assert np.array_equal(_null(lib, h, 0, v, L * L), ref.null(0, v))
        X1, Y1, d1 = ref.dims(1)
        for v in range(nvecs[1]):
            assert rel_err(_null(lib, h, 1, v, X1 * Y1 * d1), ref.null(1, v)) < 1e-5
        for lvl in (1, 2):
            Xc, Yc, nc = ref.dims(lvl)
            cl1, hp1, sh1 = _stencil(lib, h, lvl, Xc, Yc, nc)
            clr, hpr, shr = ref.stencil(lvl)
            tol = 1e-13 if lvl == 1 else 1e-5
            assert rel_err(cl1, clr) < tol and rel_err(hp1, hpr) < tol
            assert sh1 == [complex(mass), 0j, 0j]
        assert _null_counts(lib, h, 2) == ref.null_counts()
    finally:
        lib.glbx_mg_destroy(h)

    # (b) the driver's defaults on every level: the outer solve VPGCR(64) + V cycle behaves like the reference's
    kw = dict(seed=21)
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, mass, blocks, nvecs, **kw)
    h, fine, keep = _mock_setup(lib, ref, L, L, blocks, nvecs, **kw)
    try:
        for v in range(nvecs[0]):
            assert np.array_equal(_null(lib, h, 0, v, L * L), ref.null(0, v))
        ref.set_precond()
        lib.glbx_mg_set(h, glb.Multigrid.SMOOTH["GCR"], 6, 6, glb.Multigrid.INNER["GCR"], 1024, 64, 1e-2, 0, 1)
        with quiet_stdout():
            xo, want = ref.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
        x = np.zeros_like(b)
        res = glb.Result()
        assert lib.glbx_mg_vpgcr(h, _p(x), _p(b), 1000, 5e-7, 64, 0, C.byref(res)) == 0
        got = res.as_dict()
        assert want["success"] and got["success"]
        assert abs(got["iter"] - want["iter"]) <= max(2, 0.25 * want["iter"]), (got, want)
        assert rel_err(x, xo) < 1e-4
    finally:
        lib.glbx_mg_destroy(h)


@pytest.mark.parametrize("null_prec", [1, 2])
def test_preconditioned_null_solves_below_the_top_level(lib, null_prec):
    """level 1 uses the top/bottom (colour-half) versions: prepare, m^2 - D_tb D_bt, reconstruct (null_gen.cpp:276-286),
    or sigma_3 D sigma_3 and the non-Galerkin normal operator (:304-311).  Few smoothing iterations on level 1 keep the
    rounding amplification of the block Gram-Schmidt small (see test_three_level_setup_and_solve)."""
    L, blocks, nvecs, mass = 32, [4, 2], [4, 4], 0.02
    orc = oracle_py.load("ref")
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, mass, blocks, nvecs, seed=13, max_iter=[500, 3], null_gen="CG",
                                    null_prec=null_prec)
    cl, hp, sh = ref.stencil(0)
    c2 = [(C.c_double * 2)(complex(z).real, complex(z).imag) for z in sh]
    fine = C.c_void_p()
    assert lib.glb_op_create_stencil2d(None, _p(cl), _p(hp), None, L, L, 1, c2[0], c2[1], c2[2], C.byref(fine)) == 0
    two = C.c_int * 2
    h = lib.glbx_mg_setup(fine, L, L, 2, two(*blocks), two(*nvecs), 1, 1e-2, lib._glb.Multigrid.SMOOTH["CG"],
                          (C.c_double * 2)(5e-5, 5e-5), two(500, 3), 0, -1, 0, 0, 13, 0, null_prec, 0, None)
    assert h
    try:
        for v in range(nvecs[0]):
            assert np.array_equal(_null(lib, h, 0, v, L * L), ref.null(0, v))
        X1, Y1, d1 = ref.dims(1)
        for v in range(nvecs[1]):
            assert rel_err(_null(lib, h, 1, v, X1 * Y1 * d1), ref.null(1, v)) < 1e-5
        assert _null_counts(lib, h, 2) == ref.null_counts()
    finally:
        lib.glbx_mg_destroy(h)


def test_coarse_partition_uses_the_reference_colour_period(lib):
    """null_partition_coarse takes the colour index modulo n_vectors[curr_level] (null_gen.cpp:114) -- the number of
    vectors being BUILT on that level, not the dofs per site of that level.  With 4 dofs per site and 8 vectors on
    level 1 that splits by site parity in x instead of by colour half; a drop-in has to do the same."""
    L, blocks, nvecs, mass = 32, [4, 2], [4, 8], 0.02
    orc = oracle_py.load("ref")
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, mass, blocks, nvecs, seed=4, max_iter=[500, 3], do_ortho_eo=True)
    cl, hp, sh = ref.stencil(0)
    c2 = [(C.c_double * 2)(complex(z).real, complex(z).imag) for z in sh]
    fine = C.c_void_p()
    assert lib.glb_op_create_stencil2d(None, _p(cl), _p(hp), None, L, L, 1, c2[0], c2[1], c2[2], C.byref(fine)) == 0
    two = C.c_int * 2
    h = lib.glbx_mg_setup(fine, L, L, 2, two(*blocks), two(*nvecs), 1, 1e-2, lib._glb.Multigrid.SMOOTH["BICGSTAB"],
                          (C.c_double * 2)(5e-5, 5e-5), two(500, 3), 0, -1, 1, 0, 4, 0, 0, 0, None)
    assert h
    try:
        X1, Y1, d1 = ref.dims(1)
        for v in range(nvecs[1]):
            want = ref.null(1, v)
            assert rel_err(_null(lib, h, 1, v, X1 * Y1 * d1), want) < 1e-5
            # the support pattern of the quirk: vectors 0..3 live on index % 8 < 4, vectors 4..7 on the rest
            idx = np.arange(want.size) % 8
            assert np.all(want[(idx >= 4) if v < 4 else (idx < 4)] == 0)
        assert _null_counts(lib, h, 2) == ref.null_counts()
    finally:
        lib.glbx_mg_destroy(h)


def test_unsupported_strategies_fail_loudly(lib):
    """BLOCK_TOPO without the gauge links cannot build its projectors: the set-up refuses instead of doing something else"""
    L = 16
    orc = oracle_py.load("ref")
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    with quiet_stdout():
        ref = oracle_py.RefMg.setup(orc, L, L, U, 0.05, [4], [4], seed=1, max_iter=5)
    cl, hp, sh = ref.stencil(0)
    z = (C.c_double * 2)(0.05, 0.0)
    zero = (C.c_double * 2)(0.0, 0.0)
    fine = C.c_void_p()
    assert lib.glb_op_create_stencil2d(None, _p(cl), _p(hp), None, L, L, 1, z, zero, zero, C.byref(fine)) == 0
    one = (C.c_int * 1)
    saved = os.dup(2)
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 2)
    try:
        h = lib.glbx_mg_setup(fine, L, L, 1, one(4), one(4), 3, 1e-2, 3, (C.c_double * 1)(5e-5), one(5), 0, -1, 0, 0, 1, 0, 0, 0, None)
    finally:
        os.dup2(saved, 2)
        os.close(null)
        os.close(saved)
    assert not h
    a = (C.c_double * 2)()
    assert lib.glb_op_get_shifts(fine, a, None, None) == 0 and a[0] == 0.05    # the caller's shift is restored
