"""Even/odd preconditioning of the staggered operator (operators.cpp:456-616) without a GPU:
the port against the reference-compiled checker (bit equality), the algebraic identities the reference's
tests/staggered_pieces checks, and the host shells on the CPU mock of the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, load_pkg, synthetic

MOCK = os.path.join(ROOT, "tests", "mock", "libglb200_inverters_mock.so")
KINDS = ["STAG_DEO_U1", "STAG_DOE_U1", "STAG_M2MDEODOE_U1"]


def parity_masks(L):
    i = np.arange(L * L)
    even = ((i % L + i // L) % 2) == 0
    return even, ~even


@pytest.mark.skipif("ref" not in oracle_py.available(), reason="needs oracle/_ref")
@pytest.mark.parametrize("L", [8, 12])
def test_port_equals_reference(L):
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    U, b = synthetic(ref, L)
    w = ref.rng(9).gaussian(L * L)
    for k in KINDS:
        assert np.array_equal(ref.op(k, L, L, mass=0.13, links=U).apply(b), port.op(k, L, L, mass=0.13, links=U).apply(b))
    ro, po = ref.op("STAG_U1", L, L, mass=0.13, links=U), port.op("STAG_U1", L, L, mass=0.13, links=U)
    assert np.array_equal(ro.eoprec_prepare(b), po.eoprec_prepare(b))
    assert np.array_equal(ro.eoprec_reconstruct(b, w), po.eoprec_reconstruct(b, w))


def test_even_odd_identities(orc):
    """D = m + D_eo + D_oe ; (m^2 - D_eo D_oe) is Hermitian on the even sublattice ; prepare / solve / reconstruct
    reproduces D^-1 b  (tests/staggered_pieces/staggered_pieces.cpp TEST 5-8)"""
    L, m = 12, 0.2
    U, b = synthetic(orc, L)
    even, odd = parity_masks(L)
    D = orc.op("STAG_U1", L, L, mass=m, links=U)
    Deo, Doe = orc.op("STAG_DEO_U1", L, L, mass=m, links=U), orc.op("STAG_DOE_U1", L, L, mass=m, links=U)
    M = orc.op("STAG_M2MDEODOE_U1", L, L, mass=m, links=U)
    assert np.allclose(D.apply(b), m * b + Deo.apply(b) + Doe.apply(b), rtol=0, atol=1e-14)
    assert np.all(Deo.apply(b)[odd] == 0) and np.all(Doe.apply(b)[even] == 0) and np.all(M.apply(b)[odd] == 0)
    v, w = np.where(even, b, 0), np.where(even, orc.rng(3).gaussian(L * L), 0)
    assert abs(np.vdot(w, M.apply(v)) - np.vdot(M.apply(w), v)) < 1e-12
    be = D.eoprec_prepare(b)
    xe, info = orc.solve("CG", M, be, max_iter=4000, eps=1e-12)
    x = D.eoprec_reconstruct(xe, b)
    assert info["success"] and np.linalg.norm(D.apply(x) - b) / np.linalg.norm(b) < 1e-10
    # D^dag D = m^2 - D_eo D_oe - D_oe D_eo is block diagonal in parity: the e/o system needs no more iterations than
    # CGNE (its gain is that only the even half carries information)
    N = orc.op("STAG_NORMAL_U1", L, L, mass=m, links=U)
    _, plain = orc.solve("CG", N, orc.op("STAG_DAGGER_U1", L, L, mass=m, links=U).apply(b), max_iter=4000, eps=1e-12)
    assert info["iter"] <= plain["iter"]


def test_shells_on_mock_bit_identical():
    """host entry points (operators.h: square_staggered_{deo,doe,m2mdeodoe}_u1, eoprec_prepare / _reconstruct) and a
    CG solve of the e/o system on the CPU mock equal the oracle bit for bit"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    glb = load_pkg()
    lib = C.CDLL(MOCK, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    lib.glbx_host_apply.argtypes = [C.POINTER(glb.OpDesc), vp, vp]
    lib.glbx_host_eoprec_prepare.argtypes = [C.POINTER(glb.OpDesc), vp, vp]
    lib.glbx_host_eoprec_reconstruct.argtypes = [C.POINTER(glb.OpDesc), vp, vp, vp]
    lib.glbx_host_solve.argtypes = [ci, C.POINTER(glb.OpDesc), vp, vp, ci, cd, ci, ci, ci, C.POINTER(glb.Result)]
    lib.glbx_force_host_scalars.argtypes = [ci]
    lib.glbx_force_host_scalars(1)
    orc = oracle_py.load("best")
    L, m = 16, 0.1
    U, b = synthetic(orc, L)
    p = lambda a: a.ctypes.data_as(vp)

    def desc(kind):
        d = glb.OpDesc()
        d.kind, d.X, d.Y, d.Nc, d.mass = glb.OP[kind], L, L, 1, m
        d.links = p(U)
        return d
    for k in KINDS:
        out = np.empty_like(b)
        assert lib.glbx_host_apply(C.byref(desc(k)), p(out), p(b)) == 0
        assert np.array_equal(out, orc.op(k, L, L, mass=m, links=U).apply(b))
    D = orc.op("STAG_U1", L, L, mass=m, links=U)
    be = np.empty_like(b)
    assert lib.glbx_host_eoprec_prepare(C.byref(desc("STAG_U1")), p(be), p(b)) == 0
    assert np.array_equal(be, D.eoprec_prepare(b))
    xe = np.zeros_like(b)
    res = glb.Result()
    assert lib.glbx_host_solve(glb.SOLVER["CG"], C.byref(desc("STAG_M2MDEODOE_U1")), p(xe), p(be), 4000, 1e-10, 0, 0, 0,
                               C.byref(res)) == 0
    xo, want = orc.solve("CG", orc.op("STAG_M2MDEODOE_U1", L, L, mass=m, links=U), be, max_iter=4000, eps=1e-10)
    assert res.as_dict() == want and np.array_equal(xe, xo)
    x = np.empty_like(b)
    assert lib.glbx_host_eoprec_reconstruct(C.byref(desc("STAG_U1")), p(x), p(xe), p(b)) == 0
    assert np.array_equal(x, D.eoprec_reconstruct(xo, b))


@pytest.mark.skipif("ref" not in oracle_py.available(), reason="needs oracle/_ref")
@pytest.mark.parametrize("X,Y,nc,two", [(6, 8, 4, False), (8, 6, 2, True), (5, 7, 3, True), (4, 4, 1, False), (8, 8, 8, False)])
def test_partial_stencil_applies_port_equals_reference(X, Y, nc, two):
    """apply_stencil_2d_{eo,oe,tb,bt} (coarse_stencil.cpp:395-1512): port == reference, bit for bit"""
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    rg = np.random.default_rng(X * 100 + nc)
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    V = X * Y
    cl, hp, tl, v = rc(V * nc * nc), rc(4 * V * nc * nc), (rc(8 * V * nc * nc) if two else None), rc(V * nc)
    kw = dict(Nc=nc, clover=cl, hopping=hp, two_link=tl, shift=0.3, eo_shift=0.1j, dof_shift=0.2)
    a, b = ref.op("STENCIL", X, Y, **kw), port.op("STENCIL", X, Y, **kw)
    for part in ("EO", "OE", "TB", "BT"):
        assert np.array_equal(a.apply_part(part, v), b.apply_part(part, v))


def test_partial_stencil_identities(orc):
    """on the staggered stencil: D = m + D_eo + D_oe, and the function operators' even/odd pieces equal the stencil's
    (tests/staggered_pieces/staggered_pieces.cpp TEST 9-10); on any stencil the four colour blocks add up:
    (tt + bb) + tb + bt = clover + hopping, checked through linearity"""
    L, m = 8, 0.2
    U, b = synthetic(orc, L)
    S = orc.op("STENCIL_FROM_STAG", L, L, mass=m, links=U)
    assert np.allclose(S.apply(b), m * b + S.apply_part("EO", b) + S.apply_part("OE", b), rtol=0, atol=1e-14)
    assert np.allclose(S.apply_part("EO", b), orc.op("STAG_DEO_U1", L, L, mass=m, links=U).apply(b), rtol=0, atol=1e-15)
    assert np.allclose(S.apply_part("OE", b), orc.op("STAG_DOE_U1", L, L, mass=m, links=U).apply(b), rtol=0, atol=1e-15)
    rg = np.random.default_rng(5)
    nc, V = 4, L * L
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    T = orc.op("STENCIL", L, L, Nc=nc, clover=rc(V * nc * nc), hopping=rc(4 * V * nc * nc))
    v = rc(V * nc)
    top = (np.arange(V * nc) % nc) < nc // 2
    vt, vb = np.where(top, v, 0), np.where(~top, v, 0)
    full = T.apply(v)
    # rows of the top half: T v = (T vt)_top + tb(v);  rows of the bottom half likewise with bt
    assert np.allclose(np.where(top, full, 0), np.where(top, T.apply(vt), 0) + T.apply_part("TB", v), rtol=0, atol=1e-13)
    assert np.allclose(np.where(~top, full, 0), np.where(~top, T.apply(vb), 0) + T.apply_part("BT", v), rtol=0, atol=1e-13)


def test_partial_stencil_host_entry_points_on_mock():
    """apply_stencil_2d_{eo,oe,tb,bt} of host/coarse_stencil.h on the CPU mock == the oracle"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    glb = load_pkg()
    lib = C.CDLL(MOCK, mode=C.RTLD_LOCAL)
    vp, ci = C.c_void_p, C.c_int
    lib.glbx_host_stencil_part.argtypes = [C.POINTER(glb.OpDesc), ci, vp, vp]
    orc = oracle_py.load("best")
    X, Y, nc = 6, 8, 4
    V = X * Y
    rg = np.random.default_rng(8)
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    cl, hp, tl, v = rc(V * nc * nc), rc(4 * V * nc * nc), rc(8 * V * nc * nc), rc(V * nc)
    p = lambda a: a.ctypes.data_as(vp)
    d = glb.OpDesc()
    d.kind, d.X, d.Y, d.Nc, d.mass = glb.OP["STENCIL"], X, Y, nc, 0.0
    d.clover, d.hopping, d.two_link, d.has_two = p(cl), p(hp), p(tl), 1
    oop = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, two_link=tl)
    for part, code in (("EO", 1), ("OE", 2), ("TB", 3), ("BT", 4)):
        out = np.empty_like(v)
        assert lib.glbx_host_stencil_part(C.byref(d), code, p(out), p(v)) == 0
        assert np.array_equal(out, oop.apply_part(part, v))
