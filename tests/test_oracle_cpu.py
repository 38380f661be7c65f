"""Pins the CPU oracle (no GPU needed).

1. both checkers reproduce the golden numbers generated from the unmodified reference
   (oracle/gen_golden.py; the same numbers BASELINE.md records from the reference's own binaries:
   unit_test.cpp iteration/ops/residual triplets, the 150-iteration config-1 solve, the five
   values printed by tests/staggered_stencil, iteration counts on the seeded synthetic inputs);
2. when oracle/_ref (the reference compiled from /root/reference) is present, the port agrees
   with it BIT FOR BIT on operators, BLAS-1, inputs and complete solves.
"""
import os

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, synthetic

KINDS = oracle_py.available()
pytestmark = pytest.mark.skipif(not KINDS, reason="no oracle library built")


@pytest.fixture(scope="module", params=KINDS)
def any_orc(request):
    return oracle_py.load(request.param)


def test_unit_test_triplets(any_orc, golden):
    g = golden["unit_test_128"]
    N = g["N"]
    b = np.zeros(N * N)
    b[g["src_index"]] = 1.0
    op = any_orc.op("LAPLACE_REAL", N, N, mass=g["mass_sq"])
    for name, want in g["results"].items():
        call = dict(want["call"])
        solver = call.pop("solver")
        x, got = any_orc.solve(solver, op, b, x0=b.copy(), eps=g["tol"], **call)
        assert (got["iter"], got["ops_count"], got["success"], got["name"]) == \
            (want["iter"], want["ops_count"], want["success"], want["name"]), name
        assert got["resSq"] == want["resSq"], name  # bit-identical residual


def test_unit_test_matches_published_numbers(golden):
    """the numbers BASELINE.md section 2 records from the reference's own unit_test binary"""
    r = golden["unit_test_128"]["results"]
    assert (r["CG"]["iter"], r["CG"]["ops_count"]) == (174, 176)
    assert abs(r["CG"]["resSq"] ** 0.5 - 9.870329542587672e-07) < 1e-21
    assert (r["CR"]["iter"], r["GCR"]["iter"], r["GMRES"]["iter"], r["GMRES"]["ops_count"]) == (164, 164, 164, 329)
    assert (r["BiCGStab"]["iter"], r["BiCGStab"]["ops_count"]) == (110, 222)
    assert (r["BiCGStab-2"]["iter"], r["BiCGStab-8"]["iter"]) == (111, 113)
    assert (r["CG(8)"]["iter"], r["GCR(8)"]["iter"], r["GMRES(8)"]["ops_count"]) == (516, 481, 1023)


def test_config1_cg_150_iterations(any_orc, golden):
    g = golden["config1_laplace64_cg"]
    N = g["N"]
    b = np.zeros(N * N)
    b[N // 2 + (N // 2) * N] = 1.0
    x0 = np.zeros(N * N)
    x0[N // 2 + (N // 2) * N + 1] = 1.0
    op = any_orc.op("LAPLACE_REAL", N, N, mass=g["mass_sq"])
    x, got = any_orc.solve("CG", op, b, x0=x0, max_iter=g["max_iter"], eps=g["tol"])
    assert got["iter"] == 150 == g["result"]["iter"]
    assert got["resSq"] == g["result"]["resSq"]
    assert np.array_equal(x, np.load(os.path.join(ROOT, "tests", "golden", "config1_solution.npy")))


def _cfg_links(L):
    """u1_utils.cpp:17-33: file order x outer, y, mu inner; stored as lattice[y*2L + 2x + mu]"""
    ph = np.load(os.path.join(ROOT, "tests", "golden", "l64t64b60_heatbath_phases.npy")).reshape(L, L, 2)
    return np.ascontiguousarray(np.exp(1j * ph.transpose(1, 0, 2)).reshape(-1))


def test_staggered_stencil_fixture(any_orc, golden):
    g = golden["staggered_stencil_64"]
    L = g["L"]
    U = _cfg_links(L)
    plaq = any_orc.plaquette(U, L, L)
    assert abs(plaq.real - 9.184146e-01) < 1e-6 and abs(plaq.imag - 4.699586e-04) < 1e-9
    src = np.zeros(L * L, dtype=np.complex128)
    src[g["src_index"]] = 1.0
    fn = any_orc.op("STAG_U1", L, L, mass=g["mass"], links=U).apply(src)
    st = any_orc.op("STENCIL_FROM_STAG", L, L, mass=g["mass"], links=U).apply(src)
    assert any_orc.diffnorm2sq(fn, st) == 0.0  # staggered_stencil.cpp:232 prints exactly 0
    nz = np.flatnonzero(fn)
    assert list(nz) == g["nonzero_index"]
    # polar(1,theta) vs exp(i theta) may differ in the last bit: compare to 1e-15
    assert np.allclose(fn[nz].real, g["nonzero_re"], rtol=0, atol=1e-15)
    assert np.allclose(fn[nz].imag, g["nonzero_im"], rtol=0, atol=1e-15)
    # the five values the reference test prints (staggered_stencil.cpp:220)
    by = dict(zip(nz, fn[nz]))
    assert abs(by[L + 1] - 1e-2) < 1e-16
    assert abs(by[L] - complex(-2.979263e-01, 4.015469e-01)) < 1e-6      # -x neighbour sees +x hop... 
    assert abs(by[L + 2] - complex(-4.456911e-01, 2.266262e-01)) < 1e-6
    assert abs(by[1] - complex(2.430415e-01, -4.369564e-01)) < 1e-6
    assert abs(by[2 * L + 1] - complex(-2.735168e-01, 4.185553e-01)) < 1e-6


@pytest.mark.parametrize("L", [64, 256])
def test_synthetic_iteration_counts(any_orc, golden, L):
    g = golden["synthetic_beta6_m0.1"][str(L)]
    U, b = synthetic(any_orc, L)
    assert abs(any_orc.plaquette(U, L, L).real - g["plaquette"]) < 1e-15
    D = any_orc.op("STAG_U1", L, L, mass=0.1, links=U)
    DdD = any_orc.op("STAG_NORMAL_U1", L, L, mass=0.1, links=U)
    bp = any_orc.op("STAG_DAGGER_U1", L, L, mass=0.1, links=U).apply(b)
    checks = [("CGNE", lambda: any_orc.solve("CG", DdD, bp, max_iter=100000, eps=1e-10)[1]),
              ("BiCGStab", lambda: any_orc.solve("BICGSTAB", D, b, max_iter=100000, eps=1e-10)[1])]
    if L == 64:  # the long-recurrence ones only at the small size (CPU suite stays within minutes)
        checks += [("GMRES(20)", lambda: any_orc.solve("GMRES_RESTART", D, b, max_iter=100000, eps=1e-8,
                                                       restart_freq=20)[1]),
                   ("GCR(20)", lambda: any_orc.solve("GCR_RESTART", D, b, max_iter=100000, eps=1e-8,
                                                     restart_freq=20)[1]),
                   ("CG-M", lambda: any_orc.solve_cg_m(DdD, bp, [0.0, 0.01, 0.05, 0.25], max_iter=100000,
                                                       eps=1e-10)[1])]
    for name, run in checks:
        got, want = run(), g[name]
        assert (got["iter"], got["ops_count"], got["success"]) == (want["iter"], want["ops_count"], want["success"]), name
        assert got["resSq"] == want["resSq"], name


def test_published_synthetic_counts(golden):
    """BASELINE.md section 2, 256^2 and 64^2 rows"""
    g = golden["synthetic_beta6_m0.1"]
    assert (g["256"]["CGNE"]["iter"], g["256"]["CGNE"]["ops_count"]) == (168, 170)
    assert (g["256"]["BiCGStab"]["iter"], g["256"]["BiCGStab"]["ops_count"]) == (314, 630)
    assert (g["256"]["GMRES(20)"]["iter"], g["256"]["GMRES(20)"]["ops_count"]) == (269, 581)
    assert (g["256"]["GCR(20)"]["iter"], g["256"]["GCR(20)"]["ops_count"]) == (285, 315)
    assert (g["256"]["CG-M"]["iter"], g["256"]["CG-M"]["ops_count"]) == (171, 175)
    assert [g["64"][k]["iter"] for k in ("CGNE", "BiCGStab", "GMRES(20)", "GCR(20)", "CG-M")] == [163, 294, 264, 277, 171]


# ------------------------------------------------------------------ port == reference, bit for bit
both = pytest.mark.skipif(set(KINDS) != {"ref", "port"}, reason="needs both oracle/_ref and the port")


@both
def test_port_inputs_bit_identical():
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    for L in (8, 32):
        Ur, br = synthetic(ref, L)
        Up, bp = synthetic(port, L)
        assert np.array_equal(Ur, Up) and np.array_equal(br, bp)
    a, b = ref.rng(5).gaussian(100, np.float64), port.rng(5).gaussian(100, np.float64)
    assert np.array_equal(a, b)


@both
@pytest.mark.parametrize("L", [4, 6, 32])
def test_port_operators_bit_identical(L):
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    U, _ = synthetic(ref, L)
    for kind in ["LAPLACE_IMAG", "LAPLACE_NC", "LAPLACE_U1", "STAG_FREE", "STAG_U1", "STAG_GAMMA5_U1",
                 "STAG_GAMMA5_FREE", "STAG_DAGGER_U1", "STAG_NORMAL_U1", "GAMMA5", "STENCIL_FROM_STAG",
                 "LAPLACE_REAL", "LAPLACE_REAL_NC", "STAG_FREE_REAL"]:
        kw = dict(mass=0.1, links=U)
        if kind in ("LAPLACE_NC", "LAPLACE_REAL_NC"):
            kw["Nc"] = 3
        a, b = ref.op(kind, L, L, **kw), port.op(kind, L, L, **kw)
        v = ref.rng(7).gaussian(a.size, a.dtype)
        assert np.array_equal(a.apply(v), b.apply(v)), kind


@both
@pytest.mark.parametrize("nc,two", [(1, False), (2, True), (4, False), (8, False), (3, True)])
def test_port_coarse_stencil_bit_identical(nc, two):
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    X, Y = 6, 8
    V = X * Y
    rg = np.random.default_rng(nc)
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    cl, hp, tl = rc(V * nc * nc), rc(4 * V * nc * nc), (rc(8 * V * nc * nc) if two else None)
    kw = dict(Nc=nc, clover=cl, hopping=hp, two_link=tl, shift=0.3 + 0.1j, eo_shift=0.2 - 0.5j, dof_shift=0.7j)
    v = rc(V * nc)
    assert np.array_equal(ref.op("STENCIL", X, Y, **kw).apply(v), port.op("STENCIL", X, Y, **kw).apply(v))


@both
def test_port_blas_bit_identical():
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    rg = np.random.default_rng(0)
    for dt in (np.float64, np.complex128):
        x = rg.standard_normal(1001).astype(dt)
        y = rg.standard_normal(1001).astype(dt)
        if dt == np.complex128:
            x = x + 1j * rg.standard_normal(1001)
            y = y - 1j * rg.standard_normal(1001)
        assert ref.dot(x, y) == port.dot(x, y)
        assert ref.norm2sq(x) == port.norm2sq(x)
        assert ref.diffnorm2sq(x, y) == port.diffnorm2sq(x, y)


SOLVES = [("STAG_NORMAL_U1", "CG", {}), ("STAG_NORMAL_U1", "CG_RESTART", dict(restart_freq=32)),
          ("STAG_NORMAL_U1", "CR", {}), ("STAG_NORMAL_U1", "CR_RESTART", dict(restart_freq=32)),
          ("STAG_U1", "GCR", dict(max_iter=300)), ("STAG_U1", "GCR_RESTART", dict(restart_freq=20)),
          ("STAG_U1", "BICGSTAB", {}), ("STAG_U1", "BICGSTAB_RESTART", dict(restart_freq=20)),
          ("STAG_U1", "BICGSTAB_L", dict(l=4)), ("STAG_U1", "BICGSTAB_L_RESTART", dict(l=2, restart_freq=20)),
          ("STAG_U1", "GMRES", dict(max_iter=120)), ("STAG_U1", "GMRES_RESTART", dict(restart_freq=20)),
          ("LAPLACE_REAL", "CG", {}), ("LAPLACE_REAL", "CR", {}), ("LAPLACE_REAL", "GCR", {}),
          ("LAPLACE_REAL", "BICGSTAB", {}), ("LAPLACE_REAL", "BICGSTAB_L", dict(l=2)),
          ("LAPLACE_REAL", "GMRES", {}), ("LAPLACE_REAL", "GMRES_RESTART", dict(restart_freq=8)),
          # hitting max_iter exercises the success/iter quirks (CR complex never fails; GMRES double decrement)
          ("STAG_U1", "CG", dict(max_iter=7)), ("STAG_U1", "CR", dict(max_iter=7)),
          ("STAG_U1", "GCR", dict(max_iter=7)), ("STAG_U1", "BICGSTAB", dict(max_iter=7)),
          ("STAG_U1", "BICGSTAB_L", dict(max_iter=7, l=2)), ("STAG_U1", "GMRES", dict(max_iter=7)),
          ("LAPLACE_REAL", "CR", dict(max_iter=7)), ("LAPLACE_REAL", "GMRES", dict(max_iter=7)),
          ("STAG_FREE_REAL", "BICGSTAB", {}), ("STAG_FREE_REAL", "GCR", dict(max_iter=200))]


@both
@pytest.mark.parametrize("kind,solver,kw", SOLVES)
def test_port_solvers_bit_identical(kind, solver, kw):
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    L = 16
    U, b = synthetic(ref, L)
    out = []
    for o in (ref, port):
        op = o.op(kind, L, L, mass=0.1, links=U)
        bb = b if op.is_complex else np.ascontiguousarray(b.real)
        args = dict(max_iter=4000, eps=1e-9)
        args.update(kw)
        out.append(o.solve(solver, op, bb, **args))
    assert out[0][1] == out[1][1]
    assert np.array_equal(out[0][0], out[1][0])


@both
@pytest.mark.parametrize("kind", ["STAG_NORMAL_U1", "LAPLACE_REAL"])
def test_port_multishift_bit_identical(kind):
    ref, port = oracle_py.load("ref"), oracle_py.load("port")
    L = 16
    U, b = synthetic(ref, L)
    out = []
    for o in (ref, port):
        op = o.op(kind, L, L, mass=0.1, links=U)
        bb = b if op.is_complex else np.ascontiguousarray(b.real)
        out.append(o.solve_cg_m(op, bb, [0.25, 0.0, 0.05, 0.01], resid_freq_check=3, max_iter=4000, eps=1e-10))
    assert out[0][1] == out[1][1]
    assert np.array_equal(out[0][2], out[1][2])  # shifts restored to caller order
    for a, b_ in zip(out[0][0], out[1][0]):
        assert np.array_equal(a, b_)
