"""SURVEY 8f-4 shells (multishift CR / BiCGStab, PCG, FPCG, VPGCR, PBiCGStab with the stock preconditioners of
generic_precond.h, the enum dispatch) without a GPU: on the CPU mock of the C ABI they must reproduce the
reference's own solvers BIT FOR BIT -- solutions, iteration / ops counts, success flags, names, residuals.
The reference side is the reference-compiled checker (oracle/_ref); there is no port of this family."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, load_pkg, synthetic

MOCK = os.path.join(ROOT, "tests", "mock", "libglb200_inverters_mock.so")

pytestmark = pytest.mark.skipif("ref" not in oracle_py.available(),
                                reason="this solver family is only in oracle/_ref/libref_oracle.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def mock():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    glb = load_pkg()
    lib = C.CDLL(MOCK, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    lib.glbx_host_solve_multi.argtypes = [ci, C.POINTER(glb.OpDesc), C.POINTER(vp), vp, ci, ci, ci, cd, vp, ci, ci,
                                          C.POINTER(glb.Result)]
    lib.glbx_host_solve_precond.argtypes = [ci, C.POINTER(glb.OpDesc), vp, vp, ci, cd, ci, ci, ci, cd, ci,
                                            C.POINTER(glb.Result)]
    lib.glbx_host_solve_relax.argtypes = [ci, C.POINTER(glb.OpDesc), vp, vp, ci, cd, cd, ci, C.POINTER(glb.Result)]
    lib.glbx_force_host_scalars.argtypes = [ci]
    lib.glbx_force_host_scalars(1)
    return glb, lib


def desc(glb, kind, X, Y, mass=0.0, Nc=1, links=None):
    d = glb.OpDesc()
    d.kind, d.X, d.Y, d.Nc, d.mass = glb.OP[kind], X, Y, Nc, mass
    d.links = links.ctypes.data_as(C.c_void_p) if links is not None else None
    d._keep = links
    return d


@pytest.mark.parametrize("which,kind,shifts,kw", [
    ("CR_M", "STAG_NORMAL_U1", [0.25, 0.0, 0.05, 0.01], dict(resid_freq_check=3)),
    ("CR_M", "LAPLACE_REAL", [0.0, 0.3, 0.1], dict(resid_freq_check=1)),
    ("CR_M", "STAG_NORMAL_U1", [0.0, 0.01, 0.05, 0.25], dict(resid_freq_check=10, worst_first=True)),
    ("BICGSTAB_M", "STAG_U1", [0.25, 0.0, 0.05, 0.01], dict(resid_freq_check=3)),
    ("BICGSTAB_M", "LAPLACE_REAL", [0.0, 0.3, 0.1], dict(resid_freq_check=1)),
    ("BICGSTAB_M", "STAG_U1", [0.0, 0.01, 0.05], dict(resid_freq_check=5, worst_first=True)),
    ("BICGSTAB_M", "STAG_U1", [0.1, 0.2], dict(resid_freq_check=2, max_iter=7)),     # hits max_iter
    ("CG_M", "STAG_NORMAL_U1", [0.25, 0.0, 0.05, 0.01], dict(resid_freq_check=3)),
])
def test_multishift_family_bit_identical(mock, which, kind, shifts, kw):
    glb, lib = mock
    orc = oracle_py.load("ref")
    L = 16
    U, b = synthetic(orc, L)
    op = orc.op(kind, L, L, mass=0.1, links=U)
    bb = b if op.is_complex else np.ascontiguousarray(b.real)
    args = dict(resid_freq_check=10, max_iter=4000, eps=1e-10, worst_first=False)
    args.update(kw)
    xo, want, sho = oracle_py.ref_solve_multi(orc, which, op, bb, shifts, **args)
    xs = [np.zeros_like(bb) for _ in shifts]
    n = len(shifts)
    ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
    sh = np.array(shifts, dtype=np.float64)
    res = glb.Result()
    d = desc(glb, kind, L, L, mass=0.1, links=U)
    assert lib.glbx_host_solve_multi(oracle_py.MULTI[which], C.byref(d), ptrs, _p(bb), n, args["resid_freq_check"],
                                     args["max_iter"], args["eps"], _p(sh), int(args["worst_first"]), 0,
                                     C.byref(res)) == 0
    assert res.as_dict() == want
    assert list(sh) == list(sho) == shifts
    for a, b_ in zip(xs, xo):
        assert np.array_equal(a, b_)


PRECOND_CASES = [
    ("PCG", "STAG_NORMAL_U1", dict(precond="IDENTITY")),
    ("PCG", "STAG_NORMAL_U1", dict(precond="GCR", n_step=3)),
    ("PCG", "LAPLACE_REAL", dict(precond="GCR", n_step=2)),
    ("FPCG", "STAG_NORMAL_U1", dict(precond="GCR", n_step=3)),
    ("FPCG", "LAPLACE_REAL", dict(precond="IDENTITY")),
    ("FPCG_RESTART", "STAG_NORMAL_U1", dict(precond="GCR", n_step=2, restart_freq=5)),
    ("FPCG_RESTART", "LAPLACE_REAL", dict(precond="GCR", n_step=2, restart_freq=4)),
    ("VPGCR", "STAG_U1", dict(precond="GCR", n_step=4)),
    ("VPGCR", "LAPLACE_REAL", dict(precond="IDENTITY")),
    ("VPGCR_RESTART", "STAG_U1", dict(precond="GCR", n_step=3, restart_freq=6)),
    ("PBICGSTAB", "STAG_U1", dict(precond="GCR", n_step=3)),
    ("PBICGSTAB", "LAPLACE_REAL", dict(precond="IDENTITY")),
    ("PBICGSTAB", "STAG_U1", dict(precond="IDENTITY", max_iter=9)),                  # fails: k == max_iter quirk
    ("PBICGSTAB_RESTART", "STAG_U1", dict(precond="GCR", n_step=2, restart_freq=7)),
    ("VPGCR", "STAG_U1", dict(precond="MINRES", n_step=4)),                          # minres_preconditioner
    ("FPCG", "STAG_NORMAL_U1", dict(precond="MINRES", n_step=3)),
    ("VPGCR_RESTART", "LAPLACE_REAL", dict(precond="MINRES", n_step=2, restart_freq=6)),
]


@pytest.mark.parametrize("solver,kind,kw", PRECOND_CASES)
def test_preconditioned_family_bit_identical(mock, solver, kind, kw):
    glb, lib = mock
    orc = oracle_py.load("ref")
    L = 16
    U, b = synthetic(orc, L)
    op = orc.op(kind, L, L, mass=0.1, links=U)
    bb = b if op.is_complex else np.ascontiguousarray(b.real)
    args = dict(max_iter=4000, eps=1e-9, restart_freq=0, precond="IDENTITY", n_step=4, rel_res=1e-20)
    args.update(kw)
    xo, want = oracle_py.ref_solve_precond(orc, solver, op, bb, **args)
    x = np.zeros_like(bb)
    res = glb.Result()
    d = desc(glb, kind, L, L, mass=0.1, links=U)
    assert lib.glbx_host_solve_precond(oracle_py.PRECOND_SOLVER[solver], C.byref(d), _p(x), _p(bb), args["max_iter"],
                                       args["eps"], args["restart_freq"], oracle_py.PRECOND[args["precond"]],
                                       args["n_step"], args["rel_res"], 0, C.byref(res)) == 0
    assert res.as_dict() == want
    assert np.array_equal(x, xo)


RELAX_CASES = [
    # SOR only converges when |1 - omega lambda| < 1 for the whole spectrum: the (massive) Laplacians qualify
    ("SOR", "LAPLACE_REAL", dict(omega=0.2, eps=1e-6)),
    ("SOR", "LAPLACE_NC", dict(omega=0.15, eps=1e-5)),                       # complex overload: iter is not bumped
    ("SOR", "LAPLACE_REAL", dict(omega=0.2, eps=1e-12, max_iter=25)),         # hits max_iter
    ("SOR", "STAG_U1", dict(omega=0.3, eps=1e-8, max_iter=12)),               # diverges; the shell must follow it
    ("MINRES", "LAPLACE_REAL", dict(omega=1.0, eps=1e-8)),
    ("MINRES", "STAG_U1", dict(omega=1.0, eps=1e-7)),
    ("MINRES", "STAG_U1", dict(omega=0.67, eps=1e-7)),                        # the multigrid smoother's relaxation
    ("MINRES", "STAG_NORMAL_U1", dict(omega=0.85, eps=1e-9, max_iter=40)),    # hits max_iter
]


@pytest.mark.parametrize("which,kind,kw", RELAX_CASES)
def test_sor_minres_bit_identical(mock, which, kind, kw):
    """minv_vector_sor / minv_vector_minres (generic_sor.cpp, generic_minres.cpp): solution, counts, success flag,
    name (it carries omega) and residual equal the reference's, including the overload differences (the real
    versions count the last iteration, the complex ones do not) and a non-zero initial guess"""
    glb, lib = mock
    orc = oracle_py.load("ref")
    L = 16
    U, b = synthetic(orc, L)
    Nc = 2 if kind == "LAPLACE_NC" else 1
    op = orc.op(kind, L, L, mass=0.1, links=U, Nc=Nc) if Nc > 1 else orc.op(kind, L, L, mass=0.1, links=U)
    rg = np.random.default_rng(3)
    n = op.size
    bb = (rg.standard_normal(n) + 1j * rg.standard_normal(n)) if op.is_complex else rg.standard_normal(n)
    x0 = 0.1 * ((rg.standard_normal(n) + 1j * rg.standard_normal(n)) if op.is_complex else rg.standard_normal(n))
    args = dict(max_iter=3000, eps=1e-8, omega=1.0)
    args.update(kw)
    xo, want = oracle_py.ref_solve_relax(orc, which, op, bb, x0=x0, **args)
    x = np.array(x0, copy=True)
    res = glb.Result()
    d = desc(glb, kind, L, L, mass=0.1, Nc=Nc, links=U)
    assert lib.glbx_host_solve_relax(dict(SOR=0, MINRES=1)[which], C.byref(d), _p(x), _p(bb), args["max_iter"],
                                     args["eps"], args["omega"], 0, C.byref(res)) == 0
    assert res.as_dict() == want
    assert np.array_equal(x, xo)
