"""Host logic of the C++ solver shells, without a GPU.

The shells (generic-linalg_b200/host/*.cpp) are linked against tests/mock/glb200_mock.cpp -- a
host-memory implementation of the C ABI whose vector kernels are plain loops with the reference's
serial reductions.  On that mock the shells must reproduce the reference's solvers BIT FOR BIT
(solutions, iteration / ops counts, success flags, names, residuals): that pins every scalar
recurrence, stopping test, counting convention and permutation in the host code.  (On the GPU the
only difference is the summation order inside the reductions.)
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, load_pkg, synthetic

MOCK = os.path.join(ROOT, "tests", "mock", "libglb200_inverters_mock.so")


@pytest.fixture(scope="module")
def mock():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    glb = load_pkg()
    lib = C.CDLL(MOCK, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    lib.glbx_host_solve.restype = ci
    lib.glbx_host_solve.argtypes = [ci, C.POINTER(glb.OpDesc), vp, vp, ci, cd, ci, ci, ci, C.POINTER(glb.Result)]
    lib.glbx_host_solve_cg_m.restype = ci
    lib.glbx_host_solve_cg_m.argtypes = [C.POINTER(glb.OpDesc), C.POINTER(vp), vp, ci, ci, ci, cd, vp, ci, ci,
                                         C.POINTER(glb.Result)]
    lib.glbx_host_apply.argtypes = [C.POINTER(glb.OpDesc), vp, vp]
    lib.glbx_force_host_scalars.argtypes = [ci]
    lib.glbx_force_host_scalars(1)  # glb_cg_solve is CUDA-only
    return glb, lib


def desc(glb, kind, X, Y, mass=0.0, Nc=1, links=None):
    d = glb.OpDesc()
    d.kind, d.X, d.Y, d.Nc, d.mass = glb.OP[kind], X, Y, Nc, mass
    d.links = links.ctypes.data_as(C.c_void_p) if links is not None else None
    d._keep = links
    return d


def solve(mock, solver, d, b, x0=None, **kw):
    glb, lib = mock
    x = np.zeros_like(b) if x0 is None else x0.copy()
    res = glb.Result()
    args = dict(max_iter=4000, eps=1e-9, restart_freq=0, l=0)
    args.update(kw)
    rc = lib.glbx_host_solve(glb.SOLVER[solver], C.byref(d), x.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                             args["max_iter"], args["eps"], args["restart_freq"], args["l"], 0, C.byref(res))
    assert rc == 0
    return x, res.as_dict()


from test_oracle_cpu import SOLVES  # the same table that pins the port against the reference


@pytest.mark.parametrize("kind,solver,kw", SOLVES)
def test_shells_bit_identical_to_oracle(mock, kind, solver, kw):
    orc = oracle_py.load("best")
    L = 16
    U, b = synthetic(orc, L)
    op = orc.op(kind, L, L, mass=0.1, links=U)
    bb = b if op.is_complex else np.ascontiguousarray(b.real)
    args = dict(max_iter=4000, eps=1e-9)
    args.update(kw)
    xo, want = orc.solve(solver, op, bb, **args)
    xs, got = solve(mock, solver, desc(mock[0], kind, L, L, mass=0.1, links=U), bb, **args)
    assert got == want
    assert np.array_equal(xs, xo)


@pytest.mark.parametrize("kind", ["STAG_NORMAL_U1", "LAPLACE_REAL"])
def test_multishift_shell_bit_identical(mock, kind):
    glb, lib = mock
    orc = oracle_py.load("best")
    L = 16
    U, b = synthetic(orc, L)
    op = orc.op(kind, L, L, mass=0.1, links=U)
    bb = b if op.is_complex else np.ascontiguousarray(b.real)
    shifts = [0.25, 0.0, 0.05, 0.01]
    xo, want, sho = orc.solve_cg_m(op, bb, shifts, resid_freq_check=3, max_iter=4000, eps=1e-10)
    xs = [np.zeros_like(bb) for _ in shifts]
    ptrs = (C.c_void_p * 4)(*[x.ctypes.data for x in xs])
    sh = np.array(shifts)
    res = glb.Result()
    d = desc(glb, kind, L, L, mass=0.1, links=U)
    assert lib.glbx_host_solve_cg_m(C.byref(d), ptrs, bb.ctypes.data_as(C.c_void_p), 4, 3, 4000, 1e-10,
                                    sh.ctypes.data_as(C.c_void_p), 0, 0, C.byref(res)) == 0
    assert res.as_dict() == want
    assert list(sh) == shifts
    for a, b_ in zip(xs, xo):
        assert np.array_equal(a, b_)


def test_config1_through_shell(mock, golden):
    g = golden["config1_laplace64_cg"]
    N = g["N"]
    b = np.zeros(N * N)
    b[N // 2 + (N // 2) * N] = 1.0
    x0 = np.zeros(N * N)
    x0[N // 2 + (N // 2) * N + 1] = 1.0
    x, got = solve(mock, "CG", desc(mock[0], "LAPLACE_REAL", N, N, mass=g["mass_sq"]), b, x0=x0, max_iter=g["max_iter"],
                   eps=g["tol"])
    assert got == {k: v for k, v in g["result"].items() if k != "x_sha"}
    assert np.array_equal(x, np.load(os.path.join(ROOT, "tests", "golden", "config1_solution.npy")))


def test_unknown_callback_is_refused(mock):
    """no CPU path: a solver handed an operator it cannot map to a device operator reports failure"""
    glb, lib = mock
    d = glb.OpDesc()
    d.kind = 99
    res = glb.Result()
    b = np.zeros(4)
    assert lib.glbx_host_solve(0, C.byref(d), b.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 10, 1e-6, 0, 0, 0,
                               C.byref(res)) != 0
