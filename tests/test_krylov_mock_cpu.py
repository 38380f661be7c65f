"""The device-loop branches of the CG / BiCGStab / CR shells (host/dev_solvers.cpp: glb_cg_solve, glb_krylov_solve,
history replay for VERB_DETAIL, success flags, operator counts) without a GPU.

tests/mock/glb200_mock.cpp restates glb_krylov_solve (csrc/krylov.cu) and the two-kernel glb_cg_solve (csrc/cg.cu) on the host: the same sequence of vector
operations and the same scalar formulas the CUDA loop evaluates in its kernel prologues / epilogues, with the mock's
serial reductions.  Through it the branch a GPU run takes must reproduce the REFERENCE bit for bit -- solution,
iteration / operator counts, success flags, residual -- and print, at VERB_DETAIL, exactly the lines the host-scalar
shell prints (the reference's lines: tests/test_reference_programs_cpu.py compares those with the reference build).
Says nothing about the CUDA kernels; their parity is tests/test_krylov_gpu.py.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, load_pkg, synthetic
from test_oracle_cpu import SOLVES

MOCK = os.path.join(ROOT, "tests", "mock", "libglb200_inverters_mock.so")
KRYLOV = ("CG", "CG_RESTART", "CR", "CR_RESTART", "BICGSTAB", "BICGSTAB_RESTART")


@pytest.fixture(scope="module")
def mock():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    glb = load_pkg()
    lib = C.CDLL(MOCK, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    lib.glbx_host_solve.restype = ci
    lib.glbx_host_solve.argtypes = [ci, C.POINTER(glb.OpDesc), vp, vp, ci, cd, ci, ci, ci, C.POINTER(glb.Result)]
    lib.glbx_force_host_scalars.argtypes = [ci]
    lib.glb200_mock_krylov_calls.restype = C.c_ulonglong
    lib.glbx_force_host_scalars(0)
    return glb, lib


def desc(glb, kind, X, Y, mass=0.0, links=None):
    d = glb.OpDesc()
    d.kind, d.X, d.Y, d.Nc, d.mass = glb.OP[kind], X, Y, 1, mass
    d.links = links.ctypes.data_as(C.c_void_p) if links is not None else None
    d._keep = links
    return d


def solve(mock, solver, d, b, force_shell, verbosity=0, **kw):
    glb, lib = mock
    x = np.zeros_like(b)
    res = glb.Result()
    args = dict(max_iter=4000, eps=1e-9, restart_freq=0, l=0)
    args.update(kw)
    lib.glbx_force_host_scalars(1 if force_shell else 0)
    try:
        rc = lib.glbx_host_solve(glb.SOLVER[solver], C.byref(d), x.ctypes.data_as(C.c_void_p),
                                 b.ctypes.data_as(C.c_void_p), args["max_iter"], args["eps"], args["restart_freq"],
                                 args["l"], verbosity, C.byref(res))
    finally:
        lib.glbx_force_host_scalars(0)
    assert rc == 0
    return x, res.as_dict()


def printed(fn):
    """what the C++ side writes to stdout (std::cout << ... << std::endl) during fn()"""
    with tempfile.TemporaryFile(mode="w+b") as f:
        saved = os.dup(1)
        os.dup2(f.fileno(), 1)
        try:
            out = fn()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        f.seek(0)
        return out, f.read().decode()


@pytest.mark.parametrize("kind,solver,kw", [c for c in SOLVES if c[1] in KRYLOV])
def test_device_loop_branch_bit_identical_to_oracle(mock, kind, solver, kw):
    glb, lib = mock
    orc = oracle_py.load("best")
    L = 16
    U, b = synthetic(orc, L)
    op = orc.op(kind, L, L, mass=0.1, links=U)
    bb = b if op.is_complex else np.ascontiguousarray(b.real)
    args = dict(max_iter=4000, eps=1e-9)
    args.update(kw)
    xo, want = orc.solve(solver, op, bb, **args)
    before = lib.glb200_mock_krylov_calls()
    xd, got = solve(mock, solver, desc(glb, kind, L, L, mass=0.1, links=U), bb, False, **args)
    assert lib.glb200_mock_krylov_calls() > before, "the device-loop branch was not taken"
    assert got == want
    assert np.array_equal(xd, xo)
    before = lib.glb200_mock_krylov_calls()
    xs, shell = solve(mock, solver, desc(glb, kind, L, L, mass=0.1, links=U), bb, True, **args)
    assert lib.glb200_mock_krylov_calls() == before
    assert shell == want and np.array_equal(xs, xo)


@pytest.mark.parametrize("kind,solver,kw", [("STAG_U1", "BICGSTAB", {}), ("STAG_NORMAL_U1", "CR", {}),
                                            ("STAG_NORMAL_U1", "CG", {}), ("LAPLACE_REAL", "CG", dict(max_iter=7)),
                                            ("STAG_NORMAL_U1", "CG_RESTART", dict(restart_freq=32)),
                                            ("LAPLACE_REAL", "BICGSTAB", {}), ("LAPLACE_REAL", "CR", dict(max_iter=7)),
                                            ("STAG_U1", "BICGSTAB", dict(max_iter=9)),
                                            ("STAG_NORMAL_U1", "CR_RESTART", dict(restart_freq=32)),
                                            ("STAG_U1", "BICGSTAB_RESTART", dict(restart_freq=20))])
def test_device_loop_branch_prints_the_shells_lines(mock, kind, solver, kw):
    """VERB_DETAIL: one line per iteration with the reference's iteration and operator counts, replayed from the
    history the loop recorded; then the summary line"""
    glb, lib = mock
    orc = oracle_py.load("best")
    L = 16
    U, b = synthetic(orc, L)
    op = orc.op(kind, L, L, mass=0.1, links=U)
    bb = b if op.is_complex else np.ascontiguousarray(b.real)
    args = dict(max_iter=4000, eps=1e-9)
    args.update(kw)
    d = desc(glb, kind, L, L, mass=0.1, links=U)
    (xd, got), out_dev = printed(lambda: solve(mock, solver, d, bb, False, verbosity=3, **args))
    (xs, shell), out_shell = printed(lambda: solve(mock, solver, d, bb, True, verbosity=3, **args))
    assert got == shell and np.array_equal(xd, xs)
    assert out_dev == out_shell
    lines = out_dev.strip().splitlines()
    assert len(lines) >= got["iter"] + 1 and "Iter" in lines[0], lines[:3]


def test_switch_off(mock):
    """GLB200_MOCK_NO_KRYLOV=1: glb_krylov_solve_supported says no and the shells take their own loop"""
    glb, lib = mock
    orc = oracle_py.load("best")
    L = 16
    U, b = synthetic(orc, L)
    os.environ["GLB200_MOCK_NO_KRYLOV"] = "1"
    try:
        before = lib.glb200_mock_krylov_calls()
        x, got = solve(mock, "BICGSTAB", desc(glb, "STAG_U1", L, L, mass=0.1, links=U), b, False)
        assert lib.glb200_mock_krylov_calls() == before
    finally:
        del os.environ["GLB200_MOCK_NO_KRYLOV"]
    want = orc.solve("BICGSTAB", orc.op("STAG_U1", L, L, mass=0.1, links=U), b, max_iter=4000, eps=1e-9)[1]
    assert got == want
