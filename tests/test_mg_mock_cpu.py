"""Multigrid preconditioner + VPGCR shells (host/mg_dev.cpp, dev_solvers.cpp) without a GPU.

Linked against the host-memory mock of the C ABI (tests/mock), whose transfers and reductions run in
the reference's serial order, the device-side V cycle must reproduce the REFERENCE's mg_preconditioner
and minv_vector_gcr_var_precond(_restart) BIT FOR BIT on the same hierarchy: that pins the cycle's control
flow, the parameters handed to smoother and coarse solver, and the VPGCR recurrences.  Needs the
reference-compiled checker (oracle/_ref), which is where the reference's multigrid lives."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, load_pkg
from mg_common import build_reference_mg, quiet_stdout

MOCK = os.path.join(ROOT, "tests", "mock", "libglb200_inverters_mock.so")

pytestmark = pytest.mark.skipif("ref" not in oracle_py.available(),
                                reason="the reference's multigrid is only in oracle/_ref/libref_oracle.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def env():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "mock")], stdout=subprocess.DEVNULL)
    glb = load_pkg()
    lib = C.CDLL(MOCK, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    pd = C.POINTER(C.c_double)
    lib.glb_op_create_stencil2d.argtypes = [vp, vp, vp, vp, ci, ci, ci, pd, pd, pd, C.POINTER(vp)]
    lib.glb_mg_transfer_create.argtypes = [vp, ci, ci, ci, ci, ci, ci, C.POINTER(vp), C.POINTER(vp)]
    lib.glb_mg_prolong.argtypes = [vp, vp, vp]
    lib.glb_mg_restrict.argtypes = [vp, vp, vp]
    lib.glbx_mg_create.restype = vp
    lib.glbx_mg_create.argtypes = [ci, C.POINTER(vp), C.POINTER(vp)]
    lib.glbx_mg_set.argtypes = [vp, ci, ci, ci, ci, ci, ci, cd, ci, ci]
    lib.glbx_mg_vcycle.argtypes = [vp, vp, vp]
    lib.glbx_mg_vpgcr.argtypes = [vp, vp, vp, ci, cd, ci, ci, C.POINTER(glb.Result)]
    lib.glbx_mg_counts.argtypes = [vp, C.POINTER(ci)]
    lib.glbx_force_host_scalars.argtypes = [ci]
    lib.glbx_force_host_scalars(1)
    orc = oracle_py.load("ref")
    mg, U, b = build_reference_mg(orc, L=16, mass=0.01, nvec=2, block=4)
    keep = []

    def c2(z):
        a = (C.c_double * 2)(complex(z).real, complex(z).imag)
        keep.append(a)
        return a

    ops = []
    for lvl in (0, 1):
        X, Y, nc = mg.dims(lvl)
        cl, hp, sh = mg.stencil(lvl)
        keep.extend([cl, hp])
        h = vp()
        assert lib.glb_op_create_stencil2d(None, _p(cl), _p(hp), None, X, Y, nc, c2(sh[0]), c2(sh[1]), c2(sh[2]),
                                           C.byref(h)) == 0
        ops.append(h)
    nulls = [mg.null(0, v) for v in range(mg.dims(1)[2])]
    keep.append(nulls)
    ptrs = (vp * len(nulls))(*[n.ctypes.data for n in nulls])
    tr = vp()
    assert lib.glb_mg_transfer_create(None, 16, 16, 1, 4, 4, len(nulls), ptrs, C.byref(tr)) == 0
    op_p = (vp * 2)(ops[0].value, ops[1].value)
    tr_p = (vp * 1)(tr.value)
    h = lib.glbx_mg_create(1, op_p, tr_p)
    assert h
    return dict(glb=glb, lib=lib, orc=orc, mg=mg, U=U, b=b, h=h, tr=tr, keep=keep)


def test_transfers_bit_identical(env):
    mg, lib, tr = env["mg"], env["lib"], env["tr"]
    rg = np.random.default_rng(1)
    c = rg.standard_normal(mg.size(1)) + 1j * rg.standard_normal(mg.size(1))
    f = rg.standard_normal(mg.size(0)) + 1j * rg.standard_normal(mg.size(0))
    got_f, got_c = np.empty_like(f), np.empty_like(c)
    assert lib.glb_mg_prolong(tr, _p(got_f), _p(c)) == 0
    assert lib.glb_mg_restrict(tr, _p(got_c), _p(f)) == 0
    assert np.array_equal(got_f, mg.prolong(0, c))
    assert np.array_equal(got_c, mg.restrict(0, f))


@pytest.mark.parametrize("cfg", [dict(), dict(smooth="BICGSTAB", n_pre=3, n_post=2, inner="CG", rel_res=1e-3),
                                 dict(n_pre=0, n_post=4, inner="BICGSTAB"), dict(n_pre=2, n_post=0, inner="CR", n_restart=8),
                                 # (no SOR case: the reference's cycle leaves minv_inverter_params::sor_omega
                                 #  uninitialised, mg_complex.cpp:568-575; ours passes 1.0)
                                 dict(smooth="MINRES", n_pre=4, n_post=4)])
def test_vcycle_bit_identical(env, cfg):
    """one mg_preconditioner application: same bits as the reference for several smoother / coarse-solver settings"""
    mg, lib, h, b = env["mg"], env["lib"], env["h"], env["b"]
    glbmod = env["glb"]
    full = dict(smooth="GCR", n_pre=6, n_post=6, inner="GCR", n_max=1024, n_restart=64, rel_res=1e-2)
    full.update(cfg)
    mg.set_precond(**full)
    lib.glbx_mg_set(h, glbmod.Multigrid.SMOOTH[full["smooth"]], full["n_pre"], full["n_post"],
                    glbmod.Multigrid.INNER[full["inner"]], full["n_max"], full["n_restart"], full["rel_res"], 0, 1)
    with quiet_stdout():
        want = mg.vcycle(b)
    got = np.zeros_like(b)
    assert lib.glbx_mg_vcycle(h, _p(got), _p(b)) == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("restart", [64, 4, 0])
def test_vpgcr_bit_identical(env, restart):
    """the outer solve of config 5: VPGCR(64) + V cycle, tol 5e-7 -- iteration / ops counts, residual, success
    flag and the solution itself equal the reference's"""
    mg, lib, h, b = env["mg"], env["lib"], env["h"], env["b"]
    glbmod = env["glb"]
    mg.set_precond()
    lib.glbx_mg_set(h, glbmod.Multigrid.SMOOTH["GCR"], 6, 6, glbmod.Multigrid.INNER["GCR"], 1024, 64, 1e-2, 0, 1)
    with quiet_stdout():
        xo, want = mg.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=restart)
    x = np.zeros_like(b)
    res = glbmod.Result()
    assert lib.glbx_mg_vpgcr(h, _p(x), _p(b), 1000, 5e-7, restart, 0, C.byref(res)) == 0
    got = res.as_dict()
    assert (got["iter"], got["ops_count"], got["success"], got["resSq"]) == (
        want["iter"], want["ops_count"], want["success"], want["resSq"])
    assert np.array_equal(x, xo)
    assert want["success"] and want["iter"] < 40      # the preconditioner works (plain GCR needs hundreds)
