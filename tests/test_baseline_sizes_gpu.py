"""GPU parity at the BASELINE lattice sizes (1024^2 and the headline 4096^2), against

  * tests/golden/golden_large.json -- produced by the UNMODIFIED reference here (oracle/gen_golden_large.py, ~25 CPU
    minutes: digests of the applies, iteration counts and residuals of CGNE, CG-M, GMRES(20), BiCGStab, CR), and
  * the CPU oracle itself for the applies (np.array_equal; one 4096^2 apply costs the CPU 0.25 s).

Inputs: std::mt19937(1337) -> gauss_gauge_u1(beta = 6) -> gaussian rhs, drawn by the product's
own host helpers (glbx_synthetic_inputs) and pinned to the reference's stream by digest.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu
MASS = 0.1


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "golden_large.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module", params=[1024, 4096])
def case(request, ctx, glb, gold):
    L = request.param
    U, b = ctx.synthetic_inputs(L, L, 1337, 6.0)
    g = gold[str(L)]
    assert digest(U) == g["links_sha"] and digest(b) == g["rhs_sha"]   # the reference's own input stream
    return L, U, b, g


def close_iters(got, want, tol=0.02):
    return abs(got - want) <= max(1, int(round(tol * want)))


def test_applies_bit_identical(ctx, glb, orc, case):
    L, U, b, g = case
    V = L * L
    x, out = ctx.vector(V).upload(b), ctx.vector(V)
    for flags, key, kind in ((0, "apply_D_sha", "STAG_U1"), (glb.STAG_DAGGER, "apply_Ddag_sha", "STAG_DAGGER_U1"),
                             (glb.STAG_GAMMA5, "apply_g5D_sha", "STAG_GAMMA5_U1"),
                             (glb.STAG_NORMAL, "apply_DdagD_sha", "STAG_NORMAL_U1")):
        op = ctx.staggered(U, L, L, MASS, flags)
        op.apply(out, x)
        got = out.download()
        assert digest(got) == g[key], (L, key)                      # the reference's output, bit for bit
        if flags in (0, glb.STAG_NORMAL):                            # and against the oracle run here, element-wise
            assert np.array_equal(got, orc.op(kind, L, L, mass=MASS, links=U).apply(b)), (L, kind)
        op.destroy()


def test_cg_first_iterations_at_size(ctx, glb, orc, case):
    """the single-kernel CG iteration and the two-kernel loop (one-pass D^dag D with the fused direction update)
    against the reference's iterate after 1 and 2 iterations: every site of r, p, q, x has been through the kernels"""
    L, U, b, g = case
    V = L * L
    oN = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    rhs = orc.op("STAG_DAGGER_U1", L, L, mass=MASS, links=U).apply(b)
    N = ctx.staggered(U, L, L, MASS, glb.STAG_NORMAL)
    bd = ctx.vector(V).upload(rhs)
    for m in (1, 2):
        want, winfo = orc.solve("CG", oN, rhs, max_iter=m, eps=1e-30)
        for single in (True, False):
            ctx.cg_step_mode(single)
            try:
                x = ctx.vector(V).zero()
                rep = ctx.cg_device(N, x, bd, max_iter=m, eps=1e-30)
            finally:
                ctx.cg_step_mode(True)
            assert rep["iterations"] == m == winfo["iter"]
            assert rel_err(x.download(), want) < 1e-13, (L, m, single)
    N.destroy()


def test_cgne_matches_reference_run(ctx, glb, orc, case):
    L, U, b, g = case
    V = L * L
    want = g["CGNE"]
    N = ctx.staggered(U, L, L, MASS, glb.STAG_NORMAL)
    Dd = ctx.staggered(U, L, L, MASS, glb.STAG_DAGGER)
    bd, bp, x = ctx.vector(V).upload(b), ctx.vector(V), ctx.vector(V).zero()
    Dd.apply(bp, bd)
    info = ctx.solve("CG", N, x, bp, max_iter=100000, eps=1e-10)
    assert info["success"] and close_iters(info["iter"], want["iter"]) and info["ops_count"] == info["iter"] + 2
    rhs = bp.download()
    rr = float(np.linalg.norm(orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U).apply(x.download()) - rhs) /
               np.linalg.norm(rhs))
    assert rr < 1.05e-10 and abs(rr - want["true_rel_residual"]) < 2e-11, (rr, want["true_rel_residual"])
    assert 0.0 < ctx.cg_last_pred_err() < 1e-10
    for o in (N, Dd):
        o.destroy()


def test_config3_cg_m_and_gmres(ctx, glb, orc, case):
    """BASELINE config 3: minv_vector_cg_m with shifts {0, .01, .05, .25} (generic_cg_m.cpp:312) and
    minv_vector_gmres_restart(..., 1e-8, 20, ...) (generic_gmres.cpp:778) at the lattice size of the reference run"""
    L, U, b, g = case
    V = L * L
    D = ctx.staggered(U, L, L, MASS, 0)
    N = ctx.staggered(U, L, L, MASS, glb.STAG_NORMAL)
    Dd = ctx.staggered(U, L, L, MASS, glb.STAG_DAGGER)
    bd, bp = ctx.vector(V).upload(b), ctx.vector(V)
    Dd.apply(bp, bd)
    shifts = [0.0, 0.01, 0.05, 0.25]
    xs = [ctx.vector(V).zero() for _ in shifts]
    info, _ = ctx.solve_cg_m(N, xs, bp, shifts, resid_freq_check=10, max_iter=100000, eps=1e-10)
    want = g["CG-M"]
    assert info["success"] and close_iters(info["iter"], want["iter"]), (info["iter"], want["iter"])
    rhs = bp.download()
    oN = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    for s, xv, wr in zip(shifts, xs, want["true_rel_residuals"]):
        xh = xv.download()
        rr = float(np.linalg.norm(oN.apply(xh) + s * xh - rhs) / np.linalg.norm(rhs))
        assert rr < 1.05e-10 and rr < 3 * wr + 1e-12, (s, rr, wr)
    del xs
    x = ctx.vector(V).zero()
    info = ctx.solve("GMRES_RESTART", D, x, bd, max_iter=100000, eps=1e-8, restart_freq=20)
    want = g["GMRES(20)"]
    assert close_iters(info["iter"], want["iter"]) and info["success"] == want["success"], (info, want)
    rr = float(np.linalg.norm(orc.op("STAG_U1", L, L, mass=MASS, links=U).apply(x.download()) - b) / np.linalg.norm(b))
    assert rr < 1.05e-8
    x.zero()
    info = ctx.solve("BICGSTAB", D, x, bd, max_iter=100000, eps=1e-10)
    want = g["BiCGStab"]
    # BiCGStab is chaotic on this operator (the reference's own count moves by +-8 % under 1e-15 perturbations,
    # tests/test_solvers_gpu.py): the bar here is the envelope, the stable-mass case keeps the +-2 % bar
    assert info["success"] and close_iters(info["iter"], want["iter"], tol=0.10), (info["iter"], want["iter"])
    rr = float(np.linalg.norm(orc.op("STAG_U1", L, L, mass=MASS, links=U).apply(x.download()) - b) / np.linalg.norm(b))
    assert rr < 1.05e-10, rr
    # minv_vector_cr on D^dag D (generic_cr.cpp:198) through the device-resident loop of csrc/krylov.cu
    x.zero()
    info = ctx.solve("CR", N, x, bp, max_iter=100000, eps=1e-10)
    want = g["CR"]
    assert info["success"] and close_iters(info["iter"], want["iter"]), (info["iter"], want["iter"])
    rr = float(np.linalg.norm(oN.apply(x.download()) - rhs) / np.linalg.norm(rhs))
    assert rr < 1.05e-10 and abs(rr - want["true_rel_residual"]) < 2e-11, (rr, want["true_rel_residual"])
    for o in (D, N, Dd):
        o.destroy()
