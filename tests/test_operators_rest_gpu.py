"""The rest of operator_utils/operators.h on the GPU: the symmetric shifts, the staggered operator with a two-link
Laplace term and the staggered index operator (operators.cpp:625-835), through the reference-named host callbacks.
The symmetric shifts run as nc = 1 stencils whose entries are the links scaled by 1/2 -- exact -- so they equal the
reference's functions bit for bit; the two-link operator sums its thirteen terms in the stencil kernel's order instead
of the function's, and the index operator composes five applies: both agree to rounding (gate 1e-13).
(tests/test_gpu_logic_on_mock_cpu.py runs this module's Python side against the CPU mock of the C ABI.)"""
import numpy as np
import pytest

import oracle_py
from conftest import rel_err, synthetic

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif("ref" not in oracle_py.available(),
                                 reason="these operators are only in oracle/_ref/libref_oracle.so")]


@pytest.mark.parametrize("L", [8, 32, 66])
def test_symmetric_shifts_two_link_and_index_operator(ctx, glb, L):
    orc = oracle_py.load("ref")
    U, b = synthetic(orc, L)
    m, w = 0.13, 0.35
    for kind in ("SYMMSHIFT_X", "SYMMSHIFT_Y"):
        want = orc.op(kind, L, L, mass=m, links=U).apply(b)
        got = ctx.host_apply(ctx._desc(kind, L, L, mass=m, links=U), b)
        assert rel_err(got, want) < 1e-15 and np.allclose(got, want, rtol=0, atol=0), kind     # equal up to the sign of zeros
    want = orc.op("STAG_2LINK_U1", L, L, mass=m, links=U, wilson_coeff=w).apply(b)
    got = ctx.host_apply(ctx._desc("STAG_2LINK_U1", L, L, mass=m, links=U, wilson_coeff=w), b)
    assert rel_err(got, want) < 1e-13
    # w = 0: the plain staggered operator
    got0 = ctx.host_apply(ctx._desc("STAG_2LINK_U1", L, L, mass=m, links=U, wilson_coeff=0.0), b)
    assert rel_err(got0, orc.op("STAG_U1", L, L, mass=m, links=U).apply(b)) < 1e-14
    want = orc.op("STAG_INDEX", L, L, mass=m, links=U).apply(b)
    got = ctx.host_apply(ctx._desc("STAG_INDEX", L, L, mass=m, links=U), b)
    assert rel_err(got, want) < 1e-13


@pytest.mark.parametrize("kind,solver,kw", [("STAG_2LINK_U1", "BICGSTAB", dict(wilson_coeff=0.3)),
                                            ("STAG_2LINK_U1", "GCR_RESTART", dict(wilson_coeff=0.3, restart_freq=16)),
                                            ("STAG_INDEX", "GCR_RESTART", dict(restart_freq=32))])
def test_solves_on_the_remaining_operators(ctx, glb, kind, solver, kw):
    """any minv_* drop-in takes these callbacks (tests/bicgstab_l/bicgstab_l.cpp: --operator index)"""
    orc = oracle_py.load("ref")
    L, m = 32, 0.3
    U, b = synthetic(orc, L)
    opkw = dict(wilson_coeff=kw.get("wilson_coeff", 0.0))
    rf = kw.get("restart_freq", 0)
    oop = orc.op(kind, L, L, mass=m, links=U, **opkw)
    xo, want = orc.solve(solver, oop, b, max_iter=4000, eps=1e-9, restart_freq=rf)
    x = np.zeros_like(b)
    got = ctx.host_solve(solver, ctx._desc(kind, L, L, mass=m, links=U, **opkw), x, b, max_iter=4000, eps=1e-9,
                         restart_freq=rf)
    assert got["success"] == want["success"] and got["name"] == want["name"]
    tol_it = 0.10 if solver == "BICGSTAB" else 0.02
    assert abs(got["iter"] - want["iter"]) <= max(1, round(tol_it * want["iter"]))
    if want["success"]:
        assert np.linalg.norm(oop.apply(x) - b) / np.linalg.norm(b) < 1e-9 * 1.0001
        assert rel_err(x, xo) < 1e-6
