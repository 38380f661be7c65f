"""tools/mg_setup.py (numpy set-up used by the multigrid measurements) against the reference's own set-up."""
import os
import sys

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, rel_err

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif("ref" not in oracle_py.available(),
                                reason="the reference's multigrid is only in oracle/_ref/libref_oracle.so")


@pytest.mark.parametrize("L,block,nraw", [(16, 4, 2), (24, 4, 4), (16, 2, 1)])
def test_numpy_setup_matches_reference(L, block, nraw):
    import mg_setup
    orc = oracle_py.load("ref")
    rng = orc.rng(7)
    U = rng.gauss_gauge_u1(L, L, 6.0)
    mass = 0.03
    raw = [rng.gaussian(L * L) for _ in range(nraw)]
    vecs = mg_setup.split_even_odd(raw, L, L)
    mg = oracle_py.RefMg(orc, L, L, U, mass, [block], [len(vecs)], [vecs])
    # fine stencil
    cl0, hp0, sh0 = mg.stencil(0)
    cl, hp, sh = mg_setup.staggered_stencil(U, L, L, mass)
    assert np.array_equal(hp, hp0) and np.array_equal(cl, cl0) and sh == sh0[0]
    # block-orthonormal null vectors
    ortho = mg_setup.block_orthonormalize(vecs, L, L, block, block)
    for v in range(len(vecs)):
        assert rel_err(ortho[v], mg.null(0, v)) < 1e-12
    # Galerkin coarse stencil
    clc, hpc = mg_setup.coarse_stencil([mg.null(0, v) for v in range(len(vecs))], hp, sh, L, L, block, block)
    clr, hpr, shr = mg.stencil(1)
    assert np.all(shr == 0)
    assert rel_err(clc, clr) < 1e-13 and rel_err(hpc, hpr) < 1e-13
