"""Shared set-up of the multigrid parity tests (CPU mock and GPU): a small staggered problem whose
hierarchy -- block-orthonormalised null vectors, fine and coarse stencils -- is built by the REFERENCE's
own code (oracle/ref_mg_shim.cpp) and then handed, as plain arrays, both to the reference's
mg_preconditioner / VPGCR and to ours."""
import os

import numpy as np

import oracle_py


def build_reference_mg(orc, L=16, mass=0.01, nvec=2, block=4, seed=1337, relax=30):
    """returns (RefMg, U, b).  Null vectors: errors of partially solved D^dag D e = D^dag D r (rich in
    low modes), split into even/odd parts like the reference's default BLOCK_EO strategy
    (input_params.cpp:736-741), 2*nvec vectors in all."""
    rng = orc.rng(seed)
    U = rng.gauss_gauge_u1(L, L, 6.0)
    b = rng.gaussian(L * L)
    N = orc.op("STAG_NORMAL_U1", L, L, mass=mass, links=U)
    raw = []
    for _ in range(nvec):
        r = rng.gaussian(L * L)
        e, _info = orc.solve("CG", N, N.apply(r), max_iter=relax, eps=1e-12)
        raw.append(r - e)
    idx = np.arange(L * L)
    even = ((idx % L + idx // L) % 2) == 0
    vecs = [np.where(even, v, 0) for v in raw] + [np.where(~even, v, 0) for v in raw]
    mg = oracle_py.RefMg(orc, L, L, U, mass, [block], [2 * nvec], [vecs])
    return mg, U, b


class quiet_stdout:
    """the reference prints progress lines from mg_preconditioner with printf/cout: silence fd 1"""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)
