#!/usr/bin/env python
"""Generate tests/golden/*.json(.npy) from the UNMODIFIED reference (oracle/_ref).

Run in the build container (where /root/reference exists and `make -C oracle ref` has been run):

    python oracle/gen_golden.py

The fixtures pin BOTH CPU checkers (tests/test_oracle_cpu.py) and travel to the GPU box, where
/root/reference does not exist.  Inputs follow BASELINE.md / SURVEY.md section 8d:
std::mt19937(1337), gauss_gauge_u1(beta=6), gaussian rhs, mass 0.1.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_py as O  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
REFROOT = "/root/reference"


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = O.load("ref")
    gold = {"generator": "oracle/gen_golden.py", "oracle_kind": ref.kind}

    # ---- (1) unit_test.cpp:72-281 : real 128^2 Laplace, m^2 = 0.01, point source, tol 1e-6, max_iter 4000
    N = 128
    # unit_test.cpp:617-633 initialize_test: rhs AND the initial guess lhs are a delta at
    # half + half*half*2 with half = sqrt(size)/2  (= N/2 + (N/2)*N)
    half = N // 2
    src_index = half + half * half * 2
    b = np.zeros(N * N)
    b[src_index] = 1.0
    x0 = b.copy()
    op = ref.op("LAPLACE_REAL", N, N, mass=0.01)
    ut = {}
    for name, solver, kw in [
        ("CG", "CG", {}), ("CG(8)", "CG_RESTART", dict(restart_freq=8)), ("CR", "CR", {}),
        ("CR(8)", "CR_RESTART", dict(restart_freq=8)), ("GCR", "GCR", {}),
        ("GCR(8)", "GCR_RESTART", dict(restart_freq=8)), ("BiCGStab", "BICGSTAB", {}),
        ("BiCGStab(8)", "BICGSTAB_RESTART", dict(restart_freq=8)),
        ("BiCGStab-1", "BICGSTAB_L", dict(l=1, max_iter=10000)),
        ("BiCGStab-2", "BICGSTAB_L", dict(l=2, max_iter=10000)),
        ("BiCGStab-8", "BICGSTAB_L", dict(l=8, max_iter=10000)),
        ("BiCGStab-8(64)", "BICGSTAB_L_RESTART", dict(l=8, restart_freq=64, max_iter=10000)),
        ("GMRES", "GMRES", {}), ("GMRES(8)", "GMRES_RESTART", dict(restart_freq=8)),
    ]:
        kw = dict(kw)
        mi = kw.pop("max_iter", 4000)
        x, info = ref.solve(solver, op, b, x0=x0, max_iter=mi, eps=1e-6, **kw)
        info["x_sha"] = digest(x)
        info["call"] = dict(solver=solver, max_iter=mi, **kw)
        ut[name] = info
    gold["unit_test_128"] = dict(N=N, mass_sq=0.01, src_index=int(src_index), tol=1e-6, max_iter=4000, results=ut)

    # ---- (2) config 1: square_laplace.cpp with N=64, minv_vector_cg tol 1e-10 (square_laplace.cpp:52,62,78)
    N = 64
    b = np.zeros(N * N)
    b[N // 2 + (N // 2) * N] = 1.0
    x0 = np.zeros(N * N)
    x0[N // 2 + (N // 2) * N + 1] = 1.0
    op = ref.op("LAPLACE_REAL", N, N, mass=0.01)
    x, info = ref.solve("CG", op, b, x0=x0, max_iter=4000, eps=1e-10)
    info["x_sha"] = digest(x)
    np.save(os.path.join(OUT, "config1_solution.npy"), x)
    gold["config1_laplace64_cg"] = dict(N=N, mass_sq=0.01, tol=1e-10, max_iter=4000, result=info)

    # ---- (3) tests/staggered_stencil/staggered_stencil.cpp:206-251 on cfg l64t64b60_heatbath
    cfg = os.path.join(REFROOT, "multigrid/aa_mg/cfg/l64t64b60_heatbath.dat")
    L = 64
    U = ref.read_gauge(L, L, cfg)
    phases = np.loadtxt(cfg)
    np.save(os.path.join(OUT, "l64t64b60_heatbath_phases.npy"), phases)  # file order: x outer, y, mu inner
    src = np.zeros(L * L, dtype=np.complex128)
    src[L + 1] = 1.0
    fn = ref.op("STAG_U1", L, L, mass=0.01, links=U).apply(src)
    st = ref.op("STENCIL_FROM_STAG", L, L, mass=0.01, links=U).apply(src)
    nz = np.flatnonzero(fn)
    gold["staggered_stencil_64"] = dict(
        L=L, mass=0.01, src_index=L + 1, plaquette=[ref.plaquette(U, L, L).real, ref.plaquette(U, L, L).imag],
        function_vs_stencil_diffnorm2sq=float(ref.diffnorm2sq(fn, st)),
        nonzero_index=[int(i) for i in nz], nonzero_re=[float(fn[i].real) for i in nz],
        nonzero_im=[float(fn[i].imag) for i in nz], links_sha=digest(U), out_sha=digest(fn))

    # ---- (4) synthetic inputs of BASELINE.md section 2 : solver behaviour at 64^2 and 256^2
    syn = {}
    for L in (64, 256):
        r = ref.rng(1337)
        U = r.gauss_gauge_u1(L, L, 6.0)
        b = r.gaussian(L * L)
        entry = dict(links_sha=digest(U), rhs_sha=digest(b), plaquette=ref.plaquette(U, L, L).real)
        D = ref.op("STAG_U1", L, L, mass=0.1, links=U)
        DdD = ref.op("STAG_NORMAL_U1", L, L, mass=0.1, links=U)
        Dd = ref.op("STAG_DAGGER_U1", L, L, mass=0.1, links=U)
        entry["apply_D_sha"] = digest(D.apply(b))
        bprime = Dd.apply(b)  # CGNE right-hand side D^dag b
        x, entry["CGNE"] = ref.solve("CG", DdD, bprime, max_iter=100000, eps=1e-10)
        x, entry["CG_on_normal_rhs_b"] = ref.solve("CG", DdD, b, max_iter=100000, eps=1e-10)
        x, entry["CR_on_normal_rhs_b"] = ref.solve("CR", DdD, b, max_iter=100000, eps=1e-10)
        x, entry["BiCGStab"] = ref.solve("BICGSTAB", D, b, max_iter=100000, eps=1e-10)
        x, entry["BiCGStab-4"] = ref.solve("BICGSTAB_L", D, b, max_iter=100000, eps=1e-10, l=4)
        x, entry["GMRES(20)"] = ref.solve("GMRES_RESTART", D, b, max_iter=100000, eps=1e-8, restart_freq=20)
        x, entry["GCR(20)"] = ref.solve("GCR_RESTART", D, b, max_iter=100000, eps=1e-8, restart_freq=20)
        xs, info, _ = ref.solve_cg_m(DdD, bprime, [0.0, 0.01, 0.05, 0.25], resid_freq_check=10, max_iter=100000,
                                     eps=1e-10)
        entry["CG-M"] = info  # rhs = D^dag b, as in BASELINE.md section 2 (171 it / 175 ops)
        syn[str(L)] = entry
    gold["synthetic_beta6_m0.1"] = syn

    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1, sort_keys=True)[:6000])


if __name__ == "__main__":
    main()
