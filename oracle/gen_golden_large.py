#!/usr/bin/env python
"""Generate tests/golden/golden_large.json from the UNMODIFIED reference (oracle/_ref) at the BASELINE sizes.

    python oracle/gen_golden_large.py [--sizes 1024 4096] [--jobs 3]

Run in the build container (needs /root/reference and `make -C oracle ref`); takes ~20 CPU-minutes at 4096^2.
Inputs are the BASELINE.md section 3 stream: std::mt19937(1337) -> gauss_gauge_u1(beta=6) -> gaussian rhs, mass 0.1
(the same generator calls the GPU tests make through oracle_py, so both sides see identical arrays).

Recorded per size L:
  * digests of the inputs and of D b, D^dag b, gamma5 D b, D^dag D b  (operators.cpp:184,262,372,444)
  * CGNE: minv_vector_cg on D^dag D, rhs D^dag b, tol 1e-10          (generic_cg.cpp:307)   -- BASELINE config 2/4 metric
  * config 3: minv_vector_cg_m, shifts {0, .01, .05, .25}, tol 1e-10  (generic_cg_m.cpp:312)
              minv_vector_gmres_restart(..., 1e-8, 20, ...) on D      (generic_gmres.cpp:778)
  * BiCGStab on D, tol 1e-10                                          (generic_bicgstab.cpp:228)
  * CR: minv_vector_cr on D^dag D, rhs D^dag b, tol 1e-10            (generic_cr.cpp:198)   [--only CR adds it to an existing file]
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_py as O  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "golden_large.json")
MASS = 0.1


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def inputs(ref, L):
    r = ref.rng(1337)
    U = r.gauss_gauge_u1(L, L, 6.0)
    b = r.gaussian(L * L)
    return U, b


def job(args):
    L, what = args
    ref = O.load("ref")
    U, b = inputs(ref, L)
    t0 = time.time()
    D = ref.op("STAG_U1", L, L, mass=MASS, links=U)
    if what == "applies":
        out = dict(links_sha=digest(U), rhs_sha=digest(b))
        out["apply_D_sha"] = digest(D.apply(b))
        out["apply_Ddag_sha"] = digest(ref.op("STAG_DAGGER_U1", L, L, mass=MASS, links=U).apply(b))
        out["apply_g5D_sha"] = digest(ref.op("STAG_GAMMA5_U1", L, L, mass=MASS, links=U).apply(b))
        out["apply_DdagD_sha"] = digest(ref.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U).apply(b))
        return what, out, time.time() - t0
    if what in ("CGNE", "CG-M", "CR"):
        bprime = ref.op("STAG_DAGGER_U1", L, L, mass=MASS, links=U).apply(b)
        DdD = ref.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
        if what in ("CGNE", "CR"):
            x, info = ref.solve("CG" if what == "CGNE" else "CR", DdD, bprime, max_iter=100000, eps=1e-10)
            info["x_sha"] = digest(x)
            info["true_rel_residual"] = float(np.linalg.norm(DdD.apply(x) - bprime) / np.linalg.norm(bprime))
        else:
            xs, info, _ = ref.solve_cg_m(DdD, bprime, [0.0, 0.01, 0.05, 0.25], resid_freq_check=10, max_iter=100000,
                                         eps=1e-10)
            info["shifts"] = [0.0, 0.01, 0.05, 0.25]
            rel = []
            for s, x in zip(info["shifts"], xs):
                rel.append(float(np.linalg.norm(DdD.apply(x) + s * x - bprime) / np.linalg.norm(bprime)))
            info["true_rel_residuals"] = rel
        return what, info, time.time() - t0
    if what == "GMRES(20)":
        x, info = ref.solve("GMRES_RESTART", D, b, max_iter=100000, eps=1e-8, restart_freq=20)
        info["true_rel_residual"] = float(np.linalg.norm(D.apply(x) - b) / np.linalg.norm(b))
        return what, info, time.time() - t0
    if what == "BiCGStab":
        x, info = ref.solve("BICGSTAB", D, b, max_iter=100000, eps=1e-10)
        info["true_rel_residual"] = float(np.linalg.norm(D.apply(x) - b) / np.linalg.norm(b))
        return what, info, time.time() - t0
    raise ValueError(what)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[1024, 4096])
    ap.add_argument("--jobs", type=int, default=3)
    ap.add_argument("--only", nargs="+", default=None, help="run only these entries (e.g. CR) and merge them into the file")
    args = ap.parse_args()
    gold = {}
    if os.path.exists(OUT):
        gold = json.load(open(OUT))
    gold["generator"] = "oracle/gen_golden_large.py"
    gold["oracle_kind"] = O.load("ref").kind
    gold["inputs"] = "std::mt19937(1337): gauss_gauge_u1(L, L, beta=6) then gaussian(L*L); mass 0.1"
    for L in args.sizes:
        todo = [(L, w) for w in ("GMRES(20)", "CG-M", "CGNE", "BiCGStab", "CR", "applies") if not args.only or w in args.only]
        entry = gold.get(str(L), {})
        with mp.get_context("fork").Pool(args.jobs) as pool:
            for what, info, dt in pool.imap_unordered(job, todo):
                print("L=%d %-10s %.0f s  %s" % (L, what, dt, {k: v for k, v in info.items() if k != "resSqmrhs"}),
                      flush=True)
                if what == "applies":
                    entry.update(info)
                else:
                    info["cpu_seconds"] = round(dt, 1)
                    entry[what] = info
        gold[str(L)] = entry
        with open(OUT, "w") as f:
            json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
