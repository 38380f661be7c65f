"""The reference's HOST-pointer multigrid interface (multigrid/aa_mg/mg_complex.h, null_gen.h) as a drop-in.

oracle/ref_mg_shim.cpp is a client of that interface: it fills a mg_operator_struct_complex the way the reference's driver
does and calls block_orthonormalize, generate_coarse_from_fine_stencil, level_down / level_up, prolong, restrict,
fine_ / coarse_square_staggered, mg_preconditioner, minv_vector_gcr_var_precond(_restart) and
null_generate_random_smooth.  oracle/_ref/libref_oracle.so is that file compiled with the REFERENCE's sources; here the
very same file is compiled against generic-linalg_b200/host and linked with this repository's shells (on the host-memory
mock of the C ABI), and both libraries are driven through the same Python wrapper on the same inputs.  What runs the
reference's loops in the reference's order has to agree bit for bit (null vectors, transfers, level-0 applies, apply
counts); what is computed differently -- the Galerkin product is summed directly instead of probed -- to rounding, and
with it the cycle and the preconditioned solve."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from conftest import ROOT, rel_err
from mg_common import quiet_stdout

MOCK_DIR = os.path.join(ROOT, "tests", "mock")

pytestmark = pytest.mark.skipif("ref" not in oracle_py.available(),
                                reason="the reference's multigrid is only in oracle/_ref/libref_oracle.so")


class _Lib:
    kind = "reference"      # what oracle_py.RefMg asks of the library it is handed


@pytest.fixture(scope="module")
def ours(tmp_path_factory):
    subprocess.check_call(["make", "-C", MOCK_DIR], stdout=subprocess.DEVNULL)
    so = str(tmp_path_factory.mktemp("mgshim") / "libours_mg_shim.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "generic-linalg_b200", "host"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "oracle", "ref_mg_shim.cpp"), "-o", so, "-L" + MOCK_DIR,
                           "-l:libglb200_inverters_mock.so", "-Wl,-rpath," + MOCK_DIR])
    mock = C.CDLL(os.path.join(MOCK_DIR, "libglb200_inverters_mock.so"), mode=C.RTLD_GLOBAL)
    mock.glbx_force_host_scalars.argtypes = [C.c_int]
    mock.glbx_force_host_scalars(1)
    lib = _Lib()
    lib.lib = C.CDLL(so, mode=C.RTLD_LOCAL)
    return lib


def _raw_null_vectors(orc, L, mass, nvec, seed=1337, relax=30):
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
    return U, b, [np.where(even, v, 0) for v in raw] + [np.where(~even, v, 0) for v in raw]


@pytest.mark.parametrize("ignore_shifts", [False, True])
def test_hierarchy_from_given_null_vectors(ours, ignore_shifts):
    orc = oracle_py.load("ref")
    L, mass = 16, 0.01
    U, b, vecs = _raw_null_vectors(orc, L, mass, 2)
    with quiet_stdout():
        mo = oracle_py.RefMg(ours, L, L, U, mass, [4], [4], [vecs], ignore_shifts=ignore_shifts)
        mr = oracle_py.RefMg(orc, L, L, U, mass, [4], [4], [vecs], ignore_shifts=ignore_shifts)
    for v in range(4):                                               # block_orthonormalize
        assert np.array_equal(mo.null(0, v), mr.null(0, v))
    (c0, h0, s0), (cr0, hr0, sr0) = mo.stencil(0), mr.stencil(0)
    assert np.array_equal(c0, cr0) and np.array_equal(h0, hr0) and np.array_equal(s0, sr0)
    (c1, h1, s1), (cr1, hr1, sr1) = mo.stencil(1), mr.stencil(1)     # generate_coarse_from_fine_stencil
    assert rel_err(c1, cr1) < 1e-13 and rel_err(h1, hr1) < 1e-13 and np.array_equal(s1, sr1)
    rg = np.random.default_rng(1)
    c = rg.standard_normal(mr.size(1)) + 1j * rg.standard_normal(mr.size(1))
    f = rg.standard_normal(mr.size(0)) + 1j * rg.standard_normal(mr.size(0))
    assert np.array_equal(mo.prolong(0, c), mr.prolong(0, c))
    assert np.array_equal(mo.restrict(0, f), mr.restrict(0, f))
    assert np.array_equal(mo.apply_level(0, f), mr.apply_level(0, f))   # fine_square_staggered
    assert rel_err(mo.apply_level(1, c), mr.apply_level(1, c)) < 1e-13  # coarse_square_staggered
    for cfg in (dict(), dict(smooth="BICGSTAB", n_pre=3, n_post=2, inner="CG", rel_res=1e-3),
                dict(smooth="MINRES", n_pre=4, n_post=4, inner="MINRES", rel_res=1e-1)):
        mo.set_precond(**cfg)
        mr.set_precond(**cfg)
        with quiet_stdout():
            vo, vr = mo.vcycle(b), mr.vcycle(b)                       # mg_preconditioner
        assert rel_err(vo, vr) < 1e-11
    mo.set_precond()
    mr.set_precond()
    for restart in (64, 0):
        with quiet_stdout():
            xo, io = mo.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=restart)   # the drop-in solve with both callbacks
            xr, ir = mr.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=restart)
        assert (io["iter"], io["ops_count"], io["success"]) == (ir["iter"], ir["ops_count"], ir["success"])
        assert rel_err(xo, xr) < 1e-10 and abs(io["resSq"] - ir["resSq"]) <= 1e-6 * ir["resSq"]


@pytest.mark.parametrize("normal_smooth,normal_mg", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("levels", [1, 2])
def test_normal_equation_variants_of_the_cycle(ours, normal_smooth, normal_mg, levels):
    """mg_precond_struct_complex::normal_eqn_smooth (CGNR smoother: D^dag D z = D^dag r) and normal_eqn_mg (the cycle on
    D^dag D, fine and coarse) with dagger stencils on every level, wired as the reference's driver does
    (aa_mg_square_staggered_u1.cpp:550-577, :656-681, :990-1116; mg_complex.cpp:537-580, :779-803): the cycle, the
    operator counts it books, and -- for the CGNR smoother -- the preconditioned solve"""
    orc = oracle_py.load("ref")
    L, mass = 16, 0.05
    U, b, vecs = _raw_null_vectors(orc, L, mass, 2)
    blocks, nvecs, nulls = [4], [4], [vecs]
    if levels == 2:
        with quiet_stdout():
            one = oracle_py.RefMg(orc, L, L, U, mass, [4], [4], [vecs])
        _, _, more = _raw_null_vectors(orc, L, mass, 2, seed=99)      # level-1 null vectors: restricted smooth vectors
        lvl1 = [one.restrict(0, more[0] + more[2]), one.restrict(0, more[1] + more[3])]
        idx = np.arange(4 * 4 * 4)
        top = (idx % 4) < 2
        blocks, nvecs, nulls = [4, 2], [4, 4], [vecs, [np.where(top, v, 0) for v in lvl1] + [np.where(~top, v, 0) for v in lvl1]]
    with quiet_stdout():
        mo = oracle_py.RefMg(ours, L, L, U, mass, blocks, nvecs, nulls)
        mr = oracle_py.RefMg(orc, L, L, U, mass, blocks, nvecs, nulls)
        for m in (mo, mr):
            m.set_normal(normal_smooth, normal_mg)
    rng = np.random.default_rng(1)
    for lvl in range(levels + 1):                                   # fine_ / coarse_square_staggered_dagger / _normal
        X, Y, nc = mr.dims(lvl)
        f = rng.standard_normal(X * Y * nc) + 1j * rng.standard_normal(X * Y * nc)
        for which in ("dagger", "normal"):
            vo, vr = mo.apply_level_variant(lvl, f, which), mr.apply_level_variant(lvl, f, which)
            assert np.array_equal(vo, vr) if lvl == 0 else rel_err(vo, vr) < 1e-13
    # With normal_eqn_mg the cycle solves on D^dag D (condition number ~ 1/m^2 squared) and amplifies rounding: the
    # REFERENCE run twice with null vectors that differ by 1e-15 relative gives cycles that differ by 3e-8 (GCR) to 1e-7
    # (CG inside); the level operators above agree to 1e-16, the operator counts exactly.
    tol = 5e-6 if normal_mg else 1e-10
    # CG as the coarse solver only where the coarse operator is D^dag D
    for cfg in (dict(smooth="CG", n_pre=3, n_post=3, inner="CG" if normal_mg else "GCR", rel_res=1e-3),
                dict(smooth="GCR", n_pre=2, n_post=4, inner="GCR", rel_res=1e-2),
                dict(smooth="CG", n_pre=2, n_post=2, inner="GCR", rel_res=1e-2, recursive=True),
                # the other recursions of mg_complex.cpp:672-695: flexible CG (on D^dag D only), preconditioned BiCGStab
                dict(smooth="CG", n_pre=2, n_post=2, inner="CG", rel_res=1e-2, recursive=True, only_normal_mg=True),
                dict(smooth="CG", n_pre=2, n_post=2, inner="BICGSTAB", rel_res=1e-3, recursive=True, erratic=True)):
        cfg = dict(cfg)
        if cfg.pop("only_normal_mg", False) and not normal_mg:
            continue
        if cfg.get("recursive") and levels == 1:
            continue
        if cfg.get("erratic") and normal_mg:        # BiCGStab on D^dag D runs into n_max on both sides
            continue
        erratic = cfg.pop("erratic", False)
        mo.set_precond(**cfg)
        mr.set_precond(**cfg)
        before_o, before_r = mo.counts(), mr.counts()
        with quiet_stdout():
            vo, vr = mo.vcycle(b), mr.vcycle(b)
        if erratic:
            # BiCGStab preconditioned by a cycle that itself stops on a tolerance: the REFERENCE run twice with null vectors
            # 1e-15 apart differs by 1e-3 relative here (5e-2 at rel_res = 0.1) and books different operator counts.  What
            # can be held: both are equally good approximate inverses, and they agree to the accuracy of the inner solves.
            A = (lambda m, v: m.apply_level_variant(0, v, "normal")) if normal_mg else (lambda m, v: m.apply_level(0, v))
            ro, rr = rel_err(A(mo, vo), b), rel_err(A(mr, vr), b)
            assert ro < 1.0 and ro < 3 * rr + 1e-3 and rel_err(vo, vr) < 5e-2, (cfg, ro, rr)
            continue
        assert rel_err(vo, vr) < tol, cfg
        delta = lambda a, z: {k: [y - x for x, y in zip(a[k], z[k])] for k in a}
        assert delta(before_o, mo.counts()) == delta(before_r, mr.counts()), cfg
    # the preconditioned solve; with normal_eqn_mg its operator is fine_square_staggered_normal (driver :1696-1705),
    # which the drop-in runs as a composition of the level operator and its dagger on the device
    mo.set_precond(smooth="CG", n_pre=3, n_post=3, inner="GCR", rel_res=1e-2)
    mr.set_precond(smooth="CG", n_pre=3, n_post=3, inner="GCR", rel_res=1e-2)
    with quiet_stdout():
        xo, io = mo.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
        xr, ir = mr.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
    assert io["success"] and ir["success"]
    if normal_mg:
        assert abs(io["iter"] - ir["iter"]) <= 1 and rel_err(xo, xr) < 1e-4
    else:
        assert (io["iter"], io["ops_count"]) == (ir["iter"], ir["ops_count"]) and rel_err(xo, xr) < 1e-9


def test_dagger_of_a_level_without_dagger_stencils(ours):
    """coarse_square_staggered_dagger with no dagger stencil falls back to prolong -> dagger one level up -> restrict
    (mg_complex.cpp:101-111), recursively: the exact adjoint of the Galerkin operator for ANY null vectors -- here
    unpartitioned ones, for which sigma_3 D sigma_3 would not be the adjoint"""
    orc = oracle_py.load("ref")
    L, mass = 16, 0.05
    U, b, vecs = _raw_null_vectors(orc, L, mass, 2)
    whole = [vecs[0] + vecs[2], vecs[1] + vecs[3]]
    rng = np.random.default_rng(5)
    lvl1 = [rng.standard_normal(4 * 4 * 2) + 1j * rng.standard_normal(4 * 4 * 2) for _ in range(2)]
    with quiet_stdout():
        mo = oracle_py.RefMg(ours, L, L, U, mass, [4, 2], [2, 2], [whole, lvl1])
        mr = oracle_py.RefMg(orc, L, L, U, mass, [4, 2], [2, 2], [whole, lvl1])
        for m in (mo, mr):
            m.set_normal(False, False, dagger_stencils=False)
    for lvl in (0, 1, 2):
        X, Y, nc = mr.dims(lvl)
        f = rng.standard_normal(X * Y * nc) + 1j * rng.standard_normal(X * Y * nc)
        g = rng.standard_normal(X * Y * nc) + 1j * rng.standard_normal(X * Y * nc)
        for which in ("dagger", "normal"):
            vo, vr = mo.apply_level_variant(lvl, f, which), mr.apply_level_variant(lvl, f, which)
            assert np.array_equal(vo, vr) if lvl == 0 else rel_err(vo, vr) < 1e-13
        lhs, rhs = np.vdot(g, mo.apply_level(lvl, f)), np.vdot(mo.apply_level_variant(lvl, g, "dagger"), f)
        assert abs(lhs - rhs) < 1e-12 * abs(lhs)                      # <g, D f> = <D^dag g, f>
    # the CGNR-smoothed cycle and solve on this hierarchy: the device cycle gets D^dag of the lower levels as the
    # adjoint of their stencils where the reference projects
    for m in (mo, mr):
        m.set_normal(True, False, dagger_stencils=False)
        m.set_precond(smooth="CG", n_pre=3, n_post=3, inner="GCR", rel_res=1e-2)
    before_o, before_r = mo.counts(), mr.counts()
    with quiet_stdout():
        vo, vr = mo.vcycle(b), mr.vcycle(b)
    assert rel_err(vo, vr) < 1e-10
    assert {k: [y - x for x, y in zip(before_o[k], mo.counts()[k])] for k in before_o} == \
           {k: [y - x for x, y in zip(before_r[k], mr.counts()[k])] for k in before_r}
    with quiet_stdout():
        xo, io = mo.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
        xr, ir = mr.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
    assert io["success"] and (io["iter"], io["ops_count"]) == (ir["iter"], ir["ops_count"]) and rel_err(xo, xr) < 1e-9
    for m in (mo, mr):                                                 # and the cycle on D^dag D of every level
        m.set_normal(False, True, dagger_stencils=False)
        m.set_precond(smooth="CG", n_pre=3, n_post=3, inner="CG", rel_res=1e-2)
    with quiet_stdout():
        vo, vr = mo.vcycle(b), mr.vcycle(b)
    assert rel_err(vo, vr) < 5e-6


@pytest.mark.parametrize("kw", [dict(seed=11), dict(seed=5, do_ortho_eo=True), dict(seed=8, null_prec=2, null_gen="CG", tol=1e-3),
                                dict(seed=12, bstrat=2, max_iter=60), dict(seed=14, bstrat=3, max_iter=60), dict(do_free=True)])
def test_null_vector_generation_through_the_host_interface(ours, kw):
    """null_generate_random_smooth / null_generate_free / the partitions on the host struct"""
    orc = oracle_py.load("ref")
    L = 16
    U = orc.rng(7).gauss_gauge_u1(L, L, 6.0)
    nv = 8 if kw.get("bstrat") == 2 else 4
    if kw.get("do_free"):
        nv = 2
    with quiet_stdout():
        so = oracle_py.RefMg.setup(ours, L, L, U, 0.05, [4], [nv], **kw)
        sr = oracle_py.RefMg.setup(orc, L, L, U, 0.05, [4], [nv], **kw)
    for v in range(nv):
        assert np.array_equal(so.null(0, v), sr.null(0, v))
    assert so.null_counts() == sr.null_counts()
    (c1, h1, s1), (cr1, hr1, sr1) = so.stencil(1), sr.stencil(1)
    assert rel_err(c1, cr1) < 1e-13 and rel_err(h1, hr1) < 1e-13 and np.array_equal(s1, sr1)


def test_three_levels_through_the_host_interface(ours):
    """level_down / level_up, the colour-index partition and the level-1 -> level-2 Galerkin product (few smoothing
    iterations below the top level: see tests/test_mg_setup_mock_cpu.py for why)"""
    orc = oracle_py.load("ref")
    L = 32
    rng = orc.rng(7)
    U = rng.gauss_gauge_u1(L, L, 6.0)
    b = rng.gaussian(L * L)
    kw = dict(seed=21, max_iter=[500, 4])
    with quiet_stdout():
        so = oracle_py.RefMg.setup(ours, L, L, U, 0.02, [4, 2], [4, 4], **kw)
        sr = oracle_py.RefMg.setup(orc, L, L, U, 0.02, [4, 2], [4, 4], **kw)
    for v in range(4):
        assert np.array_equal(so.null(0, v), sr.null(0, v))
        assert rel_err(so.null(1, v), sr.null(1, v)) < 1e-5
    assert so.null_counts() == sr.null_counts()
    for lvl, tol in ((1, 1e-13), (2, 1e-5)):
        (c, h, s), (cr, hr, srr) = so.stencil(lvl), sr.stencil(lvl)
        assert rel_err(c, cr) < tol and rel_err(h, hr) < tol and np.array_equal(s, srr)
    so.set_precond()
    sr.set_precond()
    with quiet_stdout():
        xo, io = so.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
        xr, ir = sr.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
    assert io["success"] and ir["success"] and abs(io["iter"] - ir["iter"]) <= max(2, 0.25 * ir["iter"])
    assert rel_err(xo, xr) < 1e-4
