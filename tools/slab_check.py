#!/usr/bin/env python
"""Multi-GPU parity check of the y-slab path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_check.py [L]

Every rank holds the same seeded global field on the host (small L), builds its slab operator through
the C ABI, and the slab results gathered on rank 0 are compared with the CPU oracle: applies bit for
bit, solves by iteration count (+-2 %) and true residual.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import torch.distributed as dist
    import oracle_py
    from __graft_entry__ import _load_pkg
    glb = _load_pkg()
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = glb.Context(device=local)
    ctx.init_comm_from_torch()
    orc = oracle_py.load("best")
    X, Y = L, L
    r = orc.rng(1337)
    U = r.gauss_gauge_u1(X, Y, 6.0)
    b = r.gaussian(X * Y)
    y0, Yloc = ctx.slab_bounds(Y)
    sl = slice(y0 * X, (y0 + Yloc) * X)
    fails = []

    def gather(local_arr):
        t = torch.from_numpy(np.ascontiguousarray(local_arr).view(np.float64)).cuda()
        sizes = [0] * world
        allsz = [None] * world
        dist.all_gather_object(allsz, t.numel())
        outs = [torch.empty(n, dtype=torch.float64, device="cuda") for n in allsz]
        dist.all_gather(outs, t) if len(set(allsz)) == 1 else [dist.broadcast(outs[i] if i != rank else t, i) for i in range(world)]
        if len(set(allsz)) != 1:
            outs[rank] = t
        return np.concatenate([o.cpu().numpy() for o in outs]).view(np.complex128)

    def check(name, cond, extra=""):
        if rank == 0:
            print("%-48s %s %s" % (name, "ok" if cond else "FAIL", extra), flush=True)
        if not cond:
            fails.append(name)

    # ---- applies: slab result == global oracle rows, bit for bit
    for kind, flags in [("STAG_U1", 0), ("STAG_DAGGER_U1", 1), ("STAG_GAMMA5_U1", 2), ("STAG_NORMAL_U1", 4)]:
        op = ctx.staggered(U, X, Y, 0.1, flags)          # global host array; each rank uploads its slab
        got = gather(op.apply_host(b[sl]))
        want = orc.op(kind, X, Y, mass=0.1, links=U).apply(b)
        check("apply %s" % kind, np.array_equal(got, want), "max diff %.1e" % np.abs(got - want).max())
    rows = [(y0 - 2 + Y) % Y, (y0 - 1 + Y) % Y] + list(range(y0, y0 + Yloc)) + [(y0 + Yloc) % Y, (y0 + Yloc + 1) % Y]
    Uloc = U.reshape(Y, 2 * X)[rows].reshape(-1)
    got = gather(ctx.staggered_local(Uloc, X, Y, 0.1, 0).apply_host(b[sl]))
    check("apply STAG_U1 (slab-local links)", np.array_equal(got, orc.op("STAG_U1", X, Y, mass=0.1, links=U).apply(b)))
    got = gather(ctx.laplace_u1(U, X, Y, 0.3).apply_host(b[sl]))
    check("apply LAPLACE_U1", np.array_equal(got, orc.op("LAPLACE_U1", X, Y, mass=0.3, links=U).apply(b)))
    got = gather(ctx.laplace(X, Y, 1, 4.01 + 1j, np.complex128).apply_host(b[sl]))
    check("apply LAPLACE_IMAG", np.array_equal(got, orc.op("LAPLACE_IMAG", X, Y, mass=0.01).apply(b)))
    nc = 4
    rg = np.random.default_rng(3)
    rc = lambda n: rg.standard_normal(n) + 1j * rg.standard_normal(n)
    V = X * Y
    cl, hp, tl, v = rc(V * nc * nc), rc(4 * V * nc * nc), rc(8 * V * nc * nc), rc(V * nc)
    for two in (None, tl):
        got = gather(ctx.stencil2d(cl, hp, two, X, Y, nc, shift=0.3, eo_shift=0.2j, dof_shift=0.1).apply_host(
            v[y0 * X * nc:(y0 + Yloc) * X * nc]))
        want = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, two_link=two, shift=0.3, eo_shift=0.2j,
                      dof_shift=0.1).apply(v)
        check("apply STENCIL nc=4 two_link=%s" % (two is not None), np.array_equal(got, want))

    # ---- partial applies, composite views and the prepare / reconstruct passes on slabs (reference
    # library only).  The paired thread mapping keys on the GLOBAL row parity, so y0 matters here.
    if orc.kind == "reference":
        slc = slice(y0 * X * nc, (y0 + Yloc) * X * nc)
        sop = ctx.stencil2d(cl, hp, None, X, Y, nc, shift=0.3)
        oop = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, shift=0.3)
        w = rc(V * nc)
        dv, dw, out = ctx.vector(Yloc * X * nc).upload(v[slc]), ctx.vector(Yloc * X * nc).upload(w[slc]), ctx.vector(Yloc * X * nc)
        for part in ("EO", "OE", "TB", "BT"):
            sop.apply_part(part, out, dv)
            check("stencil part %s" % part, np.array_equal(gather(out.download()), oop.apply_part(part, v)))
        for view in ("M2MDEODOE", "M2MDTBDBT", "NORMAL_EO", "NORMAL_TB", "DAGGER_EO", "DAGGER_TB"):
            vo = sop.view(view)
            vo.apply(out, dv)
            want = orc.op("STENCIL", X, Y, Nc=nc, clover=cl, hopping=hp, shift=0.3, view=view).apply(v)
            check("stencil view %s" % view, np.array_equal(gather(out.download()), want))
            vo.destroy()
        for tb in (0, 1):
            sop.prec_prepare(tb, out, dv)
            check("stencil prec_prepare tb=%d" % tb, np.array_equal(gather(out.download()), oracle_py.ref_stencil_prec(orc, oop, tb, v)))
            sop.prec_reconstruct(tb, out, dv, dw)
            check("stencil prec_reconstruct tb=%d" % tb,
                  np.array_equal(gather(out.download()), oracle_py.ref_stencil_prec(orc, oop, tb, v, w)))
        # e/o-preconditioned solve of D x = b through the nc = 1 stencil of the staggered operator
        import mg_setup
        cl0, hp0, _ = mg_setup.staggered_stencil(U, X, Y, 0.0)
        st = ctx.stencil2d(cl0, hp0, None, X, Y, 1, shift=0.1)
        ost = orc.op("STENCIL_FROM_STAG", X, Y, mass=0.1, links=U)
        ostm = orc.op("STENCIL_FROM_STAG", X, Y, mass=0.1, links=U, view="M2MDEODOE")
        db, be, xe, xf = (ctx.vector(Yloc * X) for _ in range(4))
        db.upload(b[sl])
        st.prec_prepare(0, be, db)
        xe.zero()
        stm = st.view("M2MDEODOE")
        info = ctx.solve("CG", stm, xe, be, max_iter=5000, eps=1e-10)
        st.prec_reconstruct(0, xf, xe, db)
        _, want = orc.solve("CG", ostm, oracle_py.ref_stencil_prec(orc, ost, 0, b), max_iter=5000, eps=1e-10)
        xg = gather(xf.download())
        rr = np.linalg.norm(orc.op("STAG_U1", X, Y, mass=0.1, links=U).apply(xg) - b) / np.linalg.norm(b)
        check("solve e/o-preconditioned CG through the stencil", abs(info["iter"] - want["iter"]) <= max(1, round(0.02 * want["iter"]))
              and rr < 1e-8, "iter %d (oracle %d) true rel res %.2e" % (info["iter"], want["iter"], rr))

    # ---- multigrid on slabs: set-up on the device (null vectors, block orthonormalisation, Galerkin product with the
    # neighbours' boundary rows) and the V-cycle-preconditioned outer solve, next to the reference's own set-up + solve
    if orc.kind == "reference" and Y % (8 * world) == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from mg_common import quiet_stdout
        import mg_setup
        mass_mg = 0.05
        cl0, hp0, _ = mg_setup.staggered_stencil(U, X, Y, 0.0)
        fine = ctx.stencil2d(cl0, hp0, None, X, Y, 1, shift=mass_mg)

        def rel(a, bb):
            return float(np.linalg.norm(a - bb) / max(np.linalg.norm(bb), 1e-300))

        def gather_planes(local, nplanes):
            per = local.size // nplanes
            return np.concatenate([gather(local[d * per:(d + 1) * per]) for d in range(nplanes)])

        for blocks, nvecs in (([4], [4]), ([4, 2], [4, 4])):
            tag = "blocks %s" % blocks
            # (a) three smoothing iterations: the hierarchy itself can be compared (see tests/test_mg_setup_gpu.py)
            kw = dict(seed=17, max_iter=3)
            with quiet_stdout():
                ref = oracle_py.RefMg.setup(orc, X, Y, U, mass_mg, blocks, nvecs, **kw)
            mg = ctx.multigrid_setup(fine, X, Y, blocks, nvecs, **kw)
            worst = max(rel(gather(mg.null_vector(0, v)), ref.null(0, v)) for v in range(nvecs[0]))
            check("MG set-up on slabs, %s: top-level null vectors" % tag, worst < 1e-5, "max rel err %.1e" % worst)
            cl, hp, sh = mg.level_stencil(1)
            clr, hpr, shr = ref.stencil(1)
            e1, e2 = rel(gather(cl), clr), rel(gather_planes(hp, 4), hpr)
            check("MG set-up on slabs, %s: Galerkin coarse stencil" % tag, e1 < 1e-5 and e2 < 1e-5 and sh[0] == complex(mass_mg),
                  "clover %.1e hopping %.1e" % (e1, e2))
            check("MG set-up on slabs, %s: null-vector applies" % tag, mg.counts()["nullvectors"] == ref.null_counts())
            mg.destroy()
            # (b) the driver's defaults: outer VPGCR(64) + V cycle
            with quiet_stdout():
                ref = oracle_py.RefMg.setup(orc, X, Y, U, mass_mg, blocks, nvecs, seed=23)
                ref.set_precond()
                xo, want = ref.vpgcr(b, max_iter=1000, eps=5e-7, restart_freq=64)
            mg = ctx.multigrid_setup(fine, X, Y, blocks, nvecs, seed=23)
            mg.set()
            x = ctx.vector(Yloc * X).zero()
            got = mg.vpgcr(x, ctx.vector(Yloc * X).upload(b[sl]), max_iter=1000, eps=5e-7, restart_freq=64)
            xg = gather(x.download())
            rr = np.linalg.norm(orc.op("STAG_U1", X, Y, mass=mass_mg, links=U).apply(xg) - b) / np.linalg.norm(b)
            check("MG solve on slabs, %s: VPGCR(64) + V cycle" % tag,
                  got["success"] and abs(got["iter"] - want["iter"]) <= max(2, 0.25 * want["iter"]) and rr < 5e-7 * 1.0001,
                  "iter %d (reference %d) true rel res %.2e" % (got["iter"], want["iter"], rr))
            mg.destroy()

    # ---- reductions
    xv = ctx.vector(Yloc * X).upload(b[sl])
    d = ctx.dot(xv, xv)
    check("dot over slabs", abs(d - np.vdot(b, b)) <= 1e-12 * abs(np.vdot(b, b)))

    # ---- solves
    oD = orc.op("STAG_U1", X, Y, mass=0.1, links=U)
    oN = orc.op("STAG_NORMAL_U1", X, Y, mass=0.1, links=U)
    bprime = orc.op("STAG_DAGGER_U1", X, Y, mass=0.1, links=U).apply(b)
    D = ctx.staggered(U, X, Y, 0.1, 0)
    N = ctx.staggered(U, X, Y, 0.1, glb.STAG_NORMAL)
    for name, solver, op, oop, rhs, kw in [
            ("CGNE", "CG", N, oN, bprime, dict(eps=1e-10)), ("CR", "CR", N, oN, bprime, dict(eps=1e-10)),
            ("BiCGStab", "BICGSTAB", D, oD, b, dict(eps=1e-10)), ("BiCGStab-4", "BICGSTAB_L", D, oD, b, dict(eps=1e-10, l=4)),
            ("GCR(20)", "GCR_RESTART", D, oD, b, dict(eps=1e-8, restart_freq=20)),
            ("GMRES(20)", "GMRES_RESTART", D, oD, b, dict(eps=1e-8, restart_freq=20))]:
        _, want = orc.solve(solver, oop, rhs, max_iter=5000, **kw)
        x = ctx.vector(Yloc * X).zero()
        bd = ctx.vector(Yloc * X).upload(rhs[sl])
        info = ctx.solve(solver, op, x, bd, max_iter=5000, **kw)
        xg = gather(x.download())
        rr = np.linalg.norm(oop.apply(xg) - rhs) / np.linalg.norm(rhs)
        tol_it = 0.10 if solver.startswith("BICGSTAB") else 0.02   # BiCGStab is chaotic (see tests/test_solvers_gpu.py)
        ok = abs(info["iter"] - want["iter"]) <= max(1, round(tol_it * want["iter"])) and rr < kw["eps"] * 1.000001
        check("solve %s" % name, ok, "iter %d (oracle %d) true rel res %.2e" % (info["iter"], want["iter"], rr))
    # ---- the single-kernel CG iteration on slabs (cgstep.cu): the iterate after m = 1..5 iterations equals the
    # reference's to rounding -- every row of r, p, q next to a slab edge has then crossed NVLink m times -- on the
    # square lattice and on a thin one (four rows per rank and a partial strip in x)
    for (Xc, Yc) in ((X, Y), (132, 4 * world), (16, 8 * world)):
        rr_ = orc.rng(5)
        Uc = rr_.gauss_gauge_u1(Xc, Yc, 6.0)
        bc = rr_.gaussian(Xc * Yc)
        x0c = rr_.gaussian(Xc * Yc)
        oNc = orc.op("STAG_NORMAL_U1", Xc, Yc, mass=0.1, links=Uc)
        Nc_ = ctx.staggered(Uc, Xc, Yc, 0.1, glb.STAG_NORMAL)
        y0c, Ylc = ctx.slab_bounds(Yc)
        slc_ = slice(y0c * Xc, (y0c + Ylc) * Xc)
        worst, its = 0.0, []
        for m in (1, 2, 3, 5):
            want_x, winfo = orc.solve("CG", oNc, bc, x0=x0c, max_iter=m, eps=1e-30)
            xd = ctx.vector(Ylc * Xc).upload(x0c[slc_])
            rep = ctx.cg_device(Nc_, xd, ctx.vector(Ylc * Xc).upload(bc[slc_]), max_iter=m, eps=1e-30)
            xg = gather(xd.download())
            worst = max(worst, float(np.linalg.norm(xg - want_x) / np.linalg.norm(want_x)))
            its.append(rep["iterations"])
        check("single-kernel CG on slabs, %dx%d: first iterations" % (Xc, Yc), worst < 1e-13 and its == [1, 2, 3, 5],
              "max rel err %.1e, beta prediction err %.1e" % (worst, ctx.cg_last_pred_err()))
        Nc_.destroy()
    shifts = [0.0, 0.01, 0.05, 0.25]
    xs = [ctx.vector(Yloc * X) for _ in shifts]
    info, _ = ctx.solve_cg_m(N, xs, ctx.vector(Yloc * X).upload(bprime[sl]), shifts, max_iter=5000, eps=1e-10)
    _, want, _ = orc.solve_cg_m(oN, bprime, shifts, max_iter=5000, eps=1e-10)
    worst = 0.0
    for s, xd in zip(shifts, xs):
        xg = gather(xd.download())
        worst = max(worst, np.linalg.norm(oN.apply(xg) + s * xg - bprime) / np.linalg.norm(bprime))
    check("solve CG-M", abs(info["iter"] - want["iter"]) <= max(1, round(0.02 * want["iter"])) and worst < 1e-10 * 1.000001,
          "iter %d (oracle %d) worst rel res %.2e" % (info["iter"], want["iter"], worst))
    dist.barrier()
    if rank == 0:
        print("SLAB CHECK %s (%d ranks, %dx%d)" % ("PASSED" if not fails else "FAILED: %s" % fails, world, X, Y), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
