#!/usr/bin/env python
"""bench.py -- headline benchmark of the solver hot path (BASELINE.json metric:
"staggered Dirac apply GB/s (% HBM peak); CG solve time at 4096^2, 1-8 GPUs").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--L 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one complete CGNE solve -- minv_vector_cg on D^dag D (square_staggered_normal_u1), the
2-D U(1) staggered operator, complex<double>, mass 0.1, tol 1e-10, zero initial guess -- on an
L x (L*N) lattice, y-slab-sharded over N GPUs (weak scaling: L x L sites per GPU; N=1 is the
4096^2 configuration the metric is quoted on).

  value  : algorithmic GB/s of the whole solve with all inputs resident in HBM
           (bytes = SURVEY section 8 d-bytes accounting, see algorithmic_bytes()).
  e2e    : same metric through the reference-facing call with HOST (pinned) buffers:
           upload phi, phi0 -> solve -> download phi inside the timed region.
  roofline : the step's dominant kernel (one-pass D^dag D with the fused CG direction update, 96 B/site),
           every launch timed with CUDA events on the library's stream; the x/r update, the staggered D
           apply alone (64 B/site) and the plain one-pass D^dag D are listed beside it.
  cpu_baseline : the reference's CPU code (oracle/_ref, else the port) on a bounded sample, 1 core.

--impl reference times the reference's own CPU implementation (same metric/unit/config).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
MASS, TOL, BETA, SEED = 0.1, 1e-10, 6.0, 1337


# ------------------------------------------------------------------------------------------ inputs
def gauge_rows(X, rows, seed=SEED, beta=BETA):
    """synthetic U(1) links (theta ~ N(0,1/beta), u1_utils.cpp:92) for the given global rows, generated
    row by row from (seed, y) so that every slab decomposition sees the same field.
    Layout: [row][x][mu] complex, the reference layout lattice[y*X*2 + x*2 + mu]."""
    out = np.empty((len(rows), 2 * X), dtype=np.complex128)
    for i, y in enumerate(rows):
        th = np.random.default_rng([seed, int(y)]).standard_normal(2 * X) / np.sqrt(beta)
        out[i].real, out[i].imag = np.cos(th), np.sin(th)
    return out.reshape(-1)


def rhs_rows(X, rows, seed=SEED + 1):
    out = np.empty((len(rows), X), dtype=np.complex128)
    for i, y in enumerate(rows):
        g = np.random.default_rng([seed, int(y)]).standard_normal(2 * X)
        out[i].real, out[i].imag = g[:X], g[X:]
    return out.reshape(-1)


def algorithmic_bytes(V, iterations):
    """Fused-minimum HBM traffic of one minv_vector_cg call on D^dag D (SURVEY section 8 d-bytes, B/site):
    set-up  norm(b) 16 + apply 128 + (r = b - Ap) 48 + copy 32 + apply 128 + norm(r) 16      = 368
    per iteration that continues  [x,r update + |r|^2] 96 + [p update . D] 96 + [D^dag . <p,Ap>] 80 = 272
    the last iteration only does the x,r update                                               =  96
    true residual  apply 128 + diffnorm 32                                                    = 160"""
    return float(V) * (368.0 + 272.0 * max(iterations - 1, 0) + 96.0 + 160.0)


def ncu_traffic(kernel, L):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py), or None when there is no capture of this
    kernel at this lattice size"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t.get(str(L), {}).get(kernel, {}).get("bytes")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(L, m, seed, which="best"):
    """one bounded sample of the workload on the host: minv_vector_cg(max_iter=m) on D^dag D, L x L"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    orc = oracle_py.load(which)
    rows = list(range(L))
    U = gauge_rows(L, rows, seed)
    b = rhs_rows(L, rows, seed + 1)
    op = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    t0 = time.perf_counter()
    x, info = orc.solve("CG", op, b, max_iter=m, eps=TOL)
    dt = time.perf_counter() - t0
    return dt, info["iter"], orc.kind


def _cpu_worker(args):
    L, m, seed, reps = args
    out = []
    for _ in range(reps):
        out.append(cpu_sample(L, m, seed)[:2])
    return out


def reference_arm(args, rank, world):
    """the reference's own CPU implementation of the path, all host threads it can use: the code is
    serial (no threads anywhere in the reference), so `cores` independent replicas run side by side."""
    if rank != 0:
        return
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    kind = oracle_py.load("best").kind
    steps, warm = args.steps, args.warmup
    L = args.L if (steps + warm) <= 16 else min(args.L, 2048)
    m = 3
    ncpu = os.cpu_count() or 1
    try:
        import psutil
        mem_gb = psutil.virtual_memory().available / 2**30
    except Exception:
        mem_gb = 16.0
    per_replica_gb = 16.0 * L * L * 12 / 2**30 + 0.5   # links + ~8 work vectors + numpy temporaries
    cores = int(max(1, min(ncpu, 8, mem_gb * 0.6 // per_replica_gb)))
    V = L * L
    with mp.get_context("fork").Pool(cores) as pool:
        def one_step():
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(L, m, SEED + 17 * i, 1) for i in range(cores)])
            wall = time.perf_counter() - t0
            slow = max(r[0][0] for r in res)
            its = res[0][0][1]
            return wall, slow, its
        for _ in range(min(warm, 1)):  # one warm-up pass is enough to page the code in
            one_step()
        tot_solve, tot_bytes = 0.0, 0.0
        for _ in range(steps):
            wall, slow, its = one_step()
            tot_solve += slow                      # time of the solver calls only (input generation excluded)
            tot_bytes += cores * algorithmic_bytes(V, its)
    value = tot_bytes / tot_solve / 1e9
    sample = ("minv_vector_cg(max_iter=%d) on D^dag D, %dx%d, %d independent replicas (the reference is serial), "
              "solver time only" % (m, L, L, cores))
    line = {"impl": "reference", "metric": "staggered_cgne_solve_algorithmic_GBps", "value": value, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * tot_solve / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex<f64>",
            "data": "synthetic",
            "config": {"workload": "CGNE solve: minv_vector_cg on square_staggered_normal_u1, %dx%d per GPU, "
                                   "complex<double>, m=0.1, tol 1e-10 (bounded CPU sample)" % (args.L, args.L)},
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--Y", type=int, default=0, help="total rows of the lattice (strong scaling: fixed L x Y "
                                                      "split over the GPUs); default L rows per GPU (weak)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--apply-reps", type=int, default=50)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from __graft_entry__ import _load_pkg
    glb = _load_pkg()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = glb.Context(device=local)
    if world > 1:
        ctx.init_comm_from_torch()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    L = args.L
    X, Y = L, (args.Y if args.Y > 0 else L * world)
    if Y % world != 0:
        raise SystemExit("--Y must be a multiple of the number of GPUs")
    y0, Yloc = ctx.slab_bounds(Y)
    V_local, V_global = X * Yloc, X * Y
    # this rank's rows plus two periodic ghost rows on each side (glb_op_create_staggered_local)
    rows = [(y0 - 2 + Y) % Y, (y0 - 1 + Y) % Y] + list(range(y0, y0 + Yloc)) + [(y0 + Yloc) % Y, (y0 + Yloc + 1) % Y]
    links_local = gauge_rows(X, rows)
    b_local = rhs_rows(X, rows[2:-2])
    opN = ctx.staggered_local(links_local, X, Y, MASS, glb.STAG_NORMAL)
    opD = ctx.staggered_local(links_local, X, Y, MASS, 0)
    opDd = ctx.staggered_local(links_local, X, Y, MASS, glb.STAG_DAGGER)
    b = ctx.vector(V_local).upload(b_local)
    bp = ctx.vector(V_local)
    opDd.apply(bp, b)          # CGNE right-hand side D^dag b
    x = ctx.vector(V_local)
    ctx.sync()

    # ---- roofline leg: the staggered D apply kernel alone, events on the library's stream
    out = ctx.vector(V_local)
    stream_evt = stream
    for _ in range(5):
        opD.apply(out, b)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.apply_reps):
        opD.apply(out, b)
    e1.record(stream)
    barrier()
    apply_ms = max_over_ranks(e0.elapsed_time(e1)) / args.apply_reps
    apply_gbps = 64.0 * V_local / (apply_ms * 1e-3) / 1e9          # per GPU: 16 psi + 32 links + 16 out
    # the one-pass D^dag D kernel (the CG's dominant kernel without the fused direction update): 64 B/site
    for _ in range(3):
        opN.apply(out, b)
    barrier()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record(stream)
    for _ in range(args.apply_reps):
        opN.apply(out, b)
    n1.record(stream)
    barrier()
    normal_ms = max_over_ranks(n0.elapsed_time(n1)) / args.apply_reps
    normal_gbps = 64.0 * V_local / (normal_ms * 1e-3) / 1e9

    def solve_resident():
        x.zero()
        return ctx.solve("CG", opN, x, bp, max_iter=5000, eps=TOL)

    # ---- warm-up, then exactly K timed steps bracketed by barrier + synchronize
    for _ in range(args.warmup):
        info = solve_resident()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = ctx.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        info = solve_resident()
    ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launches() - launches0
    iters = info["iter"]
    try:
        peak_early = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak_early = 6650.0
    step_bytes = algorithmic_bytes(V_global, iters)
    value = step_bytes * args.steps / (ms_total * 1e-3) / 1e9
    true_rel = float(np.sqrt(info["resSq"]))  # |b - A x| (absolute); made relative below
    bnorm = float(np.sqrt(ctx.norm2sq(bp)))

    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline leg: the step's own kernels, each launch bracketed by CUDA events on the library's
    # stream (glb_prof_*), on one more solve of the same workload right after the timed region.  The CG
    # loop launches [x/r update][direction + D^dag D] per iteration; launches enqueued past the stopping
    # point return at once and are not counted.
    ctx.prof_enable(True)
    info_p = solve_resident()
    ctx.prof_enable(False)
    it_p = info_p["iter"]
    t_fused = ctx.prof_read(1)[:max(it_p - 1, 0)]
    t_upd = ctx.prof_read(3)[:it_p]
    fused_ms = max_over_ranks(sum(t_fused) / max(len(t_fused), 1))
    upd_ms = max_over_ranks(sum(t_upd) / max(len(t_upd), 1))
    fused_gbps = 96.0 * V_local / (fused_ms * 1e-3) / 1e9 if t_fused else 0.0   # R r,p,U 64 + W p,Ap 32
    upd_gbps = 96.0 * V_local / (upd_ms * 1e-3) / 1e9 if t_upd else 0.0         # R x,p,r,Ap 64 + W x,r 32
    step_ms = ms_total / args.steps
    fused_share = sum(t_fused) / step_ms if step_ms > 0 else 0.0
    upd_share = sum(t_upd) / step_ms if step_ms > 0 else 0.0

    # ---- end to end: the reference-facing call with host (pinned) buffers
    hx, hb = ctx.pinned(V_local), ctx.pinned(V_local)
    hb[:] = bp.download()
    e2e_iters = iters
    if world == 1:
        ctx.cache_operators(True)   # gauge field stays resident between solves (it is the "model"); vectors travel
        desc = ctx._desc("STAG_NORMAL_U1", X, Y, mass=MASS, links=links_local[4 * X:4 * X + 2 * X * Y])

        def solve_e2e():
            return ctx.host_solve("CG", desc, hx, hb, max_iter=5000, eps=TOL)
    else:
        def solve_e2e():   # slab runs: same copies, device-level call (the host-vector entry point is single-rank)
            x.upload(hx)
            bp.upload(hb)
            r = ctx.solve("CG", opN, x, bp, max_iter=5000, eps=TOL)
            x.download(hx)
            return r
    for _ in range(2):
        hx[:] = 0
        solve_e2e()
    e2e_s = 0.0
    for _ in range(args.steps):
        hx[:] = 0          # preparing the caller's initial guess is host work outside the call: not timed
        barrier()
        t0 = time.perf_counter()
        e2e_info = solve_e2e()   # synchronous: returns after the solution is back in host memory
        e2e_s += time.perf_counter() - t0
    e2e_s = max_over_ranks(e2e_s)
    e2e_iters = e2e_info["iter"]
    e2e_value = algorithmic_bytes(V_global, e2e_iters) * args.steps / e2e_s / 1e9
    if world == 1:
        ctx.cache_operators(False)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    cpu = None
    if not args.no_cpu and world == 1 and Y == L:
        # bounded sample of the same workload on the host: minv_vector_cg(max_iter=3), same arrays
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py
        orc = oracle_py.load("best")
        Lc = L
        Uc = links_local[4 * X:4 * X + 2 * X * Y] if Lc == L else gauge_rows(Lc, list(range(Lc)))
        bc = hb.copy() if Lc == L else rhs_rows(Lc, list(range(Lc)))
        oop = orc.op("STAG_NORMAL_U1", Lc, Lc, mass=MASS, links=Uc)
        t0 = time.perf_counter()
        _, cinfo = orc.solve("CG", oop, bc, max_iter=3, eps=TOL)
        ct = time.perf_counter() - t0
        v = np.ascontiguousarray(bc)
        oD = orc.op("STAG_U1", Lc, Lc, mass=MASS, links=Uc)
        t0 = time.perf_counter()
        oD.apply(v)
        cat = time.perf_counter() - t0
        cpu = {"value": algorithmic_bytes(Lc * Lc, cinfo["iter"]) / ct / 1e9, "unit": "GB/s", "cores": 1,
               "kind": orc.kind, "host_cpus": os.cpu_count(),
               "sample": "minv_vector_cg(max_iter=3) on D^dag D, %dx%d, the arrays uploaded to the GPU; %.2f s; "
                         "one square_staggered_u1 apply %.3f s = %.2f GB/s" % (Lc, Lc, ct, cat, 64.0 * Lc * Lc / cat / 1e9)}

    line = {
        "metric": "staggered_cgne_solve_algorithmic_GBps", "value": value, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": ("strong" if args.Y > 0 else "weak"), "vs_baseline": None,
        "dtype": "complex<f64>", "data": "synthetic",
        "config": {"workload": "CGNE solve: minv_vector_cg on square_staggered_normal_u1 (D^dag D), 2-D U(1) staggered, "
                               "%dx%d per GPU (global %dx%d, y-slabs), complex<double>, m=0.1, tol 1e-10, zero guess"
                               % (X, Y // world, X, Y),
                   "lattice": [X, Y], "iterations": iters, "true_rel_residual": true_rel / bnorm,
                   "l2": "working set %.1f GB per GPU >> 126 MB L2: no flush needed" % (V_local * 16 * 9 / 1e9),
                   "gauge": "gauss U(1), beta=6, per-row numpy seed %d" % SEED,
                   "comm": ("single GPU" if world == 1 else
                            ("NVLink peer memory: halo rows and rank sums written by the kernels themselves"
                             if ctx.p2p else "NCCL send/recv + allreduce on the compute stream")),
                   "bytes_model": "SURVEY 8 d-bytes fused minimum: 368 + 272*(it-1) + 96 + 160 B/site; the one-pass "
                                  "D^dag D kernel actually moves 192 B/site/iteration, so value/peak may exceed 1",
                   "actual_traffic_frac_of_peak": (192.0 * V_local * iters * args.steps / (ms_total * 1e-3) / 1e9) / peak_early},
        "solve_time_s": ms_total / args.steps * 1e-3, "iterations_per_s": iters * args.steps / (ms_total * 1e-3),
        "frac_of_hbm_peak": value / world / peak,
        "clocks": sampler.summary(),
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm",
                     "kernel": "normal_kernel<fused> (p = r + beta p ; Ap = D^dag D p ; <p,Ap> in one pass, 96 B/site): "
                               "the step's dominant kernel",
                     "achieved": fused_gbps, "peak": peak, "unit": "GB/s", "frac": fused_gbps / peak,
                     "traffic": ncu_traffic("normal_kernel_fused", L), "ms_per_launch": fused_ms,
                     "algorithmic_bytes_per_launch": 96.0 * V_local, "launches_timed": len(t_fused),
                     "share_of_step": fused_share, "peak_source": peak_src, "per_gpu": True,
                     "how": "CUDA events around every launch inside one more solve right after the timed region "
                            "(slab runs: includes waiting for the neighbours' halo rows)",
                     "other_kernels": [
                         {"kernel": "cg_update_kernel (x += alpha p ; r -= alpha Ap ; |r|^2, 96 B/site)",
                          "achieved": upd_gbps, "frac": upd_gbps / peak, "ms_per_launch": upd_ms,
                          "traffic": ncu_traffic("cg_update_kernel", L), "share_of_step": upd_share,
                          "launches_timed": len(t_upd)},
                         {"kernel": "stag_kernel (staggered D apply alone, 64 B/site; the metric's 'Dirac apply GB/s')",
                          "achieved": apply_gbps, "frac": apply_gbps / peak, "ms_per_launch": apply_ms,
                          "traffic": ncu_traffic("stag_kernel", L), "how": "loop of %d applies" % args.apply_reps},
                         {"kernel": "normal1_kernel (D^dag D alone in one pass, one site per thread, 64 B/site)",
                          "achieved": normal_gbps, "frac": normal_gbps / peak, "ms_per_launch": normal_ms,
                          "traffic": ncu_traffic("normal_kernel", L), "how": "loop of %d applies" % args.apply_reps}]},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(2 * 16 * V_local),
                "d2h_bytes_per_step": int(16 * V_local), "s_per_step": e2e_s / args.steps, "iterations": e2e_iters,
                "note": "host vectors travel every step (pinned); the gauge field is uploaded once and stays resident"},
        "wall_s_timed_region": t_wall,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


if __name__ == "__main__":
    main()
