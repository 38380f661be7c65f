#!/usr/bin/env python
"""bench.py -- headline benchmark of the solver hot path (BASELINE.json metric:
"staggered Dirac apply GB/s (% HBM peak); CG solve time at 4096^2, 1-8 GPUs").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--L 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one complete CGNE solve -- minv_vector_cg on D^dag D (square_staggered_normal_u1), the
2-D U(1) staggered operator, complex<double>, mass 0.1, tol 1e-10, zero initial guess -- on an
L x (L*N) lattice, y-slab-sharded over N GPUs (weak scaling: L x L sites per GPU; N=1 is the
4096^2 configuration the metric is quoted on).  The same JSON line carries a `strong` block: the same
solve on a FIXED global 4096^2 and 8192^2 lattice over the N GPUs with its own one-GPU time (BASELINE
config 4 and the metric's "CG solve time at 4096^2, 1-8 GPUs").

  value  : GB/s of the whole solve on the bytes the kernels MOVE (solve_bytes(): 160 B/site per CG iteration
           for the single-kernel iteration), inputs resident in HBM; value/peak <= 1.
           value_algorithmic_model is the same time on SURVEY section 8's 272 B/site textbook-fused model.
  e2e    : same metric through the reference-facing call minv_vector_cg(host pointers, square_staggered_normal_u1)
           with pinned HOST buffers: upload phi, phi0 -> solve -> download phi inside the timed region, every rank.
  roofline : the step's dominant kernel (cg_step_kernel: the whole CG iteration in one pass, 160 B/site), every launch
           timed with CUDA events on the library's stream; the staggered D apply (the metric's "Dirac apply GB/s"),
           the plain one-pass D^dag D and the two kernels of the two-kernel CG loop are listed beside it.
  cpu_baseline : the reference's CPU code (oracle/_ref, else the port) on a bounded sample, 1 core, same arrays.

--impl reference times the reference's own CPU implementation on the same configuration (same L, same input
arrays, same metric and byte count), one bounded sample (max_iter = 3) per step.
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
METRIC = "staggered_cgne_solve_GBps"


# ------------------------------------------------------------------------------------------ inputs
def gauge_rows(X, rows, seed=SEED, beta=BETA):
    """synthetic U(1) links (theta ~ N(0,1/beta), u1_utils.cpp:92) for the given global rows, generated
    row by row from (seed, y) so that every slab decomposition sees the same field (multi-GPU weak scaling,
    where no rank can afford the global mt19937 stream).  Layout: [row][x][mu] complex."""
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


def slab_rows(y0, Yloc, Y):
    """this rank's rows plus two periodic ghost rows on each side (glb_op_create_staggered_local)"""
    return [(y0 - 2 + Y) % Y, (y0 - 1 + Y) % Y] + list(range(y0, y0 + Yloc)) + [(y0 + Yloc) % Y, (y0 + Yloc + 1) % Y]


def solve_bytes(V, iterations):
    """HBM bytes one minv_vector_cg call on D^dag D MOVES with the single-kernel iteration (B/site):
    set-up    norm(b) 16 + one-pass apply 64 + (r = b - Ax) 48 + two zero fills 32          = 160
    steps     (iterations + 1) passes of cg_step_kernel: R r,q,p,x,Ux,Uy 96 + W r,p,q,x 64  = 160 each
    true residual  one-pass apply 64 + diffnorm 32                                          =  96"""
    return float(V) * (160.0 + 160.0 * (iterations + 1) + 96.0)


def model_bytes(V, iterations):
    """SURVEY section 8 d-bytes of the textbook-fused two-kernel schedule (two-pass D^dag D): 368 set-up +
    272 per continued iteration + 96 last update + 160 true residual"""
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


def golden_large():
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_large.json")))
    except Exception:
        return {}


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


def bind_to_gpu_numa_node(local):
    """run this rank's host threads and place its pinned buffers on the NUMA node its GPU hangs off (the e2e leg
    moves 768 MiB per solve per rank through host memory); best effort, returns a description for the JSON line"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "gpu numa node unknown"
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cpus) & allowed)
        note = "node %d" % node
        if use:
            os.sched_setaffinity(0, use)
            note += ", %d cpus" % len(use)
        else:
            note += ", cpus outside this process's cpuset"
        try:  # memory policy: prefer that node for this process's allocations (set_mempolicy, MPOL_PREFERRED = 1)
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            if libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64)) == 0:
                note += ", mempolicy preferred"
        except Exception:
            pass
        return note
    except Exception as e:  # noqa: BLE001
        return "not bound (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------ CPU arm
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    return oracle_py.load("best")


def cpu_sample(L, m, seed):
    """one bounded sample of the workload on the host: minv_vector_cg(max_iter=m) on D^dag D, L x L, rhs D^dag b,
    inputs = the mt19937(seed) stream of BASELINE.md section 3 (the arrays the GPU arm solves at N=1)"""
    orc = _oracle()
    r = orc.rng(seed)
    U = r.gauss_gauge_u1(L, L, BETA)
    b = r.gaussian(L * L)
    bp = orc.op("STAG_DAGGER_U1", L, L, mass=MASS, links=U).apply(b)
    op = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    t0 = time.perf_counter()
    x, info = orc.solve("CG", op, bp, max_iter=m, eps=TOL)
    dt = time.perf_counter() - t0
    return dt, info["iter"], orc.kind


def _cpu_worker(args):
    L, m, seed, steps = args
    orc = _oracle()
    r = orc.rng(seed)
    U = r.gauss_gauge_u1(L, L, BETA)
    b = r.gaussian(L * L)
    bp = orc.op("STAG_DAGGER_U1", L, L, mass=MASS, links=U).apply(b)
    op = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=U)
    out = []
    for _ in range(steps):
        t0 = time.perf_counter()
        x, info = orc.solve("CG", op, bp, max_iter=m, eps=TOL)
        out.append((time.perf_counter() - t0, info["iter"]))
    return out


def reference_arm(args, rank, world):
    """the reference's own CPU implementation of the path on the configuration of our arm: L x L per GPU, the same
    mt19937(1337) arrays as our N=1 run, same metric and byte count.  The code is serial (no threads anywhere in the
    reference), so `value` is ONE solver instance on one core; `replicas` reports what independent copies on the other
    cores add up to.  Each step is a bounded sample (max_iter = 3; CG's cost per iteration is constant)."""
    if rank != 0:
        return
    import multiprocessing as mp
    kind = _oracle().kind
    steps, warm = args.steps, args.warmup
    L, m = args.L, 3
    V = L * L
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import psutil
        mem_gb = psutil.virtual_memory().available / 2**30
    except Exception:
        mem_gb = 16.0
    per_replica_gb = 16.0 * V * 12 / 2**30 + 0.5   # links + ~8 work vectors + numpy temporaries
    extra = int(max(0, min(ncpu - 1, 7, mem_gb * 0.6 // per_replica_gb - 1)))
    # the measured instance: this process, one core, W warm-up + K timed samples
    res = _cpu_worker((L, m, SEED, min(warm, 1) + steps))[min(warm, 1):]
    tot = sum(t for t, _ in res)
    its = res[0][1]
    value = solve_bytes(V, its) * steps / tot / 1e9
    replicas = None
    if extra > 0:  # what the other cores add when they run independent copies at the same time (2 samples each)
        with mp.get_context("fork").Pool(extra + 1) as pool:
            rr = pool.map(_cpu_worker, [(L, m, SEED + 17 * i, 2) for i in range(extra + 1)])
        slow = max(sum(t for t, _ in r) for r in rr)
        replicas = {"cores": extra + 1, "aggregate_GBps": (extra + 1) * 2 * solve_bytes(V, its) / slow / 1e9}
    sample = ("minv_vector_cg(max_iter=%d) on D^dag D, %dx%d, mt19937(%d) inputs (the arrays of the GPU arm at N=1), "
              "one serial instance, solver time only; GB/s on the GPU arm's byte count (solve_bytes)" % (m, L, L, SEED))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * tot / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex<f64>",
            "data": "synthetic",
            "config": workload_config(L, L * max(args.gpus, 1), max(args.gpus, 1)),
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": kind, "sample": sample,
                             "replicas": replicas, "host_cpus": ncpu},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(X, Y, world):
    """the workload, identical in both arms (--impl ours / reference): what is solved, on what, how it is counted"""
    ws_gb = X * (Y // world) * 16 * 10 / 1e9
    return {"workload": "CGNE solve: minv_vector_cg on square_staggered_normal_u1 (D^dag D), 2-D U(1) staggered, "
                        "%dx%d per GPU (global %dx%d, y-slabs), complex<double>, m=0.1, tol 1e-10, zero guess"
                        % (X, Y // world, X, Y),
            "lattice": [X, Y],
            "l2": ("working set %.1f GB per GPU >> 126 MB L2: no flush needed" % ws_gb if ws_gb > 0.5 else
                   "working set %.2f GB per GPU is comparable to the 126 MB L2: L2-assisted numbers, not HBM bandwidth"
                   % ws_gb),
            "bytes_model": "solve_bytes(): bytes the kernels move = 160 set-up + 160*(iterations+1) + 96 true residual "
                           "B/site (single-kernel CG iteration, 160 B/site per iteration)"}


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
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block")
    ap.add_argument("--no-sweep", action="store_true", help="skip the lattice sweep 256^2 ... 2048^2 (N=1)")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations (N=1)")
    ap.add_argument("--strong-sizes", type=int, nargs="*", default=[4096, 8192])
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

    numa_note = (bind_to_gpu_numa_node(local) if world > 1 and not os.environ.get("BENCH_NO_BIND")
                 else "not bound")
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

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = ("MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks
                else "fallback 6650 GB/s (B200_PROFILING.md)")

    L = args.L
    X, Y = L, (args.Y if args.Y > 0 else L * world)
    if Y % world != 0:
        raise SystemExit("--Y must be a multiple of the number of GPUs")
    y0, Yloc = ctx.slab_bounds(Y)
    V_local, V_global = X * Yloc, X * Y
    rows = slab_rows(y0, Yloc, Y)
    if world == 1:
        # one mt19937(1337) stream, gauss_gauge_u1(beta = 6) then the gaussian right-hand side
        links_global, b_local = ctx.synthetic_inputs(X, Y, SEED, BETA)
        lg = links_global.reshape(Y, 2 * X)
        links_local = np.ascontiguousarray(lg[rows]).reshape(-1)
        gauge_note = "std::mt19937(%d): gauss_gauge_u1(beta=6) then gaussian rhs (BASELINE.md section 3)" % SEED
    else:
        links_local = gauge_rows(X, rows)
        b_local = rhs_rows(X, rows[2:-2])
        # the reference-facing call takes the GLOBAL field (lattice[y*X*2 + x*2 + mu]); a rank only ever reads its own
        # rows and the four ghost rows, so the rest of this array is never touched (nor committed by the kernel)
        links_global = np.empty(2 * X * Y, dtype=np.complex128)
        lg = links_global.reshape(Y, 2 * X)
        ll = links_local.reshape(len(rows), 2 * X)
        for i, y in enumerate(rows):
            lg[y] = ll[i]
        gauge_note = "gauss U(1), beta=6, per-row numpy seed %d (slab runs: no rank draws the global stream)" % SEED
    opN = ctx.staggered_local(links_local, X, Y, MASS, glb.STAG_NORMAL)
    opD = ctx.staggered_local(links_local, X, Y, MASS, 0)
    opDd = ctx.staggered_local(links_local, X, Y, MASS, glb.STAG_DAGGER)
    b = ctx.vector(V_local).upload(b_local)
    bp = ctx.vector(V_local)
    opDd.apply(bp, b)          # CGNE right-hand side D^dag b
    x = ctx.vector(V_local)
    ctx.sync()

    # ---- slab self-check (untimed): D and D^dag D on the 32 rows next to each slab edge against the CPU oracle on a
    # 72-row band (periodic wrap of the band only reaches its outer 2 rows) -- exercises the halo path of this run
    slab_parity = None
    if world > 1:
        slab_parity = slab_self_check(ctx, glb, opD, opN, X, Y, y0, Yloc, b_local, links_local, rows, b)
        t = torch.tensor([0.0 if slab_parity == "ok" else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if t.item() != 0.0:
            slab_parity = "MISMATCH on %d rank(s)%s" % (int(t.item()), "" if slab_parity == "ok" else ": " + slab_parity)

    def time_loop(fn, reps, warm=5):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / reps

    # ---- roofline leg: the staggered D apply kernel alone and the plain one-pass D^dag D (64 B/site each)
    out = ctx.vector(V_local)
    apply_ms = time_loop(lambda: opD.apply(out, b), args.apply_reps)
    apply_gbps = 64.0 * V_local / (apply_ms * 1e-3) / 1e9          # per GPU: 16 psi + 32 links + 16 out
    normal_ms = time_loop(lambda: opN.apply(out, b), args.apply_reps, warm=3)
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
    value = solve_bytes(V_global, iters) * args.steps / (ms_total * 1e-3) / 1e9
    value_model = model_bytes(V_global, iters) * args.steps / (ms_total * 1e-3) / 1e9
    bnorm = float(np.sqrt(ctx.norm2sq(bp)))
    true_rel = float(np.sqrt(info["resSq"])) / bnorm   # |b - A x| recomputed by the shell with one more apply
    pred_err = ctx.cg_last_pred_err()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline leg: the step's own kernels, each launch bracketed by CUDA events on the library's
    # stream (glb_prof_*), on one more solve of the same workload right after the timed region.  Launches
    # enqueued past the stopping point return at once and are not counted.
    ctx.prof_enable(True)
    info_p = solve_resident()
    ctx.prof_enable(False)
    it_p = info_p["iter"]
    t_step = ctx.prof_read(7)
    single_kernel = len(t_step) > 0
    step_ms = ms_total / args.steps
    if single_kernel:
        persistent = (len(t_step) == 1)     # one launch ran all it_p + 1 steps of the solve
        if not persistent:
            t_step = t_step[:it_p + 1]
        steps_per_launch = (it_p + 1) if persistent else 1
        dom_ms = max_over_ranks(sum(t_step) / len(t_step))
        dom_bytes = 160.0 * V_local * steps_per_launch
        dom_share = sum(t_step) / step_ms
        dom_name = ("cg_step_kernel (%s; per CG step, in one pass: r -= a q ; x += a p ; p = r + b p ; q = D^dag D p ; "
                    "|r|^2, <p,q>, <r,q>, |q|^2 ; R r,q,p,x,Ux,Uy 96 + W r,p,q,x 64 = 160 B/site): the step's dominant kernel"
                    % ("persistent: ONE launch runs all %d CG steps of the solve, meeting at one grid-wide reduction per "
                       "step" % steps_per_launch if persistent else "one launch per CG step"))
        dom_key, dom_n = "cg_step_kernel", len(t_step)
        dom_extra = {"cg_steps_per_launch": steps_per_launch, "ms_per_cg_step": dom_ms / steps_per_launch}
    # the two-kernel loop (glb_cg_step_mode(0)): its kernels are timed the same way on one more solve
    ctx.cg_step_mode(False)
    solve_resident()
    ctx.prof_enable(True)
    t2 = time.perf_counter()
    info_2 = solve_resident()
    ctx.sync()
    t2 = time.perf_counter() - t2
    ctx.prof_enable(False)
    ctx.cg_step_mode(True)
    it_2 = info_2["iter"]
    t_fused = ctx.prof_read(1)[:max(it_2 - 1, 0)]
    t_upd = ctx.prof_read(3)[:it_2]
    fused_ms = max_over_ranks(sum(t_fused) / max(len(t_fused), 1))
    upd_ms = max_over_ranks(sum(t_upd) / max(len(t_upd), 1))
    fused_gbps = 96.0 * V_local / (fused_ms * 1e-3) / 1e9 if t_fused else 0.0   # R r,p,U 64 + W p,Ap 32
    upd_gbps = 96.0 * V_local / (upd_ms * 1e-3) / 1e9 if t_upd else 0.0         # R x,p,r,Ap 64 + W x,r 32
    if not single_kernel:   # GLB_CGSTEP=0 in the environment: the two-kernel loop IS the step
        dom_ms, dom_bytes, dom_share = fused_ms, 96.0 * V_local, sum(t_fused) / step_ms
        dom_name = "normal_kernel<fused> (p = r + beta p ; Ap = D^dag D p ; <p,Ap> in one pass, 96 B/site)"
        dom_key, dom_n = "normal_kernel_fused", len(t_fused)
        dom_extra = {}
    dom_gbps = dom_bytes / (dom_ms * 1e-3) / 1e9

    # ---- end to end: the reference-facing call with host (pinned) buffers, on every rank (slab form: the rank's
    # rows of the vectors, the global gauge field of which it reads its own rows)
    hx, hb = ctx.pinned(V_local), ctx.pinned(V_local)
    hb[:] = bp.download()
    ctx.cache_operators(True)   # gauge field stays resident between solves (it is the "model"); vectors travel
    desc = ctx._desc("STAG_NORMAL_U1", X, Y, mass=MASS, links=links_global)

    def solve_e2e():
        return ctx.host_solve("CG", desc, hx, hb, max_iter=5000, eps=TOL)
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
    e2e_value = solve_bytes(V_global, e2e_iters) * args.steps / e2e_s / 1e9
    e2e_x = np.array(hx, copy=True) if world == 1 else None
    ctx.cache_operators(False)
    for o in (opD, opDd):
        o.destroy()
    del out

    # ---- strong scaling: the same solve on a fixed global lattice (mt19937 stream on every rank), with its own
    # single-GPU time taken by rank 0 on a communicator-free context
    strong = None
    if not args.no_strong and args.Y == 0:
        strong = {}
        opN.destroy()
        del x, b, bp
        for Lg in args.strong_sizes:
            if Lg % world != 0 or Lg // world < 8:
                continue
            try:
                strong[str(Lg)] = strong_case(ctx, glb, torch, dist, stream, barrier, max_over_ranks, Lg, world, rank,
                                              local, peak, L, step_ms if world == 1 else None, iters)
            except Exception as e:  # noqa: BLE001
                strong[str(Lg)] = {"error": "%s: %s" % (type(e).__name__, e)}
        opN = None

    # ---- the other BASELINE configurations on one GPU (2: 256^2 CGNE + BiCGStab, 3: 4096^2 CG-M + GMRES(20))
    configs = None
    if world == 1 and not args.no_configs:
        try:
            configs = other_configs(ctx, glb, L, peak)
        except Exception as e:  # noqa: BLE001
            configs = [{"error": "%s: %s" % (type(e).__name__, e)}]

    # ---- lattice sweep (north_star: stencil GB/s and iterations/s from 256^2 up, as a fraction of the HBM roofline);
    # 4096^2 is the headline run, 8192^2 the strong block
    sweep = None
    if world == 1 and not args.no_sweep:
        try:
            sweep = lattice_sweep(ctx, glb, torch, stream, peak)
        except Exception as e:  # noqa: BLE001
            sweep = [{"error": "%s: %s" % (type(e).__name__, e)}]

    if rank != 0:
        return

    cpu = None
    if not args.no_cpu and world == 1 and Y == L:
        # bounded sample of the same workload on the host: minv_vector_cg(max_iter=3), same arrays; and the oracle's
        # verdict on the GPU solution (true residual recomputed on the CPU by the reference operator)
        orc = _oracle()
        Uc = links_global
        bc = np.array(hb, copy=True)
        oop = orc.op("STAG_NORMAL_U1", L, L, mass=MASS, links=Uc)
        t0 = time.perf_counter()
        _, cinfo = orc.solve("CG", oop, bc, max_iter=3, eps=TOL)
        ct = time.perf_counter() - t0
        oD = orc.op("STAG_U1", L, L, mass=MASS, links=Uc)
        t0 = time.perf_counter()
        oD.apply(bc)
        cat = time.perf_counter() - t0
        oracle_rel = float(np.linalg.norm(oop.apply(e2e_x) - bc) / np.linalg.norm(bc))
        cpu = {"value": solve_bytes(L * L, cinfo["iter"]) / ct / 1e9, "unit": "GB/s", "cores": 1,
               "kind": orc.kind, "host_cpus": os.cpu_count(),
               "sample": "minv_vector_cg(max_iter=3) on D^dag D, %dx%d, the arrays uploaded to the GPU; %.2f s; "
                         "one square_staggered_u1 apply %.3f s = %.2f GB/s" % (L, L, ct, cat, 64.0 * L * L / cat / 1e9),
               "oracle_true_rel_residual_of_gpu_solution": oracle_rel}

    gl = golden_large().get(str(L), {}) if world == 1 and Y == L else {}
    ref_iters = gl.get("CGNE", {}).get("iter")
    cfg = workload_config(X, Y, world)
    run = {
        "iterations": iters, "reference_iterations": ref_iters, "true_rel_residual": true_rel,
        "beta_prediction_max_rel_err": pred_err,
        "gauge": gauge_note,
        "comm": ("single GPU" if world == 1 else
                 ("NVLink peer memory: halo rows and rank sums written by the kernels themselves, one rank-wide reduction "
                  "per CG iteration" if ctx.p2p else "NCCL send/recv + allreduce on the compute stream")),
        "cg_loop": ("single-kernel iteration (cgstep.cu): 1 launch + 1 reduction, 160 B/site per iteration"
                    if single_kernel else "two-kernel loop (GLB_CGSTEP=0): 192 B/site per iteration"),
        "host_numa": numa_note,
    }
    if slab_parity is not None:
        run["slab_parity"] = slab_parity
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": ("strong" if args.Y > 0 else "weak"), "vs_baseline": None,
        "dtype": "complex<f64>", "data": "synthetic",
        "config": cfg, "run": run,
        "solve_time_s": step_ms * 1e-3, "iterations_per_s": iters * args.steps / (ms_total * 1e-3),
        "frac_of_hbm_peak": value / world / peak,
        "value_algorithmic_model": value_model,
        "clocks": sampler.summary(),
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom_name,
                     "achieved": dom_gbps, "peak": peak, "unit": "GB/s", "frac": dom_gbps / peak,
                     "traffic": ncu_traffic(dom_key, L), "ms_per_launch": dom_ms,
                     "algorithmic_bytes_per_launch": dom_bytes, "launches_timed": dom_n,
                     "share_of_step": dom_share, "peak_source": peak_src, "per_gpu": True, **dom_extra,
                     "how": "CUDA events around every launch inside one more solve right after the timed region "
                            "(slab runs: includes waiting for the slowest rank in the reduction)",
                     "other_kernels": [
                         {"kernel": "stag_kernel (staggered D apply alone, 64 B/site; the metric's 'Dirac apply GB/s')",
                          "achieved": apply_gbps, "frac": apply_gbps / peak, "ms_per_launch": apply_ms,
                          "traffic": ncu_traffic("stag_kernel", L), "how": "loop of %d applies" % args.apply_reps},
                         {"kernel": "normal1_kernel (D^dag D alone in one pass, one site per thread, 64 B/site)",
                          "achieved": normal_gbps, "frac": normal_gbps / peak, "ms_per_launch": normal_ms,
                          "traffic": ncu_traffic("normal_kernel", L), "how": "loop of %d applies" % args.apply_reps},
                         {"kernel": "two-kernel loop, normal_kernel<fused> (p = r + beta p ; Ap = D^dag D p ; <p,Ap>, "
                                    "96 B/site)", "achieved": fused_gbps, "frac": fused_gbps / peak,
                          "ms_per_launch": fused_ms, "traffic": ncu_traffic("normal_kernel_fused", L),
                          "launches_timed": len(t_fused)},
                         {"kernel": "two-kernel loop, cg_update_kernel (x += alpha p ; r -= alpha Ap ; |r|^2, 96 B/site)",
                          "achieved": upd_gbps, "frac": upd_gbps / peak, "ms_per_launch": upd_ms,
                          "traffic": ncu_traffic("cg_update_kernel", L), "launches_timed": len(t_upd)}],
                     "two_kernel_loop": {"solve_time_s": t2, "iterations": it_2,
                                         "note": "the same solve with glb_cg_step_mode(0): 192 B/site per iteration"}},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(2 * 16 * V_local),
                "d2h_bytes_per_step": int(16 * V_local), "s_per_step": e2e_s / args.steps, "iterations": e2e_iters,
                "note": "minv_vector_cg(host pointers) on every rank; host vectors travel every step (pinned); the gauge "
                        "field is uploaded once and stays resident; bytes per step are per GPU"},
        "wall_s_timed_region": t_wall,
    }
    if slab_parity is not None:
        line["slab_parity"] = slab_parity
    if strong is not None:
        line["strong"] = strong
    if configs is not None:
        line["configs"] = configs
    if sweep is not None:
        line["sweep"] = sweep
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def slab_self_check(ctx, glb, opD, opN, X, Y, y0, Yloc, b_local, links_local, rows, b_dev):
    """compare the 32 rows next to each edge of this rank's slab (D and D^dag D applied to b) with the CPU oracle
    applied to a 72-row band of the global lattice around that edge; returns "ok" or a description"""
    try:
        orc = _oracle()
        V_local = X * Yloc
        out = ctx.vector(V_local)
        opD.apply(out, b_dev)
        gD = out.download().reshape(Yloc, X)
        opN.apply(out, b_dev)
        gN = out.download().reshape(Yloc, X)
        H = 72
        if Yloc < H - 4:
            return "skipped (slab thinner than the band)"
        ll = links_local.reshape(len(rows), 2 * X)
        bl = b_local.reshape(Yloc, X)
        bad = []
        for edge in ("low", "high"):
            # band rows (global): 36 rows either side of the slab edge; this rank holds two ghost rows of links beyond
            # its slab and none of b, so the band's outside half of b comes from the row-seeded generator
            e = y0 if edge == "low" else y0 + Yloc
            band = [(e - H // 2 + i) % Y for i in range(H)]
            Ub = gauge_rows(X, band)
            bb = rhs_rows(X, band)
            want_D = orc.op("STAG_U1", X, H, mass=MASS, links=Ub).apply(bb).reshape(H, X)
            want_N = orc.op("STAG_NORMAL_U1", X, H, mass=MASS, links=Ub).apply(bb).reshape(H, X)
            for i in range(4, H - 4):
                yl = band[i] - y0   # local row, if ours
                if not (0 <= yl < Yloc):
                    continue
                # eta_y = (-1)^x does not depend on y, so a band starting on any row sees the operator of the lattice
                if not (np.array_equal(gD[yl], want_D[i]) and np.array_equal(gN[yl], want_N[i])):
                    bad.append((edge, int(band[i])))
        return "ok" if not bad else "rows differ: %s" % bad[:4]
    except Exception as e:  # noqa: BLE001
        return "self-check failed to run: %s: %s" % (type(e).__name__, e)


def strong_case(ctx, glb, torch, dist, stream, barrier, max_over_ranks, Lg, world, rank, local, peak, L_weak,
                weak_ms, weak_iters):
    """CGNE on the global Lg x Lg lattice over `world` GPUs and on one GPU (rank 0, its own context)"""
    X = Y = Lg
    if world == 1 and Lg == L_weak and weak_ms is not None:
        return {"lattice": [X, Y], "ms_per_solve": weak_ms, "ms_per_solve_1gpu": weak_ms, "efficiency": 1.0,
                "iterations": weak_iters, "note": "the headline run itself"}
    links, b_h = ctx.synthetic_inputs(X, Y, SEED, BETA)
    lg = links.reshape(Y, 2 * X)
    y0, Yloc = ctx.slab_bounds(Y)
    rows = slab_rows(y0, Yloc, Y)
    links_local = np.ascontiguousarray(lg[rows]).reshape(-1)
    V_local = X * Yloc
    opN = ctx.staggered_local(links_local, X, Y, MASS, glb.STAG_NORMAL)
    opDd = ctx.staggered_local(links_local, X, Y, MASS, glb.STAG_DAGGER)
    b = ctx.vector(V_local).upload(b_h.reshape(Y, X)[y0:y0 + Yloc].reshape(-1))
    bp = ctx.vector(V_local)
    opDd.apply(bp, b)
    x = ctx.vector(V_local)

    reps = 3
    for _ in range(2):
        x.zero()
        info = ctx.solve("CG", opN, x, bp, max_iter=5000, eps=TOL)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        x.zero()
        info = ctx.solve("CG", opN, x, bp, max_iter=5000, eps=TOL)
    e1.record(stream)
    barrier()
    ms_n = max_over_ranks(e0.elapsed_time(e1)) / reps
    # per-kernel time of the iteration on this run (includes waiting for the other ranks)
    ctx.prof_enable(True)
    x.zero()
    ip = ctx.solve("CG", opN, x, bp, max_iter=5000, eps=TOL)
    ctx.prof_enable(False)
    ts = ctx.prof_read(7)
    if len(ts) != 1:
        ts = ts[:ip["iter"] + 1]
    k_ms = max_over_ranks(sum(ts) / (ip["iter"] + 1))   # per CG step, whether one launch per step or one per solve
    true_rel = float(np.sqrt(info["resSq"]) / np.sqrt(ctx.norm2sq(bp)))
    for o in (opN, opDd):
        o.destroy()
    del x, b, bp
    out = {"lattice": [X, Y], "rows_per_gpu": Yloc, "ms_per_solve": ms_n, "iterations": info["iter"],
           "true_rel_residual": true_rel, "cg_step_kernel_ms": k_ms, "us_per_iteration": 1e3 * ms_n / max(info["iter"], 1)}
    if world == 1:
        out.update({"ms_per_solve_1gpu": ms_n, "efficiency": 1.0})
        ref = golden_large().get(str(Lg), {}).get("CGNE", {}).get("iter")
        if ref:
            out["reference_iterations"] = ref
        return out
    ms_1 = 0.0
    if rank == 0:
        c1 = glb.Context(device=local, use_default=False)
        s1 = torch.cuda.ExternalStream(c1.stream(), device=torch.device("cuda", local))
        o1 = c1.staggered(links, X, Y, MASS, glb.STAG_NORMAL)
        od = c1.staggered(links, X, Y, MASS, glb.STAG_DAGGER)
        b1 = c1.vector(X * Y).upload(b_h)
        bp1 = c1.vector(X * Y)
        od.apply(bp1, b1)
        x1 = c1.vector(X * Y)
        for _ in range(2):
            x1.zero()
            i1 = c1.solve("CG", o1, x1, bp1, max_iter=5000, eps=TOL)
        c1.sync()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(s1)
        for _ in range(reps):
            x1.zero()
            i1 = c1.solve("CG", o1, x1, bp1, max_iter=5000, eps=TOL)
        a1.record(s1)
        c1.sync()
        torch.cuda.synchronize()
        ms_1 = a0.elapsed_time(a1) / reps
        out["iterations_1gpu"] = i1["iter"]
        for o in (o1, od):
            o.destroy()
        del x1, b1, bp1
        c1.close()
    barrier()
    ms_1 = max_over_ranks(ms_1)
    out.update({"ms_per_solve_1gpu": ms_1, "efficiency": ms_1 / (world * ms_n),
                "limiting": "cg_step_kernel %.1f us per CG step at %d rows per GPU against %.1f us = (1-GPU solve / "
                            "iterations / N): launch + grid tail + one rank-wide reduction per iteration"
                            % (1e3 * k_ms, Yloc, 1e3 * ms_1 / max(info["iter"], 1) / world)})
    ref = golden_large().get(str(Lg), {}).get("CGNE", {}).get("iter")
    if ref:
        out["reference_iterations"] = ref
    return out


def lattice_sweep(ctx, glb, torch, stream, peak, sizes=(256, 512, 1024, 2048)):
    """staggered D apply and CGNE solve per lattice size on one GPU (mt19937 inputs): GB/s on moved bytes, fraction of
    the measured HBM peak, iterations/s.  Up to 1024^2 the working set fits the 126 MB L2: latency / launch bound."""
    res = []
    gold = golden_large()
    for Ls in sizes:
        V = Ls * Ls
        links, b_h = ctx.synthetic_inputs(Ls, Ls, SEED, BETA)
        D = ctx.staggered(links, Ls, Ls, MASS, 0)
        N = ctx.staggered(links, Ls, Ls, MASS, glb.STAG_NORMAL)
        Dd = ctx.staggered(links, Ls, Ls, MASS, glb.STAG_DAGGER)
        b = ctx.vector(V).upload(b_h)
        bp, x, out = ctx.vector(V), ctx.vector(V), ctx.vector(V)
        Dd.apply(bp, b)

        def events(fn, reps, warm):
            for _ in range(warm):
                fn()
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r = None
            for _ in range(reps):
                r = fn()
            e1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            return r, e0.elapsed_time(e1) / reps

        _, a_ms = events(lambda: D.apply(out, b), 200 if Ls <= 1024 else 50, 5)

        def go():
            x.zero()
            return ctx.solve("CG", N, x, bp, max_iter=5000, eps=TOL)
        info, s_ms = events(go, 5, 2)
        it = info["iter"]
        rec = {"lattice": [Ls, Ls], "apply_ms": a_ms, "apply_GBps": 64.0 * V / a_ms / 1e6,
               "apply_frac_of_hbm_peak": 64.0 * V / a_ms / 1e6 / peak,
               "cgne_ms_per_solve": s_ms, "cgne_iterations": it, "cgne_iterations_per_s": it / (s_ms * 1e-3),
               "cgne_GBps": solve_bytes(V, it) / (s_ms * 1e-3) / 1e9,
               "cgne_frac_of_hbm_peak": solve_bytes(V, it) / (s_ms * 1e-3) / 1e9 / peak}
        ref = gold.get(str(Ls), {}).get("CGNE", {}).get("iter")
        if ref:
            rec["reference_iterations"] = ref
        if V * 160 <= 126e6:
            rec["note"] = "working set fits the L2: latency / launch bound"
        res.append(rec)
        for o in (D, N, Dd):
            o.destroy()
        del b, bp, x, out
    return res


def other_configs(ctx, glb, L, peak):
    """BASELINE configs 2 and 3 on one GPU, mt19937 inputs; reference iteration counts from tests/golden/"""
    res = []
    gold = golden_large()
    try:
        small = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["synthetic_beta6_m0.1"]
    except Exception:
        small = {}

    def timed(fn):
        fn()
        ctx.sync()
        t0 = time.perf_counter()
        r = fn()
        ctx.sync()
        return r, time.perf_counter() - t0

    for Lc, which in ((256, "config 2"), (1024, "1024^2 (device-resident BiCGStab / CR loops)"), (L, "config 3")):
        V = Lc * Lc
        links, b_h = ctx.synthetic_inputs(Lc, Lc, SEED, BETA)
        D = ctx.staggered(links, Lc, Lc, MASS, 0)
        N = ctx.staggered(links, Lc, Lc, MASS, glb.STAG_NORMAL)
        Dd = ctx.staggered(links, Lc, Lc, MASS, glb.STAG_DAGGER)
        b = ctx.vector(V).upload(b_h)
        bp = ctx.vector(V)
        Dd.apply(bp, b)
        x = ctx.vector(V)
        g = gold.get(str(Lc), small.get(str(Lc), {}))

        def solve(solver, op, rhs, **kw):
            def go():
                x.zero()
                return ctx.solve(solver, op, x, rhs, max_iter=100000, **kw)
            return timed(go)

        if which != "config 3":
            # bytes the kernels move per iteration and site (csrc/krylov.cu): BiCGStab 48 + 64 + 112 + 48 + 80 = 352,
            # CR 96 + 80 + 96 = 272
            cases = [("CGNE (minv_vector_cg on D^dag D)", "CG", N, bp, dict(eps=1e-10), "CGNE", 160.0),
                     ("minv_vector_bicgstab on D", "BICGSTAB", D, b, dict(eps=1e-10), "BiCGStab", 352.0)]
            if Lc != 256:
                cases.append(("minv_vector_cr on D^dag D", "CR", N, bp, dict(eps=1e-10), "CR", 272.0))
            for name, solver, op, rhs, kw, key, bytes_it in cases:
                info, dt = solve(solver, op, rhs, **kw)
                rec = {"config": which, "L": Lc, "solver": name, "seconds": dt, "iterations": info["iter"],
                       "reference_iterations": g.get(key, {}).get("iter"), "success": info["success"],
                       "us_per_iteration": 1e6 * dt / max(info["iter"], 1),
                       "moved_GBps": bytes_it * V * info["iter"] / dt / 1e9}
                rec["frac_of_hbm_peak"] = rec["moved_GBps"] / peak
                if solver != "CG":
                    # the same solve through the host-scalar shell (scalars read back 3-4 times per iteration)
                    ctx.force_host_scalars(True)
                    try:
                        info_h, dt_h = solve(solver, op, rhs, **kw)
                    finally:
                        ctx.force_host_scalars(False)
                    rec["loop"] = ("device-resident (csrc/krylov.cu)" if ctx.krylov_supported(solver, op)
                                   else "host-scalar shell")
                    rec["host_scalar_shell"] = {"seconds": dt_h, "iterations": info_h["iter"],
                                                "us_per_iteration": 1e6 * dt_h / max(info_h["iter"], 1)}
                res.append(rec)
        else:
            shifts = [0.0, 0.01, 0.05, 0.25]
            xs = [ctx.vector(V) for _ in shifts]

            def go_m():
                for v in xs:
                    v.zero()
                return ctx.solve_cg_m(N, xs, bp, shifts, resid_freq_check=10, max_iter=100000, eps=1e-10)[0]
            info, dt = timed(go_m)
            res.append({"config": which, "L": Lc, "solver": "minv_vector_cg_m on D^dag D, shifts {0,.01,.05,.25}",
                        "seconds": dt, "iterations": info["iter"], "reference_iterations": g.get("CG-M", {}).get("iter"),
                        "success": info["success"], "us_per_iteration": 1e6 * dt / max(info["iter"], 1),
                        "moved_GBps_if_all_shifts_stayed_live": (64.0 + 16.0 + 48.0 * 4 + 16.0 + (32.0 * 4 + 16.0)) * V * info["iter"] / dt / 1e9,
                        "note": "upper bound on the traffic: converged shifts drop out of the update kernels (checked every 10 iterations)"})
            del xs
            info, dt = solve("GMRES_RESTART", D, b, eps=1e-8, restart_freq=20)
            res.append({"config": which, "L": Lc, "solver": "minv_vector_gmres_restart(20) on D, tol 1e-8", "seconds": dt,
                        "iterations": info["iter"], "reference_iterations": g.get("GMRES(20)", {}).get("iter"),
                        "success": info["success"], "us_per_iteration": 1e6 * dt / max(info["iter"], 1)})
        for o in (D, N, Dd):
            o.destroy()
        del x, b, bp
    return res


if __name__ == "__main__":
    main()
