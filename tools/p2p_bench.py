#!/usr/bin/env python
"""Latency of the kernel-side rank-wide sum (csrc/p2p.cuh) in isolation, under torchrun:
every rank busy-waits `busy` us, then sums 6 doubles over ranks; prints, per variant, the mean of each order
statistic over ranks of the time spent in the sum (see p2p_bench_kernel in csrc/comm.cu)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from __graft_entry__ import _load_pkg  # noqa: E402

glb = _load_pkg()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = glb.Context(device=local)
ctx.init_comm_from_torch()
fn = ctx.cu.glb_dbg_p2p_bench
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_float)]
iters = 300
pp = ctx.cu.glb_dbg_p2p_pingpong
pp.restype = C.c_int
pp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
for peer in range(1, 2):
    buf = (C.c_float * iters)()
    dist.barrier()
    pp(ctx.h, peer, iters, buf)
    if rank == 0:
        w = np.array(buf[:], dtype=np.float64)[20:]
        print("ping-pong 0 <-> %d: round trip mean %.2f us, median %.2f, p90 %.2f, max %.2f" % (
            peer, w.mean(), np.median(w), np.percentile(w, 90), w.max()), flush=True)
# the library baseline: NCCL all-reduce of 6 doubles
t = torch.zeros(6, dtype=torch.float64, device="cuda")
for _ in range(20):
    dist.all_reduce(t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
e0.record()
for _ in range(200):
    dist.all_reduce(t)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("NCCL all_reduce of 6 doubles, back to back: %.2f us each" % (1e3 * e0.elapsed_time(e1) / 200), flush=True)
for busy in [float(b) for b in (sys.argv[1:] or ["80"])]:
    for variant in (0, 70, 80, 270, 280, 470, 480):
        buf = (C.c_float * iters)()
        dist.barrier()
        rc = fn(ctx.h, variant, iters, busy, buf)
        w = np.array(buf[:], dtype=np.float64)[20:]
        allw = [None] * world
        dist.all_gather_object(allw, w)
        if rank == 0:
            na = variant // 100 if variant >= 100 else world
            A = np.array(allw)[:na]
            W = np.sort(A, axis=0)
            print("busy %5.0f us variant %3d rc %d: mean wait %.2f us; order statistics over ranks %s" % (
                busy, variant, rc, W.mean(), np.round(W.mean(axis=1), 1)), flush=True)
            if variant in (0, 70, 80):
                print("   raw waits, ranks x 14 consecutive iterations:\n" + np.array2string(A[:, 100:114], precision=1, max_line_width=200), flush=True)
dist.barrier()
dist.destroy_process_group()
