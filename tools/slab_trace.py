#!/usr/bin/env python
"""Where does a CG step go on slabs?  Run under torchrun; every rank solves its slab of an L x Y lattice with the
single-kernel CG and dumps, for every step, when its local sums were done and when the sum over ranks was done
(GLB_CGSTEP_TRACE; %globaltimer).  Prints per rank: mean step time, mean wait inside the rank-wide reduction.

    python -m torch.distributed.run --nproc-per-node N ... tools/slab_trace.py L Y outdir
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
L, Y, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
os.makedirs(out, exist_ok=True)
path = os.path.join(out, "trace_%dx%d_n%d_rank%d.txt" % (L, Y, world, rank))
os.environ["GLB_CGSTEP_TRACE"] = path
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from __graft_entry__ import _load_pkg  # noqa: E402

glb = _load_pkg()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = glb.Context(device=local)
ctx.init_comm_from_torch()
y0, Yloc = ctx.slab_bounds(Y)
rows = bench.slab_rows(y0, Yloc, Y)
N = ctx.staggered_local(bench.gauge_rows(L, rows), L, Y, 0.1, glb.STAG_NORMAL)
b = ctx.vector(L * Yloc).upload(bench.rhs_rows(L, rows[2:-2]))
x = ctx.vector(L * Yloc).zero()
info = ctx.solve("CG", N, x, b, max_iter=5000, eps=1e-10)      # the traced solve is the first one
ctx.sync()
d = np.loadtxt(path + ".steps")
step = np.diff(d[:, 2]) / 1e3
wait = (d[:, 2] - d[:, 1]) / 1e3
line = "rank %d: %d steps, step %.1f us (min %.1f max %.1f), wait in the rank-wide sum %.2f us mean / %.2f median / %.2f max" % (
    rank, len(d), step[3:].mean(), step[3:].min(), step[3:].max(), wait[3:].mean(), np.median(wait[3:]), wait[3:].max())
lines = [None] * world
dist.all_gather_object(lines, line)
if rank == 0:
    print("\n".join(lines), flush=True)
dist.barrier()
dist.destroy_process_group()
