#!/bin/bash
TAG=${1:-t05}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in 1 2 3; do for ls in 1 0; do
GLB_STAG_LOCKSTEP=$ls GLB_NORMAL_STAGES=4 python - <<PY | tee -a $OUT/summary.txt
import os, sys, time, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import _load_pkg
import torch, bench
glb = _load_pkg(); ctx = glb.Context(device=0)
L = 4096
rows = [L-2, L-1] + list(range(L)) + [0, 1]
U = bench.gauge_rows(L, rows); b = bench.rhs_rows(L, rows[2:-2])
stream = torch.cuda.ExternalStream(ctx.stream())
res = []
for flags, name in ((0, 'D'), (1, 'Ddag'), (4, 'DdagD')):
    op = ctx.staggered_local(U, L, L, 0.1, flags)
    x = ctx.vector(L*L).upload(b); y = ctx.vector(L*L)
    for _ in range(5): op.apply(y, x)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(50): op.apply(y, x)
    e1.record(stream); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/50
    res.append('%s %.4f ms %.0f GB/s' % (name, ms, 64*L*L/ms/1e6))
    op.destroy()
print('lockstep=%s rep %s :' % (os.environ['GLB_STAG_LOCKSTEP'], '$rep'), ' | '.join(res))
PY
done; done
