"""N>1 host logic on CPU (gloo, world_size 2): the y-slab decomposition conventions the CUDA path
implements (csrc/ops.cu slab_of / upload_links, csrc/comm.cu halo_exchange):

  * rank g owns rows [g*Y/G, (g+1)*Y/G); ring of slabs (periodic lattice);
  * per apply each rank needs the psi row below and above its slab and U_y of the row below;
  * inner products are sums of per-slab partial sums;
  * bench.py's row-seeded input generator gives every decomposition the same global field.

Each rank builds its slab + ghosts from messages exchanged over torch.distributed (gloo), applies
the ORACLE operator to the extended slab and the interior must equal the global apply bit for bit.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_pkg


def _worker(rank, world, port, X, Y, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle_py
    import bench
    glb = load_pkg()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = oracle_py.load("port")
        y0, Yloc = glb.slab_bounds(Y, rank, world)
        rows = [(y0 - 1 + Y) % Y] + list(range(y0, y0 + Yloc))
        links = bench.gauge_rows(X, rows).reshape(Yloc + 1, 2 * X)      # row 0 = the row below the slab
        psi = bench.rhs_rows(X, rows[1:]).reshape(Yloc, X)
        # halo exchange: my lowest row -> down's ghost_hi ; my highest row -> up's ghost_lo
        up, down = (rank + 1) % world, (rank - 1 + world) % world
        lo = torch.zeros(2 * X, dtype=torch.float64)
        hi = torch.zeros(2 * X, dtype=torch.float64)
        send_lo = torch.from_numpy(psi[0].view(np.float64).copy())
        send_hi = torch.from_numpy(psi[-1].view(np.float64).copy())
        reqs = [dist.isend(send_lo, down, tag=1), dist.isend(send_hi, up, tag=2),
                dist.irecv(hi, up, tag=1), dist.irecv(lo, down, tag=2)]
        for r in reqs:
            r.wait()
        ghost_lo, ghost_hi = lo.numpy().view(np.complex128), hi.numpy().view(np.complex128)
        # extended slab: [ghost_lo, slab rows, ghost_hi]; links: [row below, slab rows, (unused) copy]
        ext = np.concatenate([ghost_lo, psi.reshape(-1), ghost_hi])
        ext_links = np.concatenate([links.reshape(-1), links[-1]])
        out_ext = orc.op("STAG_U1", X, Yloc + 2, mass=0.1, links=ext_links).apply(ext)
        mine = out_ext.reshape(Yloc + 2, X)[1:-1].reshape(-1)
        # global reference on every rank (same generator, all rows)
        allrows = list(range(Y))
        U = bench.gauge_rows(X, allrows)
        v = bench.rhs_rows(X, allrows)
        want = orc.op("STAG_U1", X, Y, mass=0.1, links=U).apply(v).reshape(Y, X)[y0:y0 + Yloc].reshape(-1)
        ok_apply = bool(np.array_equal(mine, want))
        # inner product = allreduce of slab partial sums
        part = torch.tensor([np.vdot(psi.reshape(-1), mine).real, np.vdot(psi.reshape(-1), mine).imag],
                            dtype=torch.float64)
        dist.all_reduce(part)
        full = np.vdot(v, orc.op("STAG_U1", X, Y, mass=0.1, links=U).apply(v))
        ok_dot = abs(complex(part[0].item(), part[1].item()) - full) <= 1e-12 * abs(full)
        q.put((rank, ok_apply, ok_dot, y0, Yloc))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("X,Y", [(8, 8), (6, 7)])
def test_slab_decomposition_world2(X, Y):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + Y
    procs = [ctx.Process(target=_worker, args=(r, 2, port, X, Y, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res), res
    assert all(r[2] for r in res), res
    assert res[0][3] == 0 and res[0][3] + res[0][4] == res[1][3] and res[1][3] + res[1][4] == Y


def test_slab_bounds_cover_lattice():
    glb = load_pkg()
    for Y in (1, 7, 64, 4096, 4097):
        for G in (1, 2, 4, 8):
            if Y < G:
                continue
            spans = [glb.slab_bounds(Y, g, G) for g in range(G)]
            assert spans[0][0] == 0 and sum(s[1] for s in spans) == Y
            for a, b in zip(spans, spans[1:]):
                assert a[0] + a[1] == b[0]
            assert max(s[1] for s in spans) - min(s[1] for s in spans) <= 1


def test_row_seeded_inputs_are_partition_independent():
    sys.path.insert(0, ROOT)
    import bench
    X, Y = 8, 12
    full = bench.gauge_rows(X, list(range(Y))).reshape(Y, 2 * X)
    for G in (2, 3, 4):
        glb = load_pkg()
        for g in range(G):
            y0, yl = glb.slab_bounds(Y, g, G)
            part = bench.gauge_rows(X, list(range(y0, y0 + yl))).reshape(yl, 2 * X)
            assert np.array_equal(part, full[y0:y0 + yl])
    assert np.allclose(np.abs(full), 1.0)
