"""Host-side (numpy) set-up of a two-level adaptive-multigrid hierarchy for the 2-D U(1) staggered operator:
what the reference does in null_generate_random_smooth / block_orthonormalize /
generate_coarse_from_fine_stencil (multigrid/aa_mg/null_gen.cpp:193, mg_complex.cpp:259, :827), vectorised so
that it is usable at 2048^2.  TEST / MEASUREMENT TOOLING: an independent restatement that both the reference's set-up
(tests/test_mg_setup_cpu.py) and the device set-up of the product (glbx_mg_setup, tests/test_mg_setup_gpu.py)
are checked against, and the host-side timing the device set-up is compared with (tools/bench_mg.py --numpy-setup)."""
import numpy as np


def staggered_stencil(U, X, Y, mass):
    """nc = 1 stencil of the staggered operator as get_square_staggered_u1_stencil builds it
    (operators_stencil.cpp:14-63): hopping[+x] = -U_x/2, [+y] = -eta U_y/2, [-x] = +conj U_x(x-1)/2,
    [-y] = +eta conj U_y(y-1)/2, shift = mass, clover = 0.  Returns (clover, hopping[4*V], shift)."""
    Ul = np.asarray(U).reshape(Y, X, 2)
    eta = (1.0 - 2.0 * (np.arange(X) % 2))[None, :]
    hop = np.empty((4, Y, X), dtype=np.complex128)
    hop[0] = -0.5 * Ul[:, :, 0]
    hop[1] = -0.5 * eta * Ul[:, :, 1]
    hop[2] = 0.5 * np.conj(np.roll(Ul[:, :, 0], 1, axis=1))
    hop[3] = 0.5 * eta * np.conj(np.roll(Ul[:, :, 1], 1, axis=0))
    return np.zeros(X * Y, dtype=np.complex128), hop.reshape(-1), complex(mass)


def split_even_odd(vecs, X, Y):
    """BLOCK_EO partition (input_params.cpp:736-741): every vector becomes an even-site and an odd-site vector;
    all even parts first, then all odd parts (null_gen.cpp layout j + k*n_null_vectors)."""
    idx = np.arange(X * Y)
    even = ((idx % X + idx // X) % 2) == 0
    return [np.where(even, v, 0) for v in vecs] + [np.where(~even, v, 0) for v in vecs]


def _to_blocks(v, X, Y, bx, by):
    nv = v.shape[0]
    return v.reshape(nv, Y // by, by, X // bx, bx).transpose(1, 3, 0, 2, 4).reshape(Y // by, X // bx, nv, by * bx)


def _from_blocks(B, X, Y, bx, by):
    Yc, Xc, nv, _ = B.shape
    return B.reshape(Yc, Xc, nv, by, bx).transpose(2, 0, 3, 1, 4).reshape(nv, Y * X)


def block_orthonormalize(vecs, X, Y, bx, by):
    """mg_complex.cpp:259-370: inside every block, Gram-Schmidt the vectors in order (each vector is normalised
    just before its successor is orthogonalised against it), then block_normalize."""
    B = _to_blocks(np.array(vecs, dtype=np.complex128), X, Y, bx, by).copy()
    nv = B.shape[2]
    for v in range(1, nv):
        nrm = np.sqrt(np.sum(np.abs(B[:, :, v - 1]) ** 2, axis=-1))
        B[:, :, v - 1] /= nrm[..., None]
        for m in range(v):
            d = np.sum(np.conj(B[:, :, m]) * B[:, :, v], axis=-1)
            B[:, :, v] -= d[..., None] * B[:, :, m]
    nrm = np.sqrt(np.sum(np.abs(B) ** 2, axis=-1))
    B /= nrm[..., None]
    return list(_from_blocks(B, X, Y, bx, by))


def coarse_stencil(null, hop, shift, X, Y, bx, by):
    """Galerkin coarse operator P^dag A P of an nc = 1 five-point fine stencil (clover 0, the given shift), as
    generate_coarse_from_fine_stencil(..., ignore_shifts=false) produces it (mg_complex.cpp:827-1026): the part
    of every fine hop that stays inside a block goes to the coarse clover, the part that leaves it to the coarse
    hopping term of that direction; the fine shift ends up in the coarse clover.
    Returns (clover[Vc*nc*nc], hopping[4*Vc*nc*nc]) in the stencil_2d layout (row-major nc x nc per site)."""
    nv = len(null)
    n = np.array(null, dtype=np.complex128).reshape(nv, Y, X)
    h = np.asarray(hop).reshape(4, Y, X)
    Yc, Xc = Y // by, X // bx
    xs, ys = np.arange(X)[None, :], np.arange(Y)[:, None]
    # direction d: value at the neighbour (np.roll) and the mask of sites whose neighbour is in the same block
    nb = [lambda a: np.roll(a, -1, axis=-1), lambda a: np.roll(a, -1, axis=-2),
          lambda a: np.roll(a, 1, axis=-1), lambda a: np.roll(a, 1, axis=-2)]
    inside = [np.broadcast_to((xs % bx) != bx - 1, (Y, X)), np.broadcast_to((ys % by) != by - 1, (Y, X)),
              np.broadcast_to((xs % bx) != 0, (Y, X)), np.broadcast_to((ys % by) != 0, (Y, X))]

    def blocksum(a):
        return a.reshape(Yc, by, Xc, bx).sum(axis=(1, 3))

    clover = np.zeros((Yc, Xc, nv, nv), dtype=np.complex128)
    hopc = np.zeros((4, Yc, Xc, nv, nv), dtype=np.complex128)
    cn = np.conj(n)
    for j in range(nv):
        diag = shift * n[j]
        for i in range(nv):
            clover[:, :, i, j] += blocksum(cn[i] * diag)
        for d in range(4):
            t = h[d] * nb[d](n[j])
            t_in = np.where(inside[d], t, 0)
            t_out = t - t_in
            for i in range(nv):
                clover[:, :, i, j] += blocksum(cn[i] * t_in)
                hopc[d, :, :, i, j] = blocksum(cn[i] * t_out)
    return clover.reshape(-1), hopc.reshape(-1)


def null_vectors_device(ctx, op, V, nvec, seed=1337, tol=5e-5, max_iter=500, solver="BICGSTAB"):
    """null_generate_random_smooth (null_gen.cpp:193-300) with the device solver: for a gaussian x0 solve
    A x = -A x0 from a zero guess (BiCGStab, tol 5e-5, at most 500 iterations: input_params.cpp:683-705) and keep
    x + x0."""
    rg = np.random.default_rng(seed)
    out = []
    x0d, rhs, x = ctx.vector(V), ctx.vector(V), ctx.vector(V)
    for _ in range(nvec):
        x0 = rg.standard_normal(V) + 1j * rg.standard_normal(V)
        x0d.upload(x0)
        op.apply(rhs, x0d)
        rhs.upload(-rhs.download())
        x.zero()
        ctx.solve(solver, op, x, rhs, max_iter=max_iter, eps=tol)
        out.append(x.download() + x0)
    return out
