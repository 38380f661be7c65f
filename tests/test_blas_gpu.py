"""GPU parity of the BLAS-1 layer (generic_vector.h) through the C ABI.

Element-wise updates must be bit-identical to the reference expressions evaluated on the host in
the same order (numpy complex arithmetic = (ac-bd, ad+bc), no FMA).  Reductions differ from the
serial CPU sum only by summation order: relative 1e-13, and they are reproducible run to run.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 31, 1000, 4096, 65537, 1 << 20]


def _vecs(ctx, rg, n, dtype, k):
    hs = []
    for _ in range(k):
        h = rg.standard_normal(n)
        if dtype == np.complex128:
            h = h + 1j * rg.standard_normal(n)
        hs.append(np.ascontiguousarray(h.astype(dtype)))
    return hs, [ctx.vector(n, dtype).upload(h) for h in hs]


def cm(a, x):
    """a*x with every product and sum rounded separately (numpy's SIMD complex multiply may fuse)"""
    if not np.iscomplexobj(x):
        return a * x
    a = complex(a)
    out = np.empty_like(x)
    out.real = a.real * x.real - a.imag * x.imag
    out.imag = a.real * x.imag + a.imag * x.real
    return out


def _c(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_reductions(ctx, glb, orc, n, dtype):
    rg = np.random.default_rng(n)
    (x, y), (dx, dy) = _vecs(ctx, rg, n, dtype, 2)
    want = orc.dot(x, y)
    got = ctx.dot(dx, dy)
    assert abs(got - want) <= 1e-13 * max(abs(want), np.sqrt(orc.norm2sq(x) * orc.norm2sq(y)))
    assert abs(ctx.norm2sq(dx) - orc.norm2sq(x)) <= 1e-13 * orc.norm2sq(x)
    assert abs(ctx.diffnorm2sq(dx, dy) - orc.diffnorm2sq(x, y)) <= 1e-13 * orc.diffnorm2sq(x, y)
    assert ctx.dot(dx, dy) == got  # reproducible
    d = (C.c_double * 3)()
    assert ctx.cu.glb_dot_norm(ctx.h, dx.dt, n, dx.ptr, dy.ptr, d) == 0
    assert abs(complex(d[0], d[1]) - want) <= 1e-13 * np.sqrt(orc.norm2sq(x) * orc.norm2sq(y))
    assert abs(d[2] - orc.norm2sq(x)) <= 1e-13 * orc.norm2sq(x)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_elementwise_bit_exact(ctx, glb, n, dtype):
    rg = np.random.default_rng(n + 1)
    hs, ds = _vecs(ctx, rg, n, dtype, 6)
    cu, h = ctx.cu, ctx.h
    dt = ds[0].dt
    a = (0.37 - 1.3j) if dtype == np.complex128 else 0.37
    b = (-0.9 + 0.2j) if dtype == np.complex128 else -0.9
    x, y, z, p, q, r = hs
    # axpy: y = y + a*x
    assert cu.glb_axpy(h, dt, n, _c(a), ds[0].ptr, ds[1].ptr) == 0
    y = y + cm(a, x)
    assert np.array_equal(ds[1].download(), y)
    # xpay: y = x + a*y
    assert cu.glb_xpay(h, dt, n, ds[0].ptr, _c(a), ds[1].ptr) == 0
    y = x + cm(a, y)
    assert np.array_equal(ds[1].download(), y)
    # axpyz: z = y + a*x ; sub ; add ; rdiv
    assert cu.glb_axpyz(h, dt, n, _c(b), ds[0].ptr, ds[1].ptr, ds[2].ptr) == 0
    z = y + cm(b, x)
    assert np.array_equal(ds[2].download(), z)
    assert cu.glb_sub(h, dt, n, ds[0].ptr, ds[1].ptr, ds[3].ptr) == 0
    p = x - y
    assert np.array_equal(ds[3].download(), p)
    assert cu.glb_add(h, dt, n, ds[0].ptr, ds[1].ptr, ds[3].ptr) == 0
    p = x + y
    assert np.array_equal(ds[3].download(), p)
    assert cu.glb_rdiv(h, dt, n, ds[3].ptr, 1.7, ds[3].ptr) == 0
    p = (p / 1.7) if dtype == np.float64 else (p.real / 1.7 + 1j * (p.imag / 1.7))  # complex/real is component-wise
    assert np.array_equal(ds[3].download(), p)
    # CG update: x = x + a p ; r = r + b q ; |r|^2
    out = C.c_double()
    assert cu.glb_update_xr_norm(h, dt, n, _c(a), ds[3].ptr, ds[0].ptr, _c(b), ds[4].ptr, ds[5].ptr, C.byref(out)) == 0
    x = x + cm(a, p)
    r = r + cm(b, q)
    assert np.array_equal(ds[0].download(), x) and np.array_equal(ds[5].download(), r)
    assert abs(out.value - np.vdot(r, r).real) <= 1e-13 * np.vdot(r, r).real
    # CR update: p = r + beta p ; Ap = Ar + beta Ap ; |Ap|^2  (vectors: r=ds5, Ar=ds4, p=ds3, Ap=ds2)
    assert cu.glb_update_p_ap_norm(h, dt, n, ds[5].ptr, ds[4].ptr, _c(a), ds[3].ptr, ds[2].ptr, C.byref(out)) == 0
    p = r + cm(a, p)
    z = q + cm(a, z)
    assert np.array_equal(ds[3].download(), p) and np.array_equal(ds[2].download(), z)
    assert abs(out.value - np.vdot(z, z).real) <= 1e-13 * np.vdot(z, z).real
    # BiCGStab p update: p = r + beta*(p - omega*Ap)
    assert cu.glb_bicgstab_pupdate(h, dt, n, ds[5].ptr, _c(a), _c(b), ds[2].ptr, ds[3].ptr) == 0
    p = r + cm(a, p - cm(b, z))
    assert np.array_equal(ds[3].download(), p)
    # BiCGStab update: x = x + alpha p + omega s ; r = s - omega As ; |r|^2, <r0,r>
    o3 = (C.c_double * 3)()
    s_, As_, r0_ = y, z, q
    assert cu.glb_bicgstab_update(h, dt, n, _c(a), ds[3].ptr, _c(b), ds[1].ptr, ds[2].ptr, ds[4].ptr, ds[0].ptr,
                                  ds[5].ptr, o3) == 0
    x = x + cm(a, p) + cm(b, s_)
    r = s_ - cm(b, As_)
    assert np.array_equal(ds[0].download(), x) and np.array_equal(ds[5].download(), r)
    assert abs(o3[0] - np.vdot(r, r).real) <= 1e-13 * np.vdot(r, r).real
    assert abs(complex(o3[1], o3[2]) - np.vdot(r0_, r)) <= 1e-13 * np.sqrt(np.vdot(r, r).real * np.vdot(r0_, r0_).real)


@pytest.mark.parametrize("k,n", [(1, 5000), (3, 5000), (8, 5000), (9, 5000), (16, 5000), (17, 5000), (40, 5000),
                                 (5, 300001), (16, 300001), (7, 128)])
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_multi_vector_ops(ctx, glb, k, n, dtype):
    """n = 5000: 39 full 128-element chunks + a ragged tail; 300001: several chunks per block; k <= 8 / > 8: one / two
    vectors per warp of multi_dot_kernel; k > 16: several launches"""
    rg = np.random.default_rng(k)
    hs, ds = _vecs(ctx, rg, n, dtype, k + 2)
    cu, h = ctx.cu, ctx.h
    dt = ds[0].dt
    X = (C.c_void_p * k)(*[d.ptr for d in ds[:k]])
    y_h, y_d = hs[k], ds[k]
    out = (C.c_double * (2 * k))()
    assert cu.glb_multi_dot(h, dt, n, k, X, y_d.ptr, out) == 0
    for j in range(k):
        want = np.vdot(hs[j], y_h)
        assert abs(complex(out[2 * j], out[2 * j + 1]) - want) <= 1e-12 * np.linalg.norm(hs[j]) * np.linalg.norm(y_h)
    out2 = (C.c_double * (2 * k))()
    assert cu.glb_multi_dot(h, dt, n, k, X, y_d.ptr, out2) == 0
    assert list(out2) == list(out)  # fixed summation order: run-to-run reproducible
    coefs = rg.standard_normal(2 * k)
    if dtype == np.float64:
        coefs[1::2] = 0.0
    cz = coefs[0::2] + 1j * coefs[1::2]
    carr = (C.c_double * (2 * k))(*coefs)
    # out = init + sum_j c_j X_j, accumulated in order
    assert cu.glb_lincomb(h, dt, n, k, carr, X, y_d.ptr, ds[k + 1].ptr) == 0
    want = y_h.copy()
    for j in range(k):
        want = want + cm(cz[j] if dtype == np.complex128 else coefs[2 * j], hs[j])
    assert np.array_equal(ds[k + 1].download(), want)
    # in place (init == out), and without init
    assert cu.glb_lincomb(h, dt, n, k, carr, X, y_d.ptr, y_d.ptr) == 0
    assert np.array_equal(y_d.download(), want)
    assert cu.glb_lincomb(h, dt, n, k, carr, X, None, ds[k + 1].ptr) == 0
    want0 = np.zeros(n, dtype=dtype)
    for j in range(k):
        want0 = want0 + cm(cz[j] if dtype == np.complex128 else coefs[2 * j], hs[j])
    assert np.array_equal(ds[k + 1].download(), want0)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_multishift_updates(ctx, glb, dtype):
    n, ns = 3001, 5
    rg = np.random.default_rng(9)
    hs, ds = _vecs(ctx, rg, n, dtype, 2 * ns + 1)
    cu, h = ctx.cu, ctx.h
    dt = ds[0].dt
    ps, xs, r = hs[:ns], hs[ns:2 * ns], hs[2 * ns]
    P = (C.c_void_p * ns)(*[d.ptr for d in ds[:ns]])
    Xp = (C.c_void_p * ns)(*[d.ptr for d in ds[ns:2 * ns]])
    c0 = rg.standard_normal(2 * ns)
    c1 = rg.standard_normal(2 * ns)
    if dtype == np.float64:
        c0[1::2] = 0
        c1[1::2] = 0
    z0, z1 = c0[0::2] + 1j * c0[1::2], c1[0::2] + 1j * c1[1::2]
    if dtype == np.float64:
        z0, z1 = z0.real, z1.real
    assert cu.glb_cgm_update_x(h, dt, n, ns, (C.c_double * (2 * ns))(*c0), P, Xp) == 0
    for s in range(ns):
        assert np.array_equal(ds[ns + s].download(), xs[s] - cm(z0[s], ps[s]))
    assert cu.glb_cgm_update_p(h, dt, n, ns, (C.c_double * (2 * ns))(*c0), (C.c_double * (2 * ns))(*c1),
                               ds[2 * ns].ptr, P) == 0
    for s in range(ns):
        assert np.array_equal(ds[s].download(), cm(z0[s], r) + cm(z1[s], ps[s]))
    out = C.c_double()
    a = -0.3 + 0.8j if dtype == np.complex128 else -0.3
    assert cu.glb_axpy_norm(h, dt, n, (C.c_double * 2)(complex(a).real, complex(a).imag), ds[0].ptr, ds[2 * ns].ptr,
                            C.byref(out)) == 0
    p0 = cm(z0[0], r) + cm(z1[0], ps[0])
    r2 = r + cm(a, p0)
    assert np.array_equal(ds[2 * ns].download(), r2)
    assert abs(out.value - np.vdot(r2, r2).real) <= 1e-13 * np.vdot(r2, r2).real
