"""ctypes loader for the CPU checkers in oracle/ (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (generic-linalg_b200/) never does.

    orc = load("ref")    # oracle/_ref/libref_oracle.so  -- unmodified reference sources
    orc = load("port")   # oracle/libport_oracle.so      -- our restatement
    orc = load("best")   # ref if it was built, else port
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# must mirror oracle/oracle_api.h
OP = dict(LAPLACE_REAL=0, LAPLACE_IMAG=1, LAPLACE_NC=2, LAPLACE_U1=3, STAG_FREE=4, STAG_U1=5,
          STAG_GAMMA5_U1=6, STAG_DAGGER_U1=7, STAG_NORMAL_U1=8, GAMMA5=9, STENCIL=10,
          STENCIL_FROM_STAG=11, STAG_GAMMA5_FREE=12, LAPLACE_REAL_NC=13, STAG_FREE_REAL=14,
          STAG_DEO_U1=15, STAG_DOE_U1=16, STAG_M2MDEODOE_U1=17, SYMMSHIFT_X=18, SYMMSHIFT_Y=19,
          STAG_2LINK_U1=20, STAG_INDEX=21)
SOLVER = dict(CG=0, CG_RESTART=1, CR=2, CR_RESTART=3, GCR=4, GCR_RESTART=5, BICGSTAB=6,
              BICGSTAB_RESTART=7, BICGSTAB_L=8, BICGSTAB_L_RESTART=9, GMRES=10, GMRES_RESTART=11)


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("X", C.c_int), ("Y", C.c_int), ("Nc", C.c_int),
                ("mass", C.c_double), ("links", C.c_void_p), ("clover", C.c_void_p),
                ("hopping", C.c_void_p), ("two_link", C.c_void_p), ("has_two", C.c_int),
                ("shift", C.c_double * 2), ("eo_shift", C.c_double * 2), ("dof_shift", C.c_double * 2),
                ("view", C.c_int), ("wilson_coeff", C.c_double)]


class Result(C.Structure):
    _fields_ = [("resSq", C.c_double), ("iter", C.c_int), ("success", C.c_int), ("ops_count", C.c_int),
                ("n_rhs", C.c_int), ("resSqmrhs", C.c_double * 32), ("name", C.c_char * 64)]

    def as_dict(self):
        d = dict(resSq=self.resSq, iter=self.iter, success=bool(self.success), ops_count=self.ops_count,
                 name=self.name.decode())
        if self.n_rhs > 0:
            d["resSqmrhs"] = [self.resSqmrhs[i] for i in range(self.n_rhs)]
        return d


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


VIEW = dict(NONE=0, M2MDEODOE=1, M2MDTBDBT=2, NORMAL_EO=3, NORMAL_TB=4, DAGGER_EO=5, DAGGER_TB=6)


class Operator:
    def __init__(self, orc, kind, X, Y, mass=0.0, Nc=1, links=None, clover=None, hopping=None, two_link=None,
                 shift=0j, eo_shift=0j, dof_shift=0j, view=0, wilson_coeff=0.0):
        """view (stencil kinds, reference library only): VIEW name or number of the composite operator built on the
        stencil -- M2MDEODOE, M2MDTBDBT, NORMAL_EO, NORMAL_TB, DAGGER_EO, DAGGER_TB"""
        self.orc = orc
        self._keep = (links, clover, hopping, two_link)
        d = OpDesc()
        d.kind = OP[kind] if isinstance(kind, str) else kind
        d.X, d.Y, d.Nc, d.mass = X, Y, Nc, mass
        d.links, d.clover, d.hopping, d.two_link = _ptr(links), _ptr(clover), _ptr(hopping), _ptr(two_link)
        d.has_two = 1 if two_link is not None else 0
        for name, v in (("shift", shift), ("eo_shift", eo_shift), ("dof_shift", dof_shift)):
            getattr(d, name)[0] = complex(v).real
            getattr(d, name)[1] = complex(v).imag
        d.view = VIEW[view] if isinstance(view, str) else view
        d.wilson_coeff = wilson_coeff
        self.desc = d
        self.h = orc._f("op_prepare")(C.byref(d))
        if not self.h:
            raise RuntimeError("this oracle library cannot build the operator (composite stencil views need `ref`)")
        self.is_complex = bool(orc._f("op_is_complex")(self.h))
        self.size = orc._f("op_size")(self.h)
        self.dtype = np.complex128 if self.is_complex else np.float64

    def apply(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=self.dtype)
        out = np.empty_like(rhs)
        self.orc._f("op_apply")(self.h, _ptr(out), _ptr(rhs))
        return out

    PART = dict(EO=1, OE=2, TB=3, BT=4)

    def apply_part(self, part, rhs):
        """apply_stencil_2d_{eo,oe,tb,bt} (coarse_stencil.cpp:395-1512) on a STENCIL operator"""
        rhs = np.ascontiguousarray(rhs, dtype=np.complex128)
        out = np.empty_like(rhs)
        self.orc._f("stencil_apply_part")(self.h, self.PART[part], _ptr(out), _ptr(rhs))
        return out

    def eoprec_prepare(self, rhs_orig):
        """operators.cpp:528 square_staggered_eoprec_prepare with this operator's links and mass"""
        rhs_orig = np.ascontiguousarray(rhs_orig, dtype=np.complex128)
        out = np.empty_like(rhs_orig)
        self.orc._f("eoprec_prepare")(self.h, _ptr(out), _ptr(rhs_orig))
        return out

    def eoprec_reconstruct(self, lhs_e, rhs_o):
        """operators.cpp:574 square_staggered_eoprec_reconstruct"""
        lhs_e = np.ascontiguousarray(lhs_e, dtype=np.complex128)
        rhs_o = np.ascontiguousarray(rhs_o, dtype=np.complex128)
        out = np.empty_like(lhs_e)
        self.orc._f("eoprec_reconstruct")(self.h, _ptr(out), _ptr(lhs_e), _ptr(rhs_o))
        return out

    def __del__(self):
        try:
            self.orc._f("op_free")(self.h)
        except Exception:
            pass


class Oracle:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path
        L = self.lib
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        sig = {
            "kind": (C.c_char_p, []),
            "rng_new": (vp, [C.c_uint]), "rng_free": (None, [vp]),
            "gauss_gauge_u1": (None, [vp, vp, ci, ci, cd]), "unit_gauge_u1": (None, [vp, ci, ci]),
            "gaussian_real": (None, [vp, vp, ci]), "gaussian_complex": (None, [vp, vp, ci]),
            "read_gauge_u1": (ci, [vp, ci, ci, C.c_char_p]), "plaquette_u1": (None, [vp, ci, ci, vp]),
            "dot": (None, [ci, vp, vp, ci, vp]), "norm2sq": (cd, [ci, vp, ci]), "diffnorm2sq": (cd, [ci, vp, vp, ci]),
            "op_prepare": (vp, [C.POINTER(OpDesc)]), "op_free": (None, [vp]), "op_is_complex": (ci, [vp]),
            "op_size": (ci, [vp]), "op_apply": (None, [vp, vp, vp]),
            "eoprec_prepare": (None, [vp, vp, vp]), "eoprec_reconstruct": (None, [vp, vp, vp, vp]),
            "stencil_apply_part": (None, [vp, ci, vp, vp]),
            "solve": (ci, [ci, vp, vp, vp, ci, cd, ci, ci, ci, C.POINTER(Result)]),
            "solve_cg_m": (ci, [vp, C.POINTER(vp), vp, ci, ci, ci, cd, vp, ci, ci, C.POINTER(Result)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, prefix + name)
            f.restype, f.argtypes = res, args
        self.kind = self._f("kind")().decode()

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- input generation (same std::mt19937 stream as the reference's tests) ----
    class Rng:
        def __init__(self, orc, seed):
            self.orc, self.h = orc, orc._f("rng_new")(seed)

        def gauss_gauge_u1(self, X, Y, beta):
            U = np.empty(2 * X * Y, dtype=np.complex128)
            self.orc._f("gauss_gauge_u1")(self.h, _ptr(U), X, Y, beta)
            return U

        def gaussian(self, n, dtype=np.complex128):
            v = np.empty(n, dtype=dtype)
            self.orc._f("gaussian_complex" if dtype == np.complex128 else "gaussian_real")(self.h, _ptr(v), n)
            return v

        def __del__(self):
            try:
                self.orc._f("rng_free")(self.h)
            except Exception:
                pass

    def rng(self, seed=1337):
        return Oracle.Rng(self, seed)

    def unit_gauge(self, X, Y):
        U = np.empty(2 * X * Y, dtype=np.complex128)
        self._f("unit_gauge_u1")(_ptr(U), X, Y)
        return U

    def read_gauge(self, X, Y, path):
        U = np.empty(2 * X * Y, dtype=np.complex128)
        rc = self._f("read_gauge_u1")(_ptr(U), X, Y, path.encode())
        if rc != 0:
            raise IOError("cannot read gauge field %s (rc=%d)" % (path, rc))
        return U

    def plaquette(self, U, X, Y):
        out = np.zeros(2)
        self._f("plaquette_u1")(_ptr(U), X, Y, _ptr(out))
        return complex(out[0], out[1])

    # ---- BLAS-1 ----
    def dot(self, a, b):
        isc = int(a.dtype == np.complex128)
        out = np.zeros(2)
        self._f("dot")(isc, _ptr(a), _ptr(b), a.size, _ptr(out))
        return complex(out[0], out[1]) if isc else out[0]

    def norm2sq(self, a):
        return self._f("norm2sq")(int(a.dtype == np.complex128), _ptr(a), a.size)

    def diffnorm2sq(self, a, b):
        return self._f("diffnorm2sq")(int(a.dtype == np.complex128), _ptr(a), _ptr(b), a.size)

    # ---- operators / solvers ----
    def op(self, kind, X, Y, **kw):
        return Operator(self, kind, X, Y, **kw)

    def solve(self, solver, op, b, x0=None, max_iter=10000, eps=1e-10, restart_freq=0, l=0, verbosity=0):
        b = np.ascontiguousarray(b, dtype=op.dtype)
        x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=op.dtype, copy=True)
        res = Result()
        s = SOLVER[solver] if isinstance(solver, str) else solver
        self._f("solve")(s, op.h, _ptr(x), _ptr(b), max_iter, eps, restart_freq, l, verbosity, C.byref(res))
        return x, res.as_dict()

    def solve_cg_m(self, op, b, shifts, resid_freq_check=10, max_iter=10000, eps=1e-10, worst_first=False,
                   verbosity=0):
        b = np.ascontiguousarray(b, dtype=op.dtype)
        shifts = np.array(shifts, dtype=np.float64, copy=True)
        n = len(shifts)
        xs = [np.zeros_like(b) for _ in range(n)]
        ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        res = Result()
        self._f("solve_cg_m")(op.h, ptrs, _ptr(b), n, resid_freq_check, max_iter, eps, _ptr(shifts),
                              int(worst_first), verbosity, C.byref(res))
        # the solver may have permuted the pointer array; map back by address
        by_addr = {x.ctypes.data: x for x in xs}
        xs = [by_addr[ptrs[i]] for i in range(n)]
        return xs, res.as_dict(), shifts


class RefMg:
    """The reference's own adaptive multigrid (oracle/ref_mg_shim.cpp; `ref` library only): two or more levels on
    the 2-D U(1) staggered operator, built from caller-supplied null vectors.  null[l] = list of nvec[l] arrays."""
    # mg.h:4-13 inner_solver, generic_inverters.h minv_inverter, mg_complex.h:17-21 mg_multilevel_type
    INNER = dict(NONE=0, MINRES=1, CG=2, GCR=3, BICGSTAB=4, CR=5, BICGSTAB_L=6)
    SMOOTH = dict(CG=0, CR=1, GCR=2, BICGSTAB=3, BICGSTAB_L=4, GMRES=5, SOR=6, MINRES=7, INVALID=-1)

    def __init__(self, orc, X, Y, links, mass, blocks, nvecs, null, ignore_shifts=False):
        """ignore_shifts: build the stencils the way the reference's driver does (mass in the shift of every level,
        generate_coarse_from_fine_stencil(..., true)); default: shifts folded into the coarse clover."""
        self._bind(orc)
        self.n_refine = len(blocks)
        self.links = np.ascontiguousarray(links, dtype=np.complex128)
        blocks_a = np.array(blocks, dtype=np.int32)
        nvecs_a = np.array(nvecs, dtype=np.int32)
        self._null = [[np.ascontiguousarray(v, dtype=np.complex128) for v in lvl] for lvl in null]
        lvl_ptrs = []
        for lvl in self._null:
            lvl_ptrs.append((C.c_void_p * len(lvl))(*[v.ctypes.data for v in lvl]))
        top = (C.c_void_p * len(lvl_ptrs))(*[C.cast(a, C.c_void_p).value for a in lvl_ptrs])
        self._keep = (lvl_ptrs, top, blocks_a, nvecs_a)
        self.h = self.L.refmg_create2(X, Y, _ptr(self.links), mass, self.n_refine, _ptr(blocks_a), _ptr(nvecs_a),
                                      C.cast(top, C.c_void_p), int(ignore_shifts))

    def _bind(self, orc):
        if orc.kind != "reference":
            raise RuntimeError("RefMg needs oracle/_ref/libref_oracle.so (the reference-compiled checker)")
        L = orc.lib
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        for name, res, args in (("refmg_create2", vp, [ci, ci, vp, cd, ci, vp, vp, vp, ci]),
                                ("refmg_setup", vp, [ci, ci, vp, cd, ci, vp, vp, ci, cd, ci, vp, vp, ci, ci, ci, ci,
                                                     C.c_uint, ci, ci, ci]),
                                ("refmg_null_counts", None, [vp, vp]),
                                ("refmg_level_dims", None, [vp, ci, vp, vp, vp]),
                                ("refmg_get_null", None, [vp, ci, ci, vp]),
                                ("refmg_get_stencil", None, [vp, ci, vp, vp, vp]),
                                ("refmg_prolong", None, [vp, ci, vp, vp]), ("refmg_restrict", None, [vp, ci, vp, vp]),
                                ("refmg_apply_level", None, [vp, ci, vp, vp]),
                                ("refmg_set_precond", None, [vp, ci, ci, ci, ci, ci, ci, cd, ci]),
                                ("refmg_apply_level_variant", None, [vp, ci, ci, vp, vp]),
                                ("refmg_set_normal", None, [vp, ci, ci, ci]),
                                ("refmg_counts", None, [vp, vp]),
                                ("refmg_vcycle", None, [vp, vp, vp]),
                                ("refmg_vpgcr", None, [vp, vp, vp, ci, cd, ci, ci, vp])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.L = L

    @classmethod
    def setup(cls, orc, X, Y, links, mass, blocks, nvecs, bstrat=1, null_mass=1e-2, null_gen="BICGSTAB", tol=5e-5,
              max_iter=500, restart_freq=0, bicgstab_l=-1, do_ortho_eo=False, do_global_ortho_conj=False, seed=1337,
              verbosity=0, null_prec=0, do_free=False):
        """The reference driver's complete set-up (aa_mg_square_staggered_u1.cpp:716-1143; oracle/ref_mg_shim.cpp
        refmg_setup): null vectors from null_generate_random_smooth with a std::mt19937(seed), block_orthonormalize,
        generate_coarse_from_fine_stencil(ignore_shifts=true) with the shift copied down.  nvecs[l] = total vectors
        of refinement l (after the partition); bstrat 0 = BLOCK_NONE, 1 = BLOCK_EO, 2 = BLOCK_CORNER; do_free: null_generate_free instead of
        the smoothing solves; null_prec 0 = none, 1 = even/odd
        (top/bottom below the top level), 2 = normal equations (null_gen.h:24-29)."""
        self = cls.__new__(cls)
        self._bind(orc)
        self.n_refine = len(blocks)
        self.links = np.ascontiguousarray(links, dtype=np.complex128)
        blocks_a = np.array(blocks, dtype=np.int32)
        nvecs_a = np.array(nvecs, dtype=np.int32)
        tol_a = np.array([tol] * self.n_refine if np.isscalar(tol) else tol, dtype=np.float64)
        it_a = np.array([max_iter] * self.n_refine if np.isscalar(max_iter) else max_iter, dtype=np.int32)
        self._keep = (blocks_a, nvecs_a, tol_a, it_a)
        self.h = self.L.refmg_setup(X, Y, _ptr(self.links), mass, self.n_refine, _ptr(blocks_a), _ptr(nvecs_a), bstrat,
                                    null_mass, self.SMOOTH[null_gen], _ptr(tol_a), _ptr(it_a), restart_freq, bicgstab_l,
                                    int(do_ortho_eo), int(do_global_ortho_conj), seed, verbosity, null_prec,
                                    int(do_free))
        return self

    def null_counts(self):
        out = np.zeros(self.n_refine + 1, dtype=np.int32)
        self.L.refmg_null_counts(self.h, _ptr(out))
        return [int(v) for v in out]

    def dims(self, level):
        x, y, n = C.c_int(), C.c_int(), C.c_int()
        self.L.refmg_level_dims(self.h, level, C.byref(x), C.byref(y), C.byref(n))
        return x.value, y.value, n.value

    def size(self, level):
        x, y, n = self.dims(level)
        return x * y * n

    def null(self, level, v):
        out = np.empty(self.size(level), dtype=np.complex128)
        self.L.refmg_get_null(self.h, level, v, _ptr(out))
        return out

    def stencil(self, level):
        x, y, n = self.dims(level)
        cl = np.empty(x * y * n * n, dtype=np.complex128)
        hp = np.empty(4 * x * y * n * n, dtype=np.complex128)
        sh = np.empty(3, dtype=np.complex128)
        self.L.refmg_get_stencil(self.h, level, _ptr(cl), _ptr(hp), _ptr(sh))
        return cl, hp, sh

    def prolong(self, level, coarse):
        coarse = np.ascontiguousarray(coarse, dtype=np.complex128)
        fine = np.empty(self.size(level), dtype=np.complex128)
        self.L.refmg_prolong(self.h, level, _ptr(fine), _ptr(coarse))
        return fine

    def restrict(self, level, fine):
        fine = np.ascontiguousarray(fine, dtype=np.complex128)
        coarse = np.empty(self.size(level + 1), dtype=np.complex128)
        self.L.refmg_restrict(self.h, level, _ptr(coarse), _ptr(fine))
        return coarse

    def apply_level(self, level, v):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.zeros_like(v)
        self.L.refmg_apply_level(self.h, level, _ptr(out), _ptr(v))
        return out

    def apply_level_variant(self, level, v, which):
        """which = "dagger" / "normal": fine_ / coarse_square_staggered_dagger / _normal at that level"""
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.zeros_like(v)
        self.L.refmg_apply_level_variant(self.h, level, dict(plain=0, dagger=1, normal=2)[which], _ptr(out), _ptr(v))
        return out

    def set_precond(self, smooth="GCR", n_pre=6, n_post=6, inner="GCR", n_max=1024, n_restart=64, rel_res=1e-2,
                    recursive=False):
        self.L.refmg_set_precond(self.h, self.SMOOTH[smooth], n_pre, n_post, self.INNER[inner], n_max, n_restart,
                                 rel_res, 1 if recursive else 0)

    def set_normal(self, normal_smooth, normal_mg, ignore_shifts=False, dagger_stencils=True):
        """normal-equation variants of the cycle (mg_precond_struct_complex::normal_eqn_smooth / normal_eqn_mg) with
        dagger stencils on every level, as the reference's driver wires them; dagger_stencils=False leaves them out"""
        self.L.refmg_set_normal(self.h, int(normal_smooth), int(normal_mg), int(ignore_shifts) if dagger_stencils else -1)

    def counts(self):
        n = self.n_refine + 1
        out = np.zeros(4 * n, dtype=np.int32)
        self.L.refmg_counts(self.h, _ptr(out))
        return dict(krylov=out[:n].tolist(), presmooth=out[n:2 * n].tolist(), postsmooth=out[2 * n:3 * n].tolist(),
                    residual=out[3 * n:].tolist())

    def vcycle(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.complex128)
        out = np.zeros_like(rhs)
        self.L.refmg_vcycle(self.h, _ptr(out), _ptr(rhs))
        return out

    def vpgcr(self, b, x0=None, max_iter=1000, eps=5e-7, restart_freq=64, verbosity=0):
        b = np.ascontiguousarray(b, dtype=np.complex128)
        x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.complex128, copy=True)
        out = np.zeros(4)
        self.L.refmg_vpgcr(self.h, _ptr(x), _ptr(b), max_iter, eps, restart_freq, verbosity, _ptr(out))
        return x, dict(resSq=out[0], iter=int(out[1]), success=bool(out[2]), ops_count=int(out[3]))


MULTI = dict(CG_M=0, CR_M=1, BICGSTAB_M=2)
PRECOND_SOLVER = dict(PCG=0, FPCG=1, FPCG_RESTART=2, VPGCR=3, VPGCR_RESTART=4, PBICGSTAB=5, PBICGSTAB_RESTART=6)
PRECOND = dict(IDENTITY=0, GCR=1, MINRES=2)


def ref_solve_multi(orc, which, op, b, shifts, resid_freq_check=10, max_iter=10000, eps=1e-10, worst_first=False):
    """minv_vector_{cg,cr,bicgstab}_m of the reference (`ref` library only)"""
    f = orc.lib.ref_solve_multi
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                  C.c_void_p, C.c_int, C.c_int, C.POINTER(Result)]
    b = np.ascontiguousarray(b, dtype=op.dtype)
    shifts = np.array(shifts, dtype=np.float64, copy=True)
    n = len(shifts)
    xs = [np.zeros_like(b) for _ in range(n)]
    ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
    res = Result()
    f(MULTI[which], op.h, ptrs, _ptr(b), n, resid_freq_check, max_iter, eps, _ptr(shifts), int(worst_first), 0,
      C.byref(res))
    by_addr = {x.ctypes.data: x for x in xs}
    return [by_addr[ptrs[i]] for i in range(n)], res.as_dict(), shifts


def ref_stencil_prec(orc, op, top_bottom, a, b=None):
    """apply_square_staggered_{eo,tb}prec_prepare_stencil (b is None) / _reconstruct_stencil (a = lhs_part,
    b = rhs_other) of the reference on a stencil operator (`ref` library only)"""
    f = orc.lib.ref_stencil_prec
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    a = np.ascontiguousarray(a, dtype=np.complex128)
    out = np.zeros_like(a)
    if b is None:
        f(op.h, int(top_bottom), 0, _ptr(out), _ptr(a), None)
    else:
        b = np.ascontiguousarray(b, dtype=np.complex128)
        f(op.h, int(top_bottom), 1, _ptr(out), _ptr(a), _ptr(b))
    return out


def ref_solve_relax(orc, which, op, b, x0=None, max_iter=10000, eps=1e-10, omega=1.0):
    """minv_vector_sor / minv_vector_minres of the reference (`ref` library only); which = "SOR" | "MINRES" """
    f = orc.lib.ref_solve_relax
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(Result)]
    b = np.ascontiguousarray(b, dtype=op.dtype)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=op.dtype, copy=True)
    res = Result()
    f(dict(SOR=0, MINRES=1)[which], op.h, _ptr(x), _ptr(b), max_iter, eps, omega, 0, C.byref(res))
    return x, res.as_dict()


def ref_solve_precond(orc, solver, op, b, x0=None, max_iter=10000, eps=1e-10, restart_freq=0, precond="IDENTITY",
                      n_step=4, rel_res=1e-20):
    """the reference's preconditioned family with its stock preconditioners (`ref` library only)"""
    f = orc.lib.ref_solve_precond
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                  C.c_double, C.c_int, C.POINTER(Result)]
    b = np.ascontiguousarray(b, dtype=op.dtype)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=op.dtype, copy=True)
    res = Result()
    f(PRECOND_SOLVER[solver], op.h, _ptr(x), _ptr(b), max_iter, eps, restart_freq, PRECOND[precond], n_step, rel_res, 0,
      C.byref(res))
    return x, res.as_dict()


def available():
    out = []
    if os.path.exists(os.path.join(HERE, "_ref", "libref_oracle.so")):
        out.append("ref")
    if os.path.exists(os.path.join(HERE, "libport_oracle.so")):
        out.append("port")
    return out


def load(which="best"):
    if which == "best":
        av = available()
        if not av:
            raise RuntimeError("no oracle library built: run `make -C oracle` (or __graft_entry__.build())")
        which = av[0]
    if which == "ref":
        return Oracle(os.path.join(HERE, "_ref", "libref_oracle.so"), "ref_")
    if which == "port":
        return Oracle(os.path.join(HERE, "libport_oracle.so"), "port_")
    raise ValueError(which)
