"""generic-linalg_b200 -- Python harness over the two in-tree native libraries.

    libglb200.so            CUDA kernels (sm_100a) + runtime + the C ABI of include/glb200.h
    libglb200_inverters.so  C++ drop-in solver shells with the reference's signatures

The product is the native code; this module is plumbing for tests/ and bench.py (ctypes
bindings, numpy <-> device copies, torch.distributed bootstrap of the slab communicator).
It never computes anything itself and there is no CPU fallback: importing works anywhere,
but creating a Context without the built libraries or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_CUDA = os.path.join(HERE, "libglb200.so")
LIB_HOST = os.path.join(HERE, "libglb200_inverters.so")

REAL, COMPLEX = 0, 1
# composite operators on a stencil2d operator (include/glb200.h GLB_SV_*)
STENCIL_VIEW = dict(NONE=0, M2MDEODOE=1, M2MDTBDBT=2, NORMAL_EO=3, NORMAL_TB=4, DAGGER_EO=5, DAGGER_TB=6)
STAG_DAGGER, STAG_GAMMA5, STAG_NORMAL = 1, 2, 4
STAG_DEO, STAG_DOE, STAG_M2MDEODOE = 8, 16, 32   # even/odd pieces (operators.cpp:456-571)

# operator / solver selectors of host/capi_solvers.cpp (same numbering as oracle/oracle_api.h)
OP = dict(LAPLACE_REAL=0, LAPLACE_IMAG=1, LAPLACE_NC=2, LAPLACE_U1=3, STAG_FREE=4, STAG_U1=5,
          STAG_GAMMA5_U1=6, STAG_DAGGER_U1=7, STAG_NORMAL_U1=8, GAMMA5=9, STENCIL=10,
          STENCIL_FROM_STAG=11, STAG_GAMMA5_FREE=12, LAPLACE_REAL_NC=13, STAG_FREE_REAL=14,
          STAG_DEO_U1=15, STAG_DOE_U1=16, STAG_M2MDEODOE_U1=17, SYMMSHIFT_X=18, SYMMSHIFT_Y=19,
          STAG_2LINK_U1=20, STAG_INDEX=21)
SOLVER = dict(CG=0, CG_RESTART=1, CR=2, CR_RESTART=3, GCR=4, GCR_RESTART=5, BICGSTAB=6,
              BICGSTAB_RESTART=7, BICGSTAB_L=8, BICGSTAB_L_RESTART=9, GMRES=10, GMRES_RESTART=11)


class GlbError(RuntimeError):
    pass


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


class CgReport(C.Structure):
    _fields_ = [("iterations", C.c_int), ("ops", C.c_int), ("hit_max_iter", C.c_int),
                ("rsq", C.c_double), ("bnorm", C.c_double)]


_libs = None


def libs():
    """Load both native libraries (once).  Raises GlbError when they were not built."""
    global _libs
    if _libs is not None:
        return _libs
    for p in (LIB_CUDA, LIB_HOST):
        if not os.path.exists(p):
            raise GlbError("native library %s is missing: run __graft_entry__.build() "
                           "(make -C generic-linalg_b200); there is no Python/CPU fallback" % p)
    # RTLD_LOCAL: the drop-in library exports the reference's own symbol names (minv_vector_cg, ...);
    # keeping them out of the global scope stops them interposing on the reference-compiled oracle.
    cu = C.CDLL(LIB_CUDA, mode=C.RTLD_LOCAL)
    ho = C.CDLL(LIB_HOST, mode=C.RTLD_LOCAL)
    vp, ci, cd, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    pd = C.POINTER(C.c_double)
    sig = {
        "glb_create": (ci, [ci, C.POINTER(vp)]), "glb_destroy": (ci, [vp]),
        "glb_last_error": (C.c_char_p, []), "glb_synchronize": (ci, [vp]), "glb_stream": (vp, [vp]),
        "glb_device": (ci, [vp]), "glb_sm_count": (ci, [vp]), "glb_kernel_launches": (C.c_ulonglong, []),
        "glb_prof_enable": (ci, [vp, ci]), "glb_prof_read": (ci, [vp, ci, ci, C.POINTER(C.c_float), C.POINTER(ci)]),
        "glb_prof_summary": (ci, [vp, ci, C.POINTER(ci), pd, pd]),
        "glb_comm_unique_id": (ci, [C.c_char_p]), "glb_comm_init": (ci, [vp, ci, ci, C.c_char_p]),
        "glb_comm_p2p_enabled": (ci, [vp]), "glb_comm_rank": (ci, [vp]), "glb_comm_size": (ci, [vp]), "glb_comm_barrier": (ci, [vp]),
        "glb_vec_alloc": (ci, [vp, ci, sz, C.POINTER(vp)]), "glb_vec_free": (ci, [vp, vp]),
        "glb_vec_upload": (ci, [vp, ci, sz, vp, vp]), "glb_vec_download": (ci, [vp, ci, sz, vp, vp]),
        "glb_vec_zero": (ci, [vp, ci, sz, vp]), "glb_vec_copy": (ci, [vp, ci, sz, vp, vp]),
        "glb_host_alloc": (ci, [vp, sz, C.POINTER(vp)]), "glb_host_free": (ci, [vp, vp]),
        "glb_op_create_laplace": (ci, [vp, ci, ci, ci, ci, cd, cd, C.POINTER(vp)]),
        "glb_op_create_laplace_u1": (ci, [vp, vp, ci, ci, cd, C.POINTER(vp)]),
        "glb_op_create_staggered_free_real": (ci, [vp, ci, ci, cd, C.POINTER(vp)]),
        "glb_op_create_staggered": (ci, [vp, vp, ci, ci, cd, C.c_uint, C.POINTER(vp)]),
        "glb_op_create_staggered_local": (ci, [vp, vp, ci, ci, cd, C.c_uint, C.POINTER(vp)]),
        "glb_slab_bounds": (ci, [vp, ci, C.POINTER(ci), C.POINTER(ci)]),
        "glb_op_create_gamma5": (ci, [vp, ci, ci, C.POINTER(vp)]),
        "glb_op_create_stencil2d": (ci, [vp, vp, vp, vp, ci, ci, ci, pd, pd, pd, C.POINTER(vp)]),
        "glb_op_destroy": (ci, [vp]), "glb_op_set_mass": (ci, [vp, cd]), "glb_op_dtype": (ci, [vp]),
        "glb_op_local_size": (sz, [vp]), "glb_op_global_size": (sz, [vp]),
        "glb_op_apply": (ci, [vp, vp, vp]), "glb_op_apply_dot": (ci, [vp, vp, vp, vp, ci, pd]),
        "glb_op_bytes_per_apply": (cd, [vp]),
        "glb_dot": (ci, [vp, ci, sz, vp, vp, pd]), "glb_norm2sq": (ci, [vp, ci, sz, vp, pd]),
        "glb_diffnorm2sq": (ci, [vp, ci, sz, vp, vp, pd]), "glb_dot_norm": (ci, [vp, ci, sz, vp, vp, pd]),
        "glb_multi_dot": (ci, [vp, ci, sz, ci, C.POINTER(vp), vp, pd]),
        "glb_sub": (ci, [vp, ci, sz, vp, vp, vp]), "glb_add": (ci, [vp, ci, sz, vp, vp, vp]),
        "glb_axpy": (ci, [vp, ci, sz, pd, vp, vp]), "glb_xpay": (ci, [vp, ci, sz, vp, pd, vp]),
        "glb_axpyz": (ci, [vp, ci, sz, pd, vp, vp, vp]), "glb_rdiv": (ci, [vp, ci, sz, vp, cd, vp]),
        "glb_axpy_norm": (ci, [vp, ci, sz, pd, vp, vp, pd]),
        "glb_lincomb": (ci, [vp, ci, sz, ci, pd, C.POINTER(vp), vp, vp]),
        "glb_update_xr_norm": (ci, [vp, ci, sz, pd, vp, vp, pd, vp, vp, pd]),
        "glb_update_p_ap_norm": (ci, [vp, ci, sz, vp, vp, pd, vp, vp, pd]),
        "glb_bicgstab_update": (ci, [vp, ci, sz, pd, vp, pd, vp, vp, vp, vp, vp, pd]),
        "glb_bicgstab_pupdate": (ci, [vp, ci, sz, vp, pd, pd, vp, vp]),
        "glb_conj": (ci, [vp, ci, sz, vp, vp]), "glb_bicgstabm_update_s": (ci, [vp, ci, sz, pd, vp, vp, vp, vp]),
        "glb_cgm_update_x": (ci, [vp, ci, sz, ci, pd, C.POINTER(vp), C.POINTER(vp)]),
        "glb_cgm_update_p": (ci, [vp, ci, sz, ci, pd, pd, vp, C.POINTER(vp)]),
        "glb_cg_solve_supported": (ci, [vp]),
        "glb_cg_last_pred_err": (cd, []), "glb_cg_step_mode": (ci, [ci, ci]),
        "glb_cg_solve": (ci, [vp, vp, vp, ci, cd, C.POINTER(CgReport), pd, ci]),
        "glb_krylov_solve_supported": (ci, [vp, ci]),
        "glb_krylov_solve": (ci, [vp, ci, vp, vp, ci, cd, C.POINTER(CgReport), pd, ci]),
        "glb_krylov_graph_mode": (ci, [ci]), "glb_krylov_last_used_graph": (ci, []),
        "glb_op_apply_part": (ci, [vp, vp, vp, ci]),
        "glb_stag_eoprec_prepare": (ci, [vp, vp, vp]), "glb_stag_eoprec_reconstruct": (ci, [vp, vp, vp, vp]),
        "glb_mg_transfer_create": (ci, [vp, ci, ci, ci, ci, ci, ci, C.POINTER(vp), C.POINTER(vp)]),
        "glb_mg_transfer_destroy": (ci, [vp]), "glb_mg_fine_size": (sz, [vp]), "glb_mg_coarse_size": (sz, [vp]),
        "glb_mg_prolong": (ci, [vp, vp, vp]), "glb_mg_restrict": (ci, [vp, vp, vp]),
        "glb_rscale": (ci, [vp, ci, sz, vp, cd, vp]),
        "glb_op_create_stencil_view": (ci, [vp, ci, ci, C.POINTER(vp)]),
        "glb_stencil_prec_prepare": (ci, [vp, ci, vp, vp]), "glb_stencil_prec_reconstruct": (ci, [vp, ci, vp, vp, vp]),
        "glb_op_set_shifts": (ci, [vp, pd, pd, pd]), "glb_op_get_shifts": (ci, [vp, pd, pd, pd]),
        "glb_op_stencil_download": (ci, [vp, vp, vp]),
        "glb_mg_transfer_create_dev": (ci, [vp, ci, ci, ci, ci, ci, ci, C.POINTER(vp), C.POINTER(vp)]),
        "glb_mg_block_orthonormalize": (ci, [vp, ci, ci, ci, ci, ci, ci, C.POINTER(vp)]),
        "glb_mg_partition": (ci, [vp, ci, ci, ci, ci, vp, vp]),
        "glb_mg_partition_corner": (ci, [vp, ci, ci, ci, ci, ci, vp, vp]),
        "glb_mg_galerkin": (ci, [vp, vp, ci, C.POINTER(vp)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(cu, name)
        f.restype, f.argtypes = res, args
    hsig = {
        "glbx_default_context": (vp, []), "glbx_set_default_context": (None, [vp]),
        "glbx_force_host_scalars": (None, [ci]), "glbx_allow_host_callback_shim": (None, [ci]),
        "glbx_cache_operators": (None, [ci]),
        "glbx_synthetic_inputs": (ci, [C.c_uint, ci, ci, cd, vp, vp]),
        "glbx_host_apply": (ci, [C.POINTER(OpDesc), vp, vp]),
        "glbx_host_solve": (ci, [ci, C.POINTER(OpDesc), vp, vp, ci, cd, ci, ci, ci, C.POINTER(Result)]),
        "glbx_host_solve_cg_m": (ci, [C.POINTER(OpDesc), C.POINTER(vp), vp, ci, ci, ci, cd, vp, ci, ci,
                                      C.POINTER(Result)]),
        "glbx_dev_solve": (ci, [ci, vp, vp, vp, ci, cd, ci, ci, ci, C.POINTER(Result)]),
        "glbx_dev_solve_cg_m": (ci, [vp, C.POINTER(vp), vp, ci, ci, ci, cd, vp, ci, ci, C.POINTER(Result)]),
        "glbx_host_stencil_part": (ci, [C.POINTER(OpDesc), ci, vp, vp]),
        "glbx_host_eoprec_prepare": (ci, [C.POINTER(OpDesc), vp, vp]),
        "glbx_host_eoprec_reconstruct": (ci, [C.POINTER(OpDesc), vp, vp, vp]),
        "glbx_host_solve_multi": (ci, [ci, C.POINTER(OpDesc), C.POINTER(vp), vp, ci, ci, ci, cd, vp, ci, ci,
                                       C.POINTER(Result)]),
        "glbx_host_solve_precond": (ci, [ci, C.POINTER(OpDesc), vp, vp, ci, cd, ci, ci, ci, cd, ci, C.POINTER(Result)]),
        "glbx_host_stencil_prec": (ci, [C.POINTER(OpDesc), ci, ci, vp, vp, vp]),
        "glbx_host_solve_relax": (ci, [ci, C.POINTER(OpDesc), vp, vp, ci, cd, cd, ci, C.POINTER(Result)]),
        "glbx_dev_solve_relax": (ci, [ci, vp, vp, vp, ci, cd, cd, ci, C.POINTER(Result)]),
        "glbx_mg_create": (vp, [ci, C.POINTER(vp), C.POINTER(vp)]), "glbx_mg_destroy": (None, [vp]),
        "glbx_mg_set": (None, [vp, ci, ci, ci, ci, ci, ci, cd, ci, ci]),
        "glbx_mg_vcycle": (ci, [vp, vp, vp]),
        "glbx_mg_vpgcr": (ci, [vp, vp, vp, ci, cd, ci, ci, C.POINTER(Result)]),
        "glbx_mg_counts": (None, [vp, C.POINTER(ci)]),
        "glbx_mg_setup": (vp, [vp, ci, ci, ci, C.POINTER(ci), C.POINTER(ci), ci, cd, ci, pd, C.POINTER(ci), ci, ci, ci,
                               ci, C.c_uint, ci, ci, ci, vp]),
        "glbx_mg_level_op": (vp, [vp, ci]), "glbx_mg_level_transfer": (vp, [vp, ci]),
        "glbx_mg_null_vector": (vp, [vp, ci, ci]), "glbx_mg_setup_seconds": (None, [vp, pd]),
    }
    for name, (res, args) in hsig.items():
        f = getattr(ho, name)
        f.restype, f.argtypes = res, args
    _libs = (cu, ho)
    return _libs


def slab_bounds(Y, rank, nranks):
    """rows [y0, y0+Yloc) of a Y-row lattice owned by `rank` -- the same arithmetic as slab_of() in
    csrc/ops.cu (pure host logic, testable without a GPU)."""
    y0 = (Y * rank) // nranks
    y1 = (Y * (rank + 1)) // nranks
    return y0, y1 - y0


def exported_symbols():
    """Names declared in include/glb200.h (used by the CPU test that checks the ABI surface)."""
    import re
    hdr = open(os.path.join(HERE, "..", "include", "glb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(glb_[a-z0-9_]+)\s*\(", hdr)))


def _chk(rc, what=""):
    if rc != 0:
        cu, _ = libs()
        raise GlbError("%s failed (code %d): %s" % (what, rc, cu.glb_last_error().decode()))


def _dt(arr_or_dtype):
    d = np.dtype(arr_or_dtype) if isinstance(arr_or_dtype, (type, np.dtype, str)) else arr_or_dtype.dtype
    if d == np.complex128:
        return COMPLEX
    if d == np.float64:
        return REAL
    raise TypeError("only float64 / complex128 vectors exist on this path")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c2(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)


class DeviceVector:
    """A device-resident vector (slab-local on multi-GPU runs)."""

    def __init__(self, ctx, n, dtype):
        self.ctx, self.n, self.dtype = ctx, int(n), np.dtype(dtype)
        self.dt = _dt(self.dtype)
        p = C.c_void_p()
        _chk(ctx.cu.glb_vec_alloc(ctx.h, self.dt, self.n, C.byref(p)), "glb_vec_alloc")
        self.ptr = p.value

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == self.n
        _chk(self.ctx.cu.glb_vec_upload(self.ctx.h, self.dt, self.n, self.ptr, _p(host)), "glb_vec_upload")
        return self

    def download(self, out=None):
        out = np.empty(self.n, dtype=self.dtype) if out is None else out
        _chk(self.ctx.cu.glb_vec_download(self.ctx.h, self.dt, self.n, _p(out), self.ptr), "glb_vec_download")
        return out

    def zero(self):
        _chk(self.ctx.cu.glb_vec_zero(self.ctx.h, self.dt, self.n, self.ptr), "glb_vec_zero")
        return self

    def copy_from(self, other):
        _chk(self.ctx.cu.glb_vec_copy(self.ctx.h, self.dt, self.n, self.ptr, other.ptr), "glb_vec_copy")
        return self

    def free(self):
        if self.ptr:
            self.ctx.cu.glb_vec_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Operator:
    def __init__(self, ctx, handle, keep=()):
        self.ctx, self.h, self._keep = ctx, handle, keep
        cu = ctx.cu
        self.dtype = np.complex128 if cu.glb_op_dtype(handle) == COMPLEX else np.float64
        self.local_size = cu.glb_op_local_size(handle)
        self.global_size = cu.glb_op_global_size(handle)
        self.bytes_per_apply = cu.glb_op_bytes_per_apply(handle)

    def apply(self, out, inp):
        _chk(self.ctx.cu.glb_op_apply(self.h, out.ptr, inp.ptr), "glb_op_apply")

    def apply_dot(self, out, inp, w=None, want_norm=False):
        d = (C.c_double * 3)()
        _chk(self.ctx.cu.glb_op_apply_dot(self.h, out.ptr, inp.ptr, w.ptr if w is not None else None,
                                          int(want_norm), d), "glb_op_apply_dot")
        return complex(d[0], d[1]), d[2]

    PART = dict(EO=1, OE=2, TB=3, BT=4)

    def apply_part(self, part, out, inp):
        """glb_op_apply_part: apply_stencil_2d_{eo,oe,tb,bt} on a stencil2d operator (coarse_stencil.cpp:395-1512)"""
        _chk(self.ctx.cu.glb_op_apply_part(self.h, out.ptr, inp.ptr, self.PART[part]), "glb_op_apply_part")

    def eoprec_prepare(self, rhs_e, rhs_orig):
        """glb_stag_eoprec_prepare: rhs_e = m rhs - D_eo rhs on even sites, 0 on odd (operators.cpp:528)"""
        _chk(self.ctx.cu.glb_stag_eoprec_prepare(self.h, rhs_e.ptr, rhs_orig.ptr), "glb_stag_eoprec_prepare")

    def eoprec_reconstruct(self, lhs_full, lhs_e, rhs_o):
        """glb_stag_eoprec_reconstruct: even sites lhs_e, odd sites (rhs_o - D_oe lhs_e)/m (operators.cpp:574)"""
        _chk(self.ctx.cu.glb_stag_eoprec_reconstruct(self.h, lhs_full.ptr, lhs_e.ptr, rhs_o.ptr),
             "glb_stag_eoprec_reconstruct")

    def set_mass(self, m):
        _chk(self.ctx.cu.glb_op_set_mass(self.h, m))

    def view(self, kind):
        """glb_op_create_stencil_view: a composite operator (STENCIL_VIEW name) sharing this stencil2d operator's
        matrices and shifts; it must not outlive this operator"""
        h = C.c_void_p()
        _chk(self.ctx.cu.glb_op_create_stencil_view(self.h, STENCIL_VIEW[kind], 0, C.byref(h)), "glb_op_create_stencil_view")
        return Operator(self.ctx, h, keep=(self,))

    def prec_prepare(self, top_bottom, rhs_part, rhs_orig):
        """glb_stencil_prec_prepare: apply_square_staggered_{eo,tb}prec_prepare_stencil"""
        _chk(self.ctx.cu.glb_stencil_prec_prepare(self.h, int(top_bottom), rhs_part.ptr, rhs_orig.ptr),
             "glb_stencil_prec_prepare")

    def prec_reconstruct(self, top_bottom, lhs_full, lhs_part, rhs_other):
        """glb_stencil_prec_reconstruct: apply_square_staggered_{eo,tb}prec_reconstruct_stencil"""
        _chk(self.ctx.cu.glb_stencil_prec_reconstruct(self.h, int(top_bottom), lhs_full.ptr, lhs_part.ptr, rhs_other.ptr),
             "glb_stencil_prec_reconstruct")

    def set_shifts(self, shift=None, eo_shift=None, dof_shift=None):
        """stencil2d operators: stencil_2d::shift / eo_shift / dof_shift (coarse_stencil.h:62-71); None keeps one"""
        a = [_c2(v) if v is not None else None for v in (shift, eo_shift, dof_shift)]
        _chk(self.ctx.cu.glb_op_set_shifts(self.h, a[0], a[1], a[2]), "glb_op_set_shifts")

    def get_shifts(self):
        a = [(C.c_double * 2)() for _ in range(3)]
        _chk(self.ctx.cu.glb_op_get_shifts(self.h, a[0], a[1], a[2]), "glb_op_get_shifts")
        return tuple(complex(v[0], v[1]) for v in a)

    def stencil_download(self, X, Y, nc):
        """(clover, hopping) of a single-rank stencil2d operator in the reference layout"""
        cl = np.empty(X * Y * nc * nc, dtype=np.complex128)
        hp = np.empty(4 * X * Y * nc * nc, dtype=np.complex128)
        _chk(self.ctx.cu.glb_op_stencil_download(self.h, _p(cl), _p(hp)), "glb_op_stencil_download")
        return cl, hp

    def apply_host(self, v):
        """upload -> apply -> download (convenience for tests)"""
        a = self.ctx.vector(self.local_size, self.dtype).upload(v)
        b = self.ctx.vector(self.local_size, self.dtype)
        self.apply(b, a)
        return b.download()

    def destroy(self):
        if self.h:
            self.ctx.cu.glb_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class MgTransfer:
    """prolong / restrict between two multigrid levels (include/glb200.h: glb_mg_*; mg_complex.cpp:372-467)"""

    def __init__(self, ctx, Xf, Yf, dof_f, bx, by, null_vectors):
        """null_vectors: host arrays, or DeviceVectors (e.g. after Context.mg_block_orthonormalize)"""
        self.ctx = ctx
        self.dims = (Xf, Yf, dof_f, bx, by)
        n = len(null_vectors)
        h = C.c_void_p()
        if n and isinstance(null_vectors[0], DeviceVector):
            ptrs = (C.c_void_p * n)(*[v.ptr for v in null_vectors])
            _chk(ctx.cu.glb_mg_transfer_create_dev(ctx.h, Xf, Yf, dof_f, bx, by, n, ptrs, C.byref(h)),
                 "glb_mg_transfer_create_dev")
        else:
            self._null = [np.ascontiguousarray(v, dtype=np.complex128) for v in null_vectors]
            ptrs = (C.c_void_p * n)(*[v.ctypes.data for v in self._null])
            _chk(ctx.cu.glb_mg_transfer_create(ctx.h, Xf, Yf, dof_f, bx, by, n, ptrs, C.byref(h)),
                 "glb_mg_transfer_create")
        self.nvec = n
        self.h = h
        self.fine_size = int(ctx.cu.glb_mg_fine_size(h))
        self.coarse_size = int(ctx.cu.glb_mg_coarse_size(h))

    def prolong(self, fine, coarse):
        _chk(self.ctx.cu.glb_mg_prolong(self.h, fine.ptr, coarse.ptr), "glb_mg_prolong")

    def restrict(self, coarse, fine):
        _chk(self.ctx.cu.glb_mg_restrict(self.h, coarse.ptr, fine.ptr), "glb_mg_restrict")

    def galerkin(self, fine_op, ignore_shifts=False):
        """glb_mg_galerkin: the coarse stencil2d operator P^dag A P (generate_coarse_from_fine_stencil,
        mg_complex.cpp:827-1026) of a five-point stencil2d fine operator"""
        h = C.c_void_p()
        _chk(self.ctx.cu.glb_mg_galerkin(self.h, fine_op.h, int(ignore_shifts), C.byref(h)), "glb_mg_galerkin")
        return Operator(self.ctx, h)

    def destroy(self):
        if self.h:
            self.ctx.cu.glb_mg_transfer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Multigrid:
    """The device multigrid preconditioner (host/mg_complex.h): ops[0] fine ... ops[n_refine] coarsest,
    transfers[l] couples level l to l+1.  Defaults are the reference's (input_params.cpp:751-800)."""
    INNER = dict(NONE=0, MINRES=1, CG=2, GCR=3, BICGSTAB=4, CR=5, BICGSTAB_L=6)
    SMOOTH = dict(CG=0, CR=1, GCR=2, BICGSTAB=3, BICGSTAB_L=4, GMRES=5, SOR=6, MINRES=7, INVALID=-1)

    def __init__(self, ctx, ops, transfers):
        assert len(ops) == len(transfers) + 1
        self.ctx, self.ops, self.transfers = ctx, list(ops), list(transfers)
        n = len(transfers)
        op_p = (C.c_void_p * (n + 1))(*[o.h.value if hasattr(o.h, "value") else o.h for o in ops])
        tr_p = (C.c_void_p * n)(*[t.h.value for t in transfers])
        self.h = ctx.ho.glbx_mg_create(n, op_p, tr_p)
        if not self.h:
            raise GlbError("glbx_mg_create failed")
        self.n_refine = n

    @classmethod
    def setup(cls, ctx, fine_op, X, Y, blocks, nvecs, bstrat=1, null_mass=1e-2, null_gen="BICGSTAB", tol=5e-5,
              max_iter=500, restart_freq=0, bicgstab_l=-1, do_ortho_eo=False, do_global_ortho_conj=False, seed=1337,
              verbosity=0, null_prec=0, do_free=False, links=None):
        """glbx_mg_setup: the reference driver's set-up sequence on the device (null_generate_random_smooth_dev,
        block_orthonormalize_dev, generate_coarse_from_fine_stencil_dev; aa_mg_square_staggered_u1.cpp:716-1143).
        fine_op: the level-0 stencil2d operator with the mass in its shift.  nvecs[l]: vectors of refinement l after
        the partition; bstrat 0 = BLOCK_NONE, 1 = BLOCK_EO, 2 = BLOCK_CORNER, 3 = BLOCK_TOPO (pass the host
        gauge field as `links`: the chiral projectors are built from its symmetric shifts); do_free: free-field vectors
        (null_generate_free_dev) instead of the smoothing solves; null_prec 0 = plain solves, 1 = even/odd (top/bottom below the
        top level), 2 = normal equations (null_gen.h:24-29)."""
        n = len(blocks)
        self = cls.__new__(cls)
        self.ctx, self.n_refine = ctx, n
        bl = (C.c_int * n)(*blocks)
        nv = (C.c_int * n)(*nvecs)
        tl = (C.c_double * n)(*([tol] * n if np.isscalar(tol) else tol))
        mi = (C.c_int * n)(*([max_iter] * n if np.isscalar(max_iter) else max_iter))
        self.h = ctx.ho.glbx_mg_setup(fine_op.h, X, Y, n, bl, nv, bstrat, null_mass, cls.SMOOTH[null_gen], tl, mi,
                                      restart_freq, bicgstab_l, int(do_ortho_eo), int(do_global_ortho_conj), seed,
                                      verbosity, null_prec, int(do_free), _p(links) if links is not None else None)
        if not self.h:
            raise GlbError("glbx_mg_setup failed (see stderr)")
        self.ops, self.transfers = [fine_op], []
        self.blocks, self.nvecs, self.top = list(blocks), list(nvecs), (X, Y)
        return self

    def level_dims(self, level):
        X, Y = self.top
        dof = 1
        for l in range(level):
            X, Y, dof = X // self.blocks[l], Y // self.blocks[l], self.nvecs[l]
        return X, Y, dof

    def _local_dims(self, level):
        """(X, Yloc, dof) of this rank's slab of level `level` (the whole lattice on one rank)"""
        X, Y, dof = self.level_dims(level)
        y0, yloc = self.ctx.slab_bounds(Y)
        return X, yloc, dof

    def level_stencil(self, level):
        """(clover, hopping, shifts) of level `level` of a hierarchy built by setup(); on y-slabs this rank's rows of
        every plane"""
        X, Y, nc = self._local_dims(level)
        h = self.ctx.ho.glbx_mg_level_op(self.h, level)
        cl = np.empty(X * Y * nc * nc, dtype=np.complex128)
        hp = np.empty(4 * X * Y * nc * nc, dtype=np.complex128)
        _chk(self.ctx.cu.glb_op_stencil_download(h, _p(cl), _p(hp)), "glb_op_stencil_download")
        a = [(C.c_double * 2)() for _ in range(3)]
        _chk(self.ctx.cu.glb_op_get_shifts(h, a[0], a[1], a[2]), "glb_op_get_shifts")
        return cl, hp, tuple(complex(v[0], v[1]) for v in a)

    def null_vector(self, level, v):
        """block-orthonormalised null vector v of refinement `level` (downloaded; on y-slabs this rank's rows)"""
        X, Y, dof = self._local_dims(level)
        out = np.empty(X * Y * dof, dtype=np.complex128)
        p = self.ctx.ho.glbx_mg_null_vector(self.h, level, v)
        if not p:
            raise GlbError("no such null vector")
        _chk(self.ctx.cu.glb_vec_download(self.ctx.h, COMPLEX, out.size, _p(out), p), "glb_vec_download")
        return out

    def setup_seconds(self):
        a = (C.c_double * 4)()
        self.ctx.ho.glbx_mg_setup_seconds(self.h, a)
        return dict(null_vectors=a[0], block_orthonormalize=a[1], galerkin=a[2], total=a[3])

    def set(self, smooth="GCR", n_pre=6, n_post=6, inner="GCR", n_max=1024, n_restart=64, rel_res=1e-2,
            recursive=False, quiet=True):
        self.ctx.ho.glbx_mg_set(self.h, self.SMOOTH[smooth], n_pre, n_post, self.INNER[inner], n_max, n_restart,
                                rel_res, 1 if recursive else 0, 1 if quiet else 0)

    def vcycle(self, out, rhs):
        _chk(self.ctx.ho.glbx_mg_vcycle(self.h, out.ptr, rhs.ptr), "glbx_mg_vcycle")

    def vpgcr(self, x, b, max_iter=1000, eps=5e-7, restart_freq=64, verbosity=0):
        res = Result()
        _chk(self.ctx.ho.glbx_mg_vpgcr(self.h, x.ptr, b.ptr, max_iter, eps, restart_freq, verbosity, C.byref(res)),
             "glbx_mg_vpgcr")
        return res.as_dict()

    def counts(self):
        n = self.n_refine + 1
        buf = (C.c_int * (5 * n))()
        self.ctx.ho.glbx_mg_counts(self.h, buf)
        names = ("krylov", "presmooth", "postsmooth", "residual", "nullvectors")
        return {nm: [buf[i * n + l] for l in range(n)] for i, nm in enumerate(names)}

    def destroy(self):
        if self.h:
            self.ctx.ho.glbx_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Context:
    """One GPU context = one process's view of the library (one per rank)."""

    def __init__(self, device=None, use_default=True):
        self.cu, self.ho = libs()
        if use_default:
            # share the C++ layer's process-wide context so host-pointer and device calls agree
            if device is not None:
                os.environ.setdefault("GLB200_DEVICE", str(device))
            h = self.ho.glbx_default_context()
            if not h:
                raise GlbError("no CUDA device / context: " + self.cu.glb_last_error().decode())
            self.h, self.owned = h, False
        else:
            p = C.c_void_p()
            _chk(self.cu.glb_create(int(device or 0), C.byref(p)), "glb_create")
            self.h, self.owned = p.value, True

    # ---- communicator bootstrap through torch.distributed (any launcher would do) ----
    def init_comm_from_torch(self):
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = C.create_string_buffer(128)
        if rank == 0:
            _chk(self.cu.glb_comm_unique_id(buf), "glb_comm_unique_id")
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        _chk(self.cu.glb_comm_init(self.h, rank, world, obj[0]), "glb_comm_init")
        return rank, world

    @property
    def rank(self):
        return self.cu.glb_comm_rank(self.h)

    @property
    def nranks(self):
        return self.cu.glb_comm_size(self.h)

    @property
    def p2p(self):
        return bool(self.cu.glb_comm_p2p_enabled(self.h))

    def barrier(self):
        _chk(self.cu.glb_comm_barrier(self.h), "glb_comm_barrier")

    def sync(self):
        _chk(self.cu.glb_synchronize(self.h), "glb_synchronize")

    def stream(self):
        return self.cu.glb_stream(self.h)

    def launches(self):
        return int(self.cu.glb_kernel_launches())

    def prof_enable(self, on=True):
        """per-kernel CUDA-event timing of the classified kernels (include/glb200.h: glb_prof_*)"""
        _chk(self.cu.glb_prof_enable(self.h, 1 if on else 0), "glb_prof_enable")

    def prof_read(self, cls):
        """durations (ms) of every launch of kernel class `cls` recorded since prof_enable, in launch order"""
        n = C.c_int(0)
        _chk(self.cu.glb_prof_read(self.h, cls, 0, None, C.byref(n)), "glb_prof_read")
        buf = (C.c_float * max(n.value, 1))()
        _chk(self.cu.glb_prof_read(self.h, cls, n.value, buf, C.byref(n)), "glb_prof_read")
        return [float(buf[i]) for i in range(n.value)]

    PROF_CLASSES = {1: "normal_kernel<fused>", 2: "normal_kernel / normal1_kernel (D^dag D, one pass)", 3: "cg_update_kernel",
                    4: "stag_kernel (staggered / gauged Laplace apply)", 5: "coarse stencil kernels (apply_stencil_2d)",
                    6: "laplace_kernel", 7: "cg_step_kernel", 8: "streaming BLAS-1 kernels (ew_kernel / ews_kernel)",
                    9: "multi_dot_kernel (batched <Ap_i, Ar>)", 10: "lincomb_kernel (p, Ap = z + sum beta_i p_i)",
                    11: "mg_prolong / mg_restrict"}

    def prof_summary(self):
        """{class name: (launches, total ms, algorithmic bytes)} of everything recorded since prof_enable"""
        out = {}
        for cls, name in self.PROF_CLASSES.items():
            n, ms, by = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
            _chk(self.cu.glb_prof_summary(self.h, cls, C.byref(n), C.byref(ms), C.byref(by)), "glb_prof_summary")
            if n.value:
                out[name] = (n.value, ms.value, by.value)
        return out

    def vector(self, n, dtype=np.complex128):
        return DeviceVector(self, n, dtype)

    def pinned(self, n, dtype=np.complex128):
        """numpy view of pinned host memory (for end-to-end timing with host buffers)"""
        dtype = np.dtype(dtype)
        p = C.c_void_p()
        _chk(self.cu.glb_host_alloc(self.h, n * dtype.itemsize, C.byref(p)), "glb_host_alloc")
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_byte)), shape=(n * dtype.itemsize,)).view(dtype)
        return arr

    # ---- operators ----
    def _op(self, fn, *args, keep=()):
        p = C.c_void_p()
        _chk(fn(self.h, *args, C.byref(p)), fn.__name__)
        return Operator(self, p.value, keep)

    def laplace(self, X, Y, Nc=1, diag=4.01, dtype=np.float64):
        diag = complex(diag)
        return self._op(self.cu.glb_op_create_laplace, _dt(dtype), X, Y, Nc, diag.real, diag.imag)

    def staggered_free_real(self, X, Y, mass):
        return self._op(self.cu.glb_op_create_staggered_free_real, X, Y, mass)

    def laplace_u1(self, links, X, Y, mass):
        links = np.ascontiguousarray(links, dtype=np.complex128)
        return self._op(self.cu.glb_op_create_laplace_u1, _p(links), X, Y, mass)

    def staggered(self, links, X, Y, mass, flags=0):
        if links is not None:
            links = np.ascontiguousarray(links, dtype=np.complex128)
        return self._op(self.cu.glb_op_create_staggered, _p(links), X, Y, mass, flags)

    def staggered_local(self, links_local, X, Y, mass, flags=0):
        """links_local: rows y0-2 .. y0+Yloc+1 of the gauge field (this rank's slab + two periodic ghost rows
        on each side), reference layout [row][x][mu]"""
        links_local = np.ascontiguousarray(links_local, dtype=np.complex128)
        return self._op(self.cu.glb_op_create_staggered_local, _p(links_local), X, Y, mass, flags)

    def slab_bounds(self, Y):
        y0, yl = C.c_int(), C.c_int()
        _chk(self.cu.glb_slab_bounds(self.h, Y, C.byref(y0), C.byref(yl)), "glb_slab_bounds")
        return y0.value, yl.value

    def gamma5(self, X, Y):
        return self._op(self.cu.glb_op_create_gamma5, X, Y)

    def stencil2d(self, clover, hopping, two_link, X, Y, nc, shift=0j, eo_shift=0j, dof_shift=0j):
        arrs = [np.ascontiguousarray(a, dtype=np.complex128) if a is not None else None
                for a in (clover, hopping, two_link)]
        return self._op(self.cu.glb_op_create_stencil2d, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), X, Y, nc,
                        _c2(shift), _c2(eo_shift), _c2(dof_shift))

    def mg_transfer(self, Xf, Yf, dof_f, bx, by, null_vectors):
        return MgTransfer(self, Xf, Yf, dof_f, bx, by, null_vectors)

    def multigrid(self, ops, transfers):
        return Multigrid(self, ops, transfers)

    def multigrid_setup(self, fine_op, X, Y, blocks, nvecs, **kw):
        return Multigrid.setup(self, fine_op, X, Y, blocks, nvecs, **kw)

    def mg_block_orthonormalize(self, Xf, Yf, dof_f, bx, by, dev_vectors):
        """block_orthonormalize + block_normalize (mg_complex.cpp:191-370) in place on DeviceVectors"""
        n = len(dev_vectors)
        ptrs = (C.c_void_p * n)(*[v.ptr for v in dev_vectors])
        _chk(self.cu.glb_mg_block_orthonormalize(self.h, Xf, Yf, dof_f, bx, by, n, ptrs), "glb_mg_block_orthonormalize")

    def mg_partition_corner(self, X, Y, dof, colour_period, which, src_io, dst_out):
        """BLOCK_CORNER partition, class `which` in 1..3 (null_gen.cpp:74-88 / :132-152)"""
        _chk(self.cu.glb_mg_partition_corner(self.h, X, Y, dof, int(colour_period), int(which), src_io.ptr, dst_out.ptr),
             "glb_mg_partition_corner")

    def mg_partition(self, X, Y, dof, colour_period, even_io, odd_out):
        """BLOCK_EO partition of one null vector: colour_period 0 = odd sites (null_gen.cpp:26-35), m > 0 = elements
        with index % m >= m/2 (null_gen.cpp:109-126)"""
        _chk(self.cu.glb_mg_partition(self.h, X, Y, dof, int(colour_period), even_io.ptr, odd_out.ptr), "glb_mg_partition")

    # ---- BLAS-1 (thin; used by the parity tests) ----
    def dot(self, x, y):
        o = (C.c_double * 2)()
        _chk(self.cu.glb_dot(self.h, x.dt, x.n, x.ptr, y.ptr, o), "glb_dot")
        return complex(o[0], o[1]) if x.dt == COMPLEX else o[0]

    def norm2sq(self, x):
        o = C.c_double()
        _chk(self.cu.glb_norm2sq(self.h, x.dt, x.n, x.ptr, C.byref(o)), "glb_norm2sq")
        return o.value

    def diffnorm2sq(self, x, y):
        o = C.c_double()
        _chk(self.cu.glb_diffnorm2sq(self.h, x.dt, x.n, x.ptr, y.ptr, C.byref(o)), "glb_diffnorm2sq")
        return o.value

    # ---- solvers on device vectors (device variant of the callback contract) ----
    def solve(self, solver, op, x, b, max_iter=10000, eps=1e-10, restart_freq=0, l=0, verbosity=0):
        res = Result()
        s = SOLVER[solver] if isinstance(solver, str) else solver
        _chk(self.ho.glbx_dev_solve(s, op.h, x.ptr, b.ptr, max_iter, eps, restart_freq, l, verbosity,
                                    C.byref(res)), "glbx_dev_solve")
        return res.as_dict()

    RELAX = dict(SOR=0, MINRES=1)

    def solve_relax(self, which, op, x, b, max_iter=10000, eps=1e-10, omega=1.0, verbosity=0):
        """minv_vector_sor_dev / minv_vector_minres_dev (generic_sor.cpp, generic_minres.cpp) on device vectors"""
        res = Result()
        _chk(self.ho.glbx_dev_solve_relax(self.RELAX[which], op.h, x.ptr, b.ptr, max_iter, eps, omega, verbosity,
                                          C.byref(res)), "glbx_dev_solve_relax")
        return res.as_dict()

    def host_stencil_prec(self, desc, top_bottom, a, b=None):
        """the reference-named prepare (b is None) / reconstruct wrappers with host vectors"""
        out = np.empty_like(a)
        _chk(self.ho.glbx_host_stencil_prec(C.byref(desc), int(top_bottom), 0 if b is None else 1, _p(out), _p(a),
                                            _p(b) if b is not None else None), "glbx_host_stencil_prec")
        return out

    def host_solve_relax(self, which, desc, x, b, max_iter=10000, eps=1e-10, omega=1.0, verbosity=0):
        """minv_vector_sor / minv_vector_minres with host vectors (x in/out)"""
        res = Result()
        _chk(self.ho.glbx_host_solve_relax(self.RELAX[which], C.byref(desc), _p(x), _p(b), max_iter, eps, omega,
                                           verbosity, C.byref(res)), "glbx_host_solve_relax")
        return res.as_dict()

    def solve_cg_m(self, op, xs, b, shifts, resid_freq_check=10, max_iter=10000, eps=1e-10, worst_first=False,
                   verbosity=0):
        n = len(xs)
        shifts = np.array(shifts, dtype=np.float64, copy=True)
        ptrs = (C.c_void_p * n)(*[x.ptr for x in xs])
        res = Result()
        _chk(self.ho.glbx_dev_solve_cg_m(op.h, ptrs, b.ptr, n, resid_freq_check, max_iter, eps, _p(shifts),
                                         int(worst_first), verbosity, C.byref(res)), "glbx_dev_solve_cg_m")
        return res.as_dict(), shifts

    def cg_device(self, op, x, b, max_iter=10000, eps=1e-10, want_history=False):
        """glb_cg_solve: the device-resident CG loop (no final true-residual apply)."""
        rep = CgReport()
        hist = np.zeros(max_iter if want_history else 0)
        _chk(self.cu.glb_cg_solve(op.h, x.ptr, b.ptr, max_iter, eps, C.byref(rep),
                                  hist.ctypes.data_as(C.POINTER(C.c_double)) if want_history else None,
                                  max_iter if want_history else 0), "glb_cg_solve")
        out = dict(iterations=rep.iterations, ops=rep.ops, hit_max_iter=bool(rep.hit_max_iter), rsq=rep.rsq,
                   bnorm=rep.bnorm)
        if want_history:
            out["history"] = hist[:rep.iterations]
        return out

    KRYLOV = dict(BICGSTAB=1, CR=2)

    def krylov_device(self, alg, op, x, b, max_iter=10000, eps=1e-10, want_history=False):
        """glb_krylov_solve: the device-resident BiCGStab / CR loop (no final true-residual apply)."""
        rep = CgReport()
        hist = np.zeros(max_iter if want_history else 0)
        _chk(self.cu.glb_krylov_solve(op.h, self.KRYLOV[alg], x.ptr, b.ptr, max_iter, eps, C.byref(rep),
                                      hist.ctypes.data_as(C.POINTER(C.c_double)) if want_history else None,
                                      max_iter if want_history else 0), "glb_krylov_solve")
        out = dict(iterations=rep.iterations, ops=rep.ops, hit_max_iter=bool(rep.hit_max_iter), rsq=rep.rsq,
                   bnorm=rep.bnorm, used_graph=bool(self.cu.glb_krylov_last_used_graph()))
        if want_history:
            out["history"] = hist[:rep.iterations]
        return out

    def krylov_supported(self, alg, op):
        return bool(self.cu.glb_krylov_solve_supported(op.h, self.KRYLOV[alg]))

    def krylov_graph_mode(self, on):
        """batches of the device-resident BiCGStab / CR loop as CUDA graphs (True, default) or direct launches"""
        return bool(self.cu.glb_krylov_graph_mode(1 if on else 0))

    def synthetic_inputs(self, X, Y, seed=1337, beta=6.0, want_rhs=True):
        """BASELINE.md section 3 inputs from the drop-in's host helpers: mt19937(seed) -> gauss_gauge_u1 -> gaussian"""
        links = np.empty(2 * X * Y, dtype=np.complex128)
        rhs = np.empty(X * Y, dtype=np.complex128) if want_rhs else None
        _chk(self.ho.glbx_synthetic_inputs(seed, X, Y, beta, _p(links), _p(rhs) if want_rhs else None),
             "glbx_synthetic_inputs")
        return links, rhs

    def cg_step_mode(self, on=True, variant=0):
        """glb_cg_step_mode: single-kernel CG iteration on/off (+ kernel shape); returns the previous setting"""
        return bool(self.cu.glb_cg_step_mode(1 if on else 0, int(variant)))

    def cg_last_pred_err(self):
        return float(self.cu.glb_cg_last_pred_err())

    # ---- the reference's own calls: HOST vectors + reference-named callbacks ----
    def _desc(self, kind, X, Y, mass=0.0, Nc=1, links=None, clover=None, hopping=None, two_link=None, shift=0j,
              eo_shift=0j, dof_shift=0j, view=0, wilson_coeff=0.0):
        d = OpDesc()
        d.wilson_coeff = wilson_coeff
        d.view = STENCIL_VIEW[view] if isinstance(view, str) else view
        d.kind = OP[kind] if isinstance(kind, str) else kind
        d.X, d.Y, d.Nc, d.mass = X, Y, Nc, mass
        d.links, d.clover, d.hopping, d.two_link = _p(links), _p(clover), _p(hopping), _p(two_link)
        d.has_two = 1 if two_link is not None else 0
        for name, v in (("shift", shift), ("eo_shift", eo_shift), ("dof_shift", dof_shift)):
            getattr(d, name)[0], getattr(d, name)[1] = complex(v).real, complex(v).imag
        d._keep = (links, clover, hopping, two_link)
        return d

    def host_apply(self, desc, rhs):
        out = np.empty_like(rhs)
        _chk(self.ho.glbx_host_apply(C.byref(desc), _p(out), _p(rhs)), "glbx_host_apply")
        return out

    def host_stencil_part(self, desc, part, rhs):
        out = np.empty_like(rhs)
        _chk(self.ho.glbx_host_stencil_part(C.byref(desc), Operator.PART[part], _p(out), _p(rhs)), "glbx_host_stencil_part")
        return out

    def host_eoprec_prepare(self, desc, rhs_orig):
        out = np.empty_like(rhs_orig)
        _chk(self.ho.glbx_host_eoprec_prepare(C.byref(desc), _p(out), _p(rhs_orig)), "glbx_host_eoprec_prepare")
        return out

    def host_eoprec_reconstruct(self, desc, lhs_e, rhs_o):
        out = np.empty_like(lhs_e)
        _chk(self.ho.glbx_host_eoprec_reconstruct(C.byref(desc), _p(out), _p(lhs_e), _p(rhs_o)),
             "glbx_host_eoprec_reconstruct")
        return out

    def host_solve(self, solver, desc, x, b, max_iter=10000, eps=1e-10, restart_freq=0, l=0, verbosity=0):
        """x (numpy, in/out) and b (numpy) are HOST arrays: copies are part of the call."""
        res = Result()
        s = SOLVER[solver] if isinstance(solver, str) else solver
        _chk(self.ho.glbx_host_solve(s, C.byref(desc), _p(x), _p(b), max_iter, eps, restart_freq, l, verbosity,
                                     C.byref(res)), "glbx_host_solve")
        return res.as_dict()

    def host_solve_cg_m(self, desc, xs, b, shifts, resid_freq_check=10, max_iter=10000, eps=1e-10,
                        worst_first=False, verbosity=0):
        n = len(xs)
        shifts = np.array(shifts, dtype=np.float64, copy=True)
        ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        res = Result()
        _chk(self.ho.glbx_host_solve_cg_m(C.byref(desc), ptrs, _p(b), n, resid_freq_check, max_iter, eps,
                                          _p(shifts), int(worst_first), verbosity, C.byref(res)),
             "glbx_host_solve_cg_m")
        return res.as_dict(), shifts

    MULTI = dict(CG_M=0, CR_M=1, BICGSTAB_M=2)
    PRECOND_SOLVER = dict(PCG=0, FPCG=1, FPCG_RESTART=2, VPGCR=3, VPGCR_RESTART=4, PBICGSTAB=5, PBICGSTAB_RESTART=6)
    PRECOND = dict(IDENTITY=0, GCR=1, MINRES=2)

    def host_solve_multi(self, which, desc, xs, b, shifts, resid_freq_check=10, max_iter=10000, eps=1e-10,
                         worst_first=False, verbosity=0):
        """minv_vector_{cg,cr,bicgstab}_m with HOST vectors (generic_cg_m.h, generic_cr_m.h, generic_bicgstab_m.h)"""
        n = len(xs)
        shifts = np.array(shifts, dtype=np.float64, copy=True)
        ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        res = Result()
        _chk(self.ho.glbx_host_solve_multi(self.MULTI[which], C.byref(desc), ptrs, _p(b), n, resid_freq_check, max_iter,
                                           eps, _p(shifts), int(worst_first), verbosity, C.byref(res)),
             "glbx_host_solve_multi")
        return res.as_dict(), shifts

    def host_solve_precond(self, solver, desc, x, b, max_iter=10000, eps=1e-10, restart_freq=0, precond="IDENTITY",
                           n_step=4, rel_res=1e-20, verbosity=0):
        """the preconditioned family (generic_inverters_precond.h) with HOST vectors and the stock preconditioners
        of generic_precond.h: identity_preconditioner, or gcr_preconditioner (n_step iterations on the same operator)"""
        res = Result()
        _chk(self.ho.glbx_host_solve_precond(self.PRECOND_SOLVER[solver], C.byref(desc), _p(x), _p(b), max_iter, eps,
                                             restart_freq, self.PRECOND[precond], n_step, rel_res, verbosity,
                                             C.byref(res)), "glbx_host_solve_precond")
        return res.as_dict()

    def force_host_scalars(self, on):
        self.ho.glbx_force_host_scalars(int(on))

    def cache_operators(self, on):
        self.ho.glbx_cache_operators(int(on))

    def close(self):
        if self.owned and self.h:
            self.cu.glb_destroy(self.h)
            self.h = None
