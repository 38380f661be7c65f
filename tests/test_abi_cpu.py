"""No-GPU checks of the drop-in boundary: the native libraries exist in-tree, load, export every
symbol include/glb200.h declares, and refuse to run without a device (no silent CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT, load_pkg

PKG = os.path.join(ROOT, "generic-linalg_b200")


def _built():
    return os.path.exists(os.path.join(PKG, "libglb200.so")) and os.path.exists(
        os.path.join(PKG, "libglb200_inverters.so"))


needs_build = pytest.mark.skipif(not _built(), reason="native libraries not built (run __graft_entry__.build())")


@needs_build
def test_every_declared_symbol_is_exported():
    glb = load_pkg()
    lib = C.CDLL(os.path.join(PKG, "libglb200.so"))
    names = glb.exported_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


@needs_build
def test_dropin_cxx_symbols_present():
    """the reference's C++ entry points exist with the reference's (mangled) signatures"""
    out = subprocess.check_output(["nm", "-DC", os.path.join(PKG, "libglb200_inverters.so")]).decode()
    for sym in ["minv_vector_cg(std::complex<double>*, std::complex<double>*, int, int, double, "
                "void (*)(std::complex<double>*, std::complex<double>*, void*), void*, inversion_verbose_struct*)",
                "minv_vector_cg(double*, double*, int, int, double, void (*)(double*, double*, void*), void*, "
                "inversion_verbose_struct*)",
                "minv_vector_bicgstab_l(std::complex<double>*, std::complex<double>*, int, int, double, int,",
                "minv_vector_gmres_restart(std::complex<double>*, std::complex<double>*, int, int, double, int,",
                "minv_vector_cg_m(std::complex<double>**, std::complex<double>*, int, int, int, int, double, double*,",
                "minv_unpreconditioned(std::complex<double>*, std::complex<double>*, int, minv_inverter, "
                "minv_inverter_params&",
                "square_staggered_u1(std::complex<double>*, std::complex<double>*, void*)",
                "square_staggered_normal_u1(std::complex<double>*, std::complex<double>*, void*)",
                "apply_stencil_2d(std::complex<double>*, std::complex<double>*, void*)",
                "gaussian_elimination(std::complex<double>*, std::complex<double>*, std::complex<double>**, int)"]:
        assert sym in out, sym


@needs_build
def test_no_oracle_in_product():
    """the product libraries must not link or reference anything under oracle/"""
    for lib in ("libglb200.so", "libglb200_inverters.so"):
        ldd = subprocess.check_output(["ldd", os.path.join(PKG, lib)]).decode()
        assert "oracle" not in ldd
    for dirpath, _, files in os.walk(PKG):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".py")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"#include\s+[\"<].*oracle", txt), f
                assert "oracle_py" not in txt and "libport_oracle" not in txt and "libref_oracle" not in txt, f


@needs_build
def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    glb = load_pkg()
    cu, _ = glb.libs()
    h = C.c_void_p()
    rc = cu.glb_create(0, C.byref(h))
    assert rc != 0
    assert b"no CUDA device" in cu.glb_last_error() or b"CPU fallback" in cu.glb_last_error()
    # in a fresh interpreter: other test modules load the host-memory mock of the C ABI with RTLD_GLOBAL, whose symbols
    # would answer for libglb200.so's in this process
    code = ("import sys; sys.path.insert(0, %r); from __graft_entry__ import _load_pkg; glb = _load_pkg()\n"
            "try:\n    glb.Context()\nexcept glb.GlbError as e:\n    print('LOUD', e)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "LOUD" in r.stdout, r.stdout + r.stderr


def test_missing_library_is_an_error(tmp_path, monkeypatch):
    glb = load_pkg()
    monkeypatch.setattr(glb, "_libs", None)
    monkeypatch.setattr(glb, "LIB_CUDA", str(tmp_path / "nope.so"))
    with pytest.raises(glb.GlbError):
        glb.libs()


def test_header_is_plain_c(tmp_path):
    """include/glb200.h is what a cgo / JNI / ctypes host binds: it must compile as C99, no C++ types in it"""
    src = tmp_path / "t.c"
    src.write_text('#include "glb200.h"\nint main(void) { return GLB_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-fsyntax-only", str(src)])


@needs_build
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree")
def test_every_reference_solver_and_operator_entry_point_has_a_drop_in():
    """every function the reference declares in its solver / operator / stencil headers on the path is exported by
    libglb200_inverters.so with the same C++ signature (nm -DC of both sides, compared by demangled prototype)"""
    ref = "/root/reference"
    headers = ["generic_cg.h", "generic_cr.h", "generic_gcr.h", "generic_bicgstab.h", "generic_bicgstab_l.h",
               "generic_gmres.h", "generic_cg_m.h", "generic_cr_m.h", "generic_bicgstab_m.h", "generic_sor.h",
               "generic_minres.h", "generic_cg_precond.h", "generic_cg_flex_precond.h", "generic_gcr_var_precond.h",
               "generic_bicgstab_precond.h", "generic_inverters.h", "generic_inverters_precond.h", "generic_gelim.h",
               "operator_utils/operators.h", "operator_utils/operators_stencil.h", "stencil_2d/coarse_stencil.h",
               "multigrid/aa_mg/mg_complex.h", "multigrid/aa_mg/null_gen.h", "u1_utils/u1_utils.h", "generic_eigenvalues.h",
               "generic_precond.h"]
    # clear_stencils is a member function of stencil_2d (inline in host/coarse_stencil.h)
    not_offered = {"clear_stencils"}
    names = set()
    for h in headers:
        txt = open(os.path.join(ref, h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"//[^\n]*", "", txt)
        names.update(re.findall(r"^\s*(?:inversion_info|eigenvalue_info|void|int|double|complex<double>)\s+(\w+)\s*\(", txt,
                                flags=re.M))
    names -= not_offered
    assert len(names) > 80
    out = subprocess.check_output(["nm", "-DC", "--defined-only", os.path.join(PKG, "libglb200_inverters.so")]).decode()
    exported = set(re.findall(r" T (\w+)\(", out))
    missing = sorted(n for n in names if n not in exported)
    assert not missing, missing


def test_library_never_uses_the_host_vector_helpers():
    """host/generic_vector.h and host/u1_utils.h exist for DRIVER programs (fill / check host arrays, gauge-field files).
    No translation unit of the two libraries includes generic_vector.h: all vector work of the path is device work."""
    for sub in ("host", "csrc"):
        for f in os.listdir(os.path.join(PKG, sub)):
            if f.endswith((".cpp", ".hpp", ".cu", ".cuh")) or (f.endswith(".h") and f != "generic_vector.h"):
                txt = open(os.path.join(PKG, sub, f)).read()
                assert not re.search(r"#include\s+[\"<]generic_vector\.h", txt), f
    out = subprocess.check_output(["nm", "-DC", "--defined-only", os.path.join(PKG, "libglb200_inverters.so")]).decode() \
        if _built() else ""
    assert "glb200_hostvec" not in out


def _build_c_example(tmp_path):
    exe = str(tmp_path / "c_abi_cgne")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_cgne.c"), "-L", PKG, "-lglb200", "-lm",
                           "-Wl,-rpath," + PKG, "-o", exe])
    return exe


@needs_build
def test_plain_c_program_links_against_the_c_abi_and_refuses_to_run_without_a_gpu(tmp_path):
    """examples/c_abi_cgne.c: what a foreign-language binding does, in C99; on this GPU-less machine it must stop at
    glb_create with the library's message -- no CPU path takes over"""
    exe = _build_c_example(tmp_path)
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    r = subprocess.run([exe, "32"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@needs_build
@pytest.mark.gpu
def test_plain_c_program_runs_on_the_gpu(tmp_path):
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, "64"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "CGNE on 64 x 64" in r.stdout and "iterations" in r.stdout
    resid = float(r.stdout.split("|Dx-b|/|b| = ")[1].split(",")[0])
    assert resid < 1e-8
