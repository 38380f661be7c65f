#!/usr/bin/env python
"""Writes tests/golden/pieces_reference.txt: the output of tests/mock/pieces_driver.cpp built against the REFERENCE's
sources (run in the container that has /root/reference; the GPU box has not).  tests/test_zz_pieces_gpu.py builds the
same driver against generic-linalg_b200/host and the CUDA library and compares."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
SRCS = ("generic_cg.cpp generic_bicgstab_l.cpp u1_utils/u1_utils.cpp operator_utils/operators.cpp "
        "operator_utils/operators_stencil.cpp stencil_2d/coarse_stencil.cpp multigrid/aa_mg/null_gen.cpp "
        "multigrid/aa_mg/mg_complex.cpp generic_cr.cpp generic_bicgstab.cpp generic_gmres.cpp generic_gcr.cpp "
        "generic_minres.cpp generic_sor.cpp generic_gelim.cpp generic_inverter.cpp generic_cg_flex_precond.cpp "
        "generic_bicgstab_precond.cpp generic_gcr_var_precond.cpp").split()
CASES = [("16", "0.1"), ("32", "0.05")]


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "pieces_ref")
        inc = ["-I" + os.path.join(REF, x) for x in ("", "u1_utils", "operator_utils", "stencil_2d", "lattice", "multigrid/aa_mg")]
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-DPIECES_DECLARE_LATTICE_FUNCTIONS"] + inc +
                              [os.path.join(ROOT, "tests", "mock", "pieces_driver.cpp")] + [os.path.join(REF, s) for s in SRCS] +
                              ["-o", exe, "-lrt"], stderr=subprocess.DEVNULL)
        out = []
        for L, m in CASES:
            r = subprocess.run([exe, L, m], capture_output=True, text=True, check=True)
            out.append("# L %s mass %s" % (L, m))
            out += [l for l in r.stdout.splitlines() if l.startswith("T")]
    with open(os.path.join(ROOT, "tests", "golden", "pieces_reference.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", len(out), "lines")


if __name__ == "__main__":
    main()
