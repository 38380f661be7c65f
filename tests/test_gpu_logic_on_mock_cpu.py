"""The Python side of some GPU test modules (multigrid set-up, preconditioned stencil paths) run against the CPU mock of the C ABI
in a child process (tests/mock/run_gpu_tests_on_mock.py): bindings, argument orders, shapes and assertions of the
`-m gpu` tests are exercised here, where no GPU exists, so that GPU minutes are spent on the kernels only.  Says
nothing about the CUDA code."""
import os
import subprocess
import sys

import pytest

import oracle_py
from conftest import ROOT

pytestmark = pytest.mark.skipif("ref" not in oracle_py.available(),
                                reason="the reference's multigrid is only in oracle/_ref/libref_oracle.so")


@pytest.mark.parametrize("module", ["test_mg_setup_gpu", "test_stencil_prec_gpu", "test_operators_rest_gpu"])
def test_gpu_test_logic_runs_on_the_mock(module):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock", "run_gpu_tests_on_mock.py"), module],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith(("PASS", "FAIL", "SKIP"))]
    # SKIP lines are tests that opt out of the mock themselves (argument checking is the CUDA library's)
    assert r.returncode == 0 and [l for l in lines if l.startswith("PASS")] and not [l for l in lines if l.startswith("FAIL")], \
        r.stdout[-3000:]
