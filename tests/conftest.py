"""pytest configuration.

  -m "not gpu" : runs anywhere (no GPU): the oracle against the golden vectors, port-vs-reference
                 bit equality (when oracle/_ref is built), host-side logic, ABI surface of the
                 native libraries, and the slab decomposition under gloo (world_size 2).
  -m gpu       : the parity tests proper -- every one of them goes through the C ABI of
                 generic-linalg_b200/libglb200.so on a real B200 and compares with the oracle.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """import generic-linalg_b200/ (the hyphen keeps it from being a plain `import`)"""
    name = "generic_linalg_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "generic-linalg_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    """the strongest CPU checker available: the compiled reference if present, else the port"""
    import oracle_py
    if not oracle_py.available():
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return oracle_py.load("best")


@pytest.fixture(scope="session")
def port():
    import oracle_py
    if "port" not in oracle_py.available():
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return oracle_py.load("port")


@pytest.fixture(scope="session")
def glb():
    return load_pkg()


@pytest.fixture(scope="session")
def ctx(glb):
    """process-wide device context; fails loudly if the CUDA library or the GPU is missing"""
    return glb.Context()


def synthetic(orc, L, beta=6.0, seed=1337):
    """gauge field + rhs exactly as BASELINE.md section 3 prescribes (one mt19937 stream)"""
    r = orc.rng(seed)
    U = r.gauss_gauge_u1(L, L, beta)
    b = r.gaussian(L * L)
    return U, b


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
