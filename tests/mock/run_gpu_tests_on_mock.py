#!/usr/bin/env python
"""TEST INFRASTRUCTURE (development aid): run the Python logic of a `-m gpu` test module against the host-memory
mock of the C ABI (tests/mock/libglb200_inverters_mock.so), in THIS process only.

    python tests/mock/run_gpu_tests_on_mock.py test_mg_setup_gpu [test_name ...]

The GPU box is a scarce resource; this catches binding mistakes, wrong argument orders, shape errors and broken
assertions in the GPU tests before they cost GPU minutes.  It proves nothing about the CUDA kernels -- the mock's
"device" is host memory and its operators are the CPU oracle's -- and nothing in the product can reach it: the
package's loader is patched here, from the outside, for the lifetime of this script."""
import ctypes as C
import importlib
import inspect
import os
import subprocess
import sys
import traceback

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), ROOT):
    sys.path.insert(0, p)
import conftest  # noqa: E402

MOCK = os.path.join(HERE, "libglb200_inverters_mock.so")


class _Tolerant:
    """a CDLL whose missing symbols raise only when CALLED (the mock does not implement the CUDA-only entry points)"""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        try:
            return getattr(self._lib, name)
        except AttributeError:
            class Missing:
                restype = None
                argtypes = None

                def __call__(self, *a):
                    raise RuntimeError("the CPU mock has no " + name)
            m = Missing()
            setattr(self, name, m)
            return m


def main():
    subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
    glb = conftest.load_pkg()
    real = C.CDLL
    glb.C.CDLL = lambda path, mode=0: _Tolerant(real(MOCK, mode=C.RTLD_LOCAL))
    glb._libs = None
    cu, ho = glb.libs()
    glb.C.CDLL = real
    ho.glbx_force_host_scalars(1)
    ctx = glb.Context()
    import oracle_py
    fixtures = dict(ctx=ctx, glb=glb, orc=oracle_py.load("best"))
    if "port" in oracle_py.available():
        fixtures["port"] = oracle_py.load("port")
    mod = importlib.import_module(sys.argv[1])
    only = sys.argv[2:]
    failed = 0
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith("test_") or (only and name not in only):
            continue
        cases = [dict()]
        for m in [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]:
            names = [n.strip() for n in m.args[0].split(",")]
            cases = [dict(c, **dict(zip(names, vals if len(names) > 1 else (vals,)))) for c in cases for vals in m.args[1]]
        params = inspect.signature(fn).parameters
        for c in cases:
            kw = dict(c)
            missing = [p for p in params if p not in kw and p not in fixtures]
            if missing:
                print("SKIP", name, "needs fixture", missing)
                continue
            kw.update({p: fixtures[p] for p in params if p in fixtures})
            try:
                fn(**kw)
                print("PASS", name, c)
            except pytest.skip.Exception as e:     # e.g. argument-error tests: the mock does no argument checking
                print("SKIP", name, c, str(e)[:120])
            except Exception as e:
                traceback.print_exc()
                print("FAIL", name, c, repr(e)[:200])
                failed += 1
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
