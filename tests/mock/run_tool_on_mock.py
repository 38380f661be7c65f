"""run a tool's main() against the CPU mock (dev aid)"""
import ctypes as C, os, sys, importlib
ROOT="/root/repo"
for p in (os.path.join(ROOT,"tests"), os.path.join(ROOT,"oracle"), os.path.join(ROOT,"tools"), ROOT):
    sys.path.insert(0,p)
import conftest
sys.path.insert(0, os.path.join(ROOT,"tests","mock"))
import run_gpu_tests_on_mock as R
glb = conftest.load_pkg()
real = C.CDLL
glb.C.CDLL = lambda path, mode=0: R._Tolerant(real(R.MOCK, mode=C.RTLD_LOCAL))
glb._libs = None
cu, ho = glb.libs()
glb.C.CDLL = real
name = sys.argv[1]
sys.argv = [name] + sys.argv[2:]
mod = importlib.import_module(name)
mod.main()
