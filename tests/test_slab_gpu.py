"""Multi-GPU slab parity (needs >= 2 GPUs; skipped on a single-GPU box): launches tools/slab_check.py
under torchrun with one rank per GPU and requires every check to pass against the CPU oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_path_matches_oracle(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "slab_check.py"),
           "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(out.stdout[-4000:])
    print(out.stderr[-2000:])
    assert out.returncode == 0 and "SLAB CHECK PASSED" in out.stdout
