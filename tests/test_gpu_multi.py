"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on a single-GPU box): launches tests/mgpu_check.py under
torch.distributed.run with a T-split and a Z-split processor grid."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("Ls", [6, 8])  # Ls = 8: fp32 Dslash through the TMA sweep kernel (dslash_tma.cu, COMM instance)
@pytest.mark.parametrize("mpi", ["1.1.1.2", "1.1.2.1"])
def test_two_gpu_parity(mpi, Ls):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_check.py"), "--mpi", mpi, "--Ls", str(Ls),
           # Ls = 6: operators incl. the clover term built from the all-gathered links; Ls = 8: TMA sweep kernel, CG iteration
           # counts, the mixed-precision Wilson-clover solve (BASELINE configs[3] in small)
           "--only", "mobius,clover" if Ls == 6 else "mobius,cg,solve"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
