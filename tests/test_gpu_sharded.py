"""-m gpu, needs >= 2 GPUs: the sharded global-batch InfoNCE equals the single-GPU result (feature-level
op and the whole model step), with the default peer-memory collectives and with NCCL.  Launches
tools/sharded_check.py under torchrun."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "sharded_check.py"), "2048"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_sharded_nccl_collectives_equal_single_gpu():
    """same check with the NCCL collectives (CVCL_B200_SYMM=0) instead of the default peer-memory
    kernels (csrc/peer_collectives.cuh)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, CVCL_B200_SYMM="0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "sharded_check.py"), "2048"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout and "exchange=nccl" in res.stdout, \
        res.stdout[-2000:] + res.stderr[-4000:]
