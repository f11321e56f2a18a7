"""rin_exchange_nccl on real GPUs: world_size 2 (and 4 when the box has them), one process per GPU, merged mesh
against the CPU oracle (tests/nccl_worker.py).  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("variant", ["default", "tiny_capacity", "nccl_data_path"])
def test_device_side_exchange_against_the_oracle(world, variant):
    """default: peer-memory exchange kernels (CUDA IPC) behind the run; tiny_capacity: the message capacity starts at 8
    entries, so the collective renegotiation runs; nccl_data_path: ncclSend/Recv + ncclAllGather instead of the peer
    kernels (the fallback when IPC is unavailable)."""
    if _device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ)
    if variant == "tiny_capacity":
        env["RIN_XCAP"] = "8"
    if variant == "nccl_data_path":
        env["RIN_NO_PEER_EXCHANGE"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + world), os.path.join(HERE, "nccl_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("nccl parity ok") == 10, p.stdout[-3000:]
