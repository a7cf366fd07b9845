"""Slab partition + NCCL halo exchange on real GPUs (needs >= 2 devices; skipped otherwise):
results must be bit-identical to the single-domain oracle for every ghost-zone width, for Cube and
Torus, for 3-D (slabs along z) and 2-D (slabs along y) grids."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_striped_simulator_nccl(tmp_path, world):
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "result")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "_striping_gpu_worker.py"), out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    for rank in range(world):
        content = open("%s.%d" % (out, rank)).read()
        assert content == "OK", content
