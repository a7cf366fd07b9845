"""Host-side slab partition / ghost-zone / halo-exchange logic (libgeodecomp_b200.striping) on CPU:
world_size 2 and 3 over gloo, the device engine replaced by tests/cpu_engine.py. Results must equal
the single-domain oracle for every ghost-zone width (as parallel_mpi_4/hiparsimulatortest.h checks
ghost widths 1..6 for the reference)."""
import os
import socket
import subprocess
import sys

import pytest

from libgeodecomp_b200.striping import slab_bounds

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_slab_bounds_are_the_striping_partition():
    assert slab_bounds(1024, 8) == [0, 128, 256, 384, 512, 640, 768, 896, 1024]
    assert slab_bounds(10, 3) == [0, 3, 6, 10]
    assert slab_bounds(5, 1) == [0, 5]


@pytest.mark.parametrize("world", [2, 3])
def test_striped_simulator_matches_single_domain(tmp_path, world):
    port = free_port()
    out = str(tmp_path / "result")
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_striping_worker.py"), out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    for rank, p in enumerate(procs):
        assert p.returncode == 0, logs[rank][-3000:]
        content = open("%s.%d" % (out, rank)).read()
        assert content == "OK", content
