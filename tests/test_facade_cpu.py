"""Host logic of the C++ façade on the CPU: tests/facade/facade_test.cpp and striping_test.cpp — the very objects
the GPU suite runs — linked against tests/facade/mock_b200geo.cpp, a host-memory stand-in for libb200geo.so that
implements the C ABI on plain arrays and delegates the sweeps to the oracle. What this covers without a GPU:
B200Grid (combined writes, cached rows, region / member byte streams against the reference's SoAGrid, status code ->
exception mapping), B200StripedGrid / B200StripedBoxGrid (routing across slabs, region split, slab bookkeeping),
B200Simulator / B200StripingSimulator (event protocol against the reference's MockWriter / MockSteerer, step fusing,
Steerer writes), the reference's SerialBOVWriter on both — each case beside the reference's SerialSimulator in the same
process; generic_soa_host_test.cpp: the generic SoA path (b200genericsoa.h) with its kernel body run in a host loop —
member table derived from LibFlatArray's generated accessors, stride dispatch, hood arithmetic over EDGE / WRAP ghosts;
stepper_test.cpp: B200Stepper beside the reference's VanillaStepper as rank 1 of 3 of a ragged StripingPartition (the set-up
of parallelization/nesting/test/parallel_mpi_1/vanillastepperregiontest.h), ghost zone widths 1-4. The kernels and the halo schedule themselves are GPU-tested (tests/test_facade_gpu.py)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", ["facade_test_cpu", "striping_test_cpu", "host_logic_test_cpu", "generic_soa_host_test_cpu", "stepper_test_cpu", "checkpoint_test_cpu", "streamed_test_cpu", "container_test_cpu", "voronoi_test_cpu"])
def test_cpp_facade_host_logic_on_the_mock_engine(name):
    binary = os.path.join(HERE, "facade", "_bin", name)
    if not os.access(binary, os.X_OK):
        pytest.skip("tests/facade/_bin/%s not built (needs /root/reference at build time)" % name)
    res = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout
    assert "MISMATCH" not in res.stdout and "DIFFERENT" not in res.stdout


def test_cpp_e2e_driver_on_the_mock_engine():
    """tests/facade/e2e_bench.cpp (bench.py: "e2e_cpp"): whole boxes and row by row, one simulator and slabs — the same
    final grid in every mode"""
    import json
    binary = os.path.join(HERE, "facade", "_bin", "e2e_bench_cpu")
    if not os.access(binary, os.X_OK):
        pytest.skip("tests/facade/_bin/e2e_bench_cpu not built (needs /root/reference at build time)")
    sums = set()
    for argv in (["20", "3", "1", "box"], ["20", "3", "1", "rows"], ["20", "3", "3", "box"], ["20", "3", "2", "rows"]):
        res = subprocess.run([binary] + argv, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        d = json.loads(res.stdout.splitlines()[0])
        assert d["value"] > 0 and "error" not in d
        sums.add(d["checksum"])
    assert len(sums) == 1
