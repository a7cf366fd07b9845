"""Runs the C++ drop-in test (tests/facade/facade_test.cpp): the reference's SerialSimulator and
B200Simulator side by side on the same user models, the reference's MockWriter/MockSteerer event
logs, and SoAGrid vs B200Grid region/member byte streams. The binary is built in the container
that has /root/reference (tests/facade/Makefile, also run by __graft_entry__.build()) and travels."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "facade", "_bin", "facade_test")


@pytest.mark.gpu
def test_cpp_facade_drop_in():
    if not os.access(BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/facade_test not built (needs /root/reference at build time)")
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


GENERIC_BIN = os.path.join(HERE, "facade", "_bin", "generic_test")


@pytest.mark.gpu
def test_generic_device_path_for_unbound_cells():
    """tests/facade/generic_test.cu: user cells without a kernel binding (the reference's TestCell suite of
    cudasimulatortest.h, a Coord<2>-addressed Game of Life, a partially assigned cell, an AoS updateLineX
    model), their own update() compiled by nvcc, bit-identical to SerialSimulator."""
    if not os.access(GENERIC_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/generic_test not built (needs /root/reference at build time)")
    res = subprocess.run([GENERIC_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


STRIPING_BIN = os.path.join(HERE, "facade", "_bin", "striping_test")


@pytest.mark.gpu
def test_cpp_striping_simulator_drop_in():
    """tests/facade/striping_test.cpp: B200StripingSimulator (one process, slabs round-robin over the GPUs
    present, direct device-to-device halos) bit-identical to SerialSimulator for 1-4 slabs."""
    if not os.access(STRIPING_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/striping_test not built (needs /root/reference at build time)")
    res = subprocess.run([STRIPING_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


def test_facade_header_has_no_oracle_dependency():
    for name in ("b200simulator.h", "b200generic.h", "b200boxgrid.h", "b200stripingsimulator.h"):
        text = open(os.path.join(HERE, "..", "include", "libgeodecomp_b200", name)).read()
        assert "oracle" not in text
