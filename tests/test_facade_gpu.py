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


def test_facade_header_has_no_oracle_dependency():
    text = open(os.path.join(HERE, "..", "include", "libgeodecomp_b200", "b200simulator.h")).read()
    assert "oracle" not in text
