"""The C-ABI library loads without a GPU and exports every symbol include/b200geo.h declares."""
import ctypes
import os
import re

import pytest

from libgeodecomp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200geo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200geo_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(name for name, _, _ in capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), name


@pytest.mark.parametrize("compiler,flags", [("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++11", "-x", "c++"])])
def test_header_is_plain_c_and_cpp(tmp_path, compiler, flags):
    """the C ABI header has no C++ or torch types in it: it compiles as C99 and as C++11 on its own"""
    import shutil
    import subprocess
    if not shutil.which(compiler):
        pytest.skip(compiler + " not installed")
    src = tmp_path / "probe.c"
    src.write_text('#include "b200geo.h"\nint probe(void) { b200geo_grid_desc d; d.n_members = 1; return (int)sizeof(d) + B200GEO_OK; }\n')
    res = subprocess.run([compiler] + flags + ["-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_version_and_error_text():
    lib = capi.lib()
    assert b"b200geo" in lib.b200geo_version()
    assert isinstance(lib.b200geo_last_error(), bytes)
    assert capi.launch_count() >= 0


def test_argument_validation_needs_no_gpu():
    """invalid descriptors are rejected before any CUDA call (std::invalid_argument)."""
    with pytest.raises(ValueError):
        capi.DeviceGrid((0, 4, 4), [8])
    with pytest.raises(ValueError):
        capi.DeviceGrid((4, 4, 4), [3])
    with pytest.raises(ValueError):
        capi.DeviceGrid((4, 4, 4), [])
    with pytest.raises(capi.LogicError):
        capi.DeviceGrid((4, 4, 4), [8], ghost_mode=[[capi.GHOST_PEER, capi.GHOST_EDGE], [0, 0], [0, 0]])


def test_no_cpu_fallback():
    """without a CUDA device grid creation fails loudly (runtime_error "CUDA error")."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.CudaError):
        capi.DeviceGrid((8, 8, 8), [8])


def test_cpp_facade_fails_loudly_without_a_gpu():
    """tests/facade/nogpu_test.cpp (built where /root/reference exists): B200Grid / B200Simulator /
    B200StripingSimulator throw std::runtime_error("CUDA error ...") when there is no device; argument
    validation maps to std::invalid_argument / std::logic_error before any CUDA call"""
    import subprocess
    binary = os.path.join(ROOT, "tests", "facade", "_bin", "nogpu_test")
    if not os.access(binary, os.X_OK):
        pytest.skip("tests/facade/_bin/nogpu_test not built (needs /root/reference at build time)")
    res = subprocess.run([binary], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "libgeodecomp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in text and "liboracle" not in text and "oracle.h" not in text, f
