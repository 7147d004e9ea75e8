import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _make(target):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), target], check=True, capture_output=True)


@pytest.fixture(scope="session")
def oracle():
    """Our CPU restatement (oracle/pf_oracle.cpp) -- the checker."""
    from oracle.bindings import Checker
    if not os.path.exists(os.path.join(ROOT, "oracle", "libpforacle.so")):
        _make("oracle")
    return Checker("oracle")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (present when built in the dev container)."""
    from oracle.bindings import Checker
    path = os.path.join(ROOT, "oracle", "_ref", "libpfref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            _make("ref")
        else:
            pytest.skip("oracle/_ref/libpfref.so not built and /root/reference absent")
    return Checker("ref")


@pytest.fixture(scope="session")
def hostemu():
    """tests/hostemu: the product's alignment state machines compiled for the CPU (development aid)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostemu")
    so = os.path.join(d, "libpfemu.so")
    src = os.path.join(d, "align_emu.cpp")
    core = os.path.join(ROOT, "ploidyfrost_b200", "csrc", "pf_align_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-fPIC", "-std=c++14", "-pthread", "-shared", "-o", so, src], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="session")
def gpu_ctx():
    from ploidyfrost_b200 import capi
    ctx = capi.Context(0)   # raises if the library or the GPU is missing: no fallback
    yield ctx
    ctx.close()
