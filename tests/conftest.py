import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the native libraries once per session if they are missing (CPU-only is fine:
    nvcc cross-compiles)."""
    need = [os.path.join(ROOT, "kmers.anno_b200", "libkmeranno.so"),
            os.path.join(ROOT, "kmers.anno_b200", "libkasynth.so"),
            os.path.join(ROOT, "oracle", "libkaoracle.so"),
            os.path.join(ROOT, "kmers.anno_b200", "bin", "kmers-anno"),
            os.path.join(ROOT, "kmers.anno_b200", "bin", "kmers-anno-selftest")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


def has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        if cuda.cuInit(0) != 0:
            return False
        n = ctypes.c_int()
        return cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False
