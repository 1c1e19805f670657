import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
sys.path.insert(0, ROOT)


# the key-switch inner product of four-chunk limbs runs as two launches only for large batches (tu_ks.cu); the tests' batches
# are small, so the threshold is lowered for every library and plugin process the tests start
os.environ.setdefault("B200HE_KS_SPLIT_MIN", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """the product's CUDA sources compiled as host C++ (tests/emu): checks kernel index arithmetic on CPU"""
    import ctypes
    import pyb200he
    if os.environ.get("B200HE_EMU_LIB"):   # e.g. an AddressSanitizer build of the same sources (tests/emu/build_emu_asan.sh)
        return pyb200he.declare(ctypes.CDLL(os.environ["B200HE_EMU_LIB"]))
    so = os.path.join(ROOT, "tests", "emu", "libb200he_emu.so")
    srcs = [os.path.join(ROOT, "reference-seal-backend_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "reference-seal-backend_b200", "csrc"))]
    srcs += [os.path.join(ROOT, "tests", "emu", f) for f in ("cuda_shim.cpp", "cuda_shim.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build_emu.sh")])
    return pyb200he.declare(ctypes.CDLL(so))


@pytest.fixture(scope="session")
def gpu_lib():
    import pyb200he
    return pyb200he.load_library()
