import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")
    if os.environ.get("WS_TEST_EMULATED"):
        # tests/test_emulated_library.py re-runs a selection of the GPU tests in a child pytest with the library's kernels
        # EXECUTED ON THE CPU (tests/host/cuemu: the same .cu sources compiled by g++ against a kernel emulator).  Test
        # scaffolding: only this hook ever hands the package that library.
        sys.path.insert(0, os.path.join(ROOT, "tests", "host", "cuemu"))
        import build_emu
        from wasmsnark_b200 import _lib
        # WS_TEST_EMULATED=asan: the AddressSanitizer build (run python under LD_PRELOAD=$(gcc -print-file-name=libasan.so))
        # WS_TEST_EMULATED=ubsan: the UndefinedBehaviorSanitizer build
        _lib.load_library(path=build_emu.build(asan=os.environ["WS_TEST_EMULATED"] == "asan", ubsan=os.environ["WS_TEST_EMULATED"] == "ubsan"))
        import emulated_torch
        emulated_torch.patch()


@pytest.fixture(scope="session")
def host_math():
    """ctypes handle on the HOST build of the arithmetic headers (tests/host/host_math.cpp)."""
    import ctypes
    src = os.path.join(ROOT, "tests", "host", "host_math.cpp")
    so = os.path.join(ROOT, "tests", "host", "libhost_math.so")
    deps = [src] + [os.path.join(ROOT, "wasmsnark_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "wasmsnark_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)
