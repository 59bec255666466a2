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
        _lib.load_library(path=build_emu.build(asan=os.environ["WS_TEST_EMULATED"] == "asan"))
        _emulated_torch_device()


def _emulated_torch_device():
    """Under emulation "device memory" is host memory: the tests' torch.cuda buffers become CPU tensors (their data_ptr()
    is what the emulated library dereferences)."""
    import torch

    def on_cpu(fn):
        def wrapper(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapper
    for name in ("zeros", "empty", "ones", "full", "tensor", "frombuffer", "arange"):
        setattr(torch, name, on_cpu(getattr(torch, name)))
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: max(1, int(os.environ.get("CUEMU_DEVICES", "1")))
    torch.cuda.set_device = lambda *a, **k: None


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
