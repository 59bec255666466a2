"""tests/host/cuemu/emulated_torch.py -- TEST SCAFFOLDING: under the kernel emulator "device memory" is host memory, so the
tests' torch.cuda buffers become CPU tensors (their data_ptr() is what the emulated library dereferences) and CUDA streams
become no-ops.  Used by tests/conftest.py (WS_TEST_EMULATED) and by the gloo workers of tests/test_emulated_library.py."""
import contextlib
import os


def drain():
    """CUEMU_ASYNC (deferred execution, cuemu.cpp): everything queued on any stream runs now.  Called where the real torch
    would order itself behind the library's work -- torch.cuda.synchronize(), stream / event .synchronize(), tensor.cpu(), a
    collective on the stream the work was queued on.  A no-op in the synchronous mode."""
    from wasmsnark_b200 import _lib
    lib = _lib.load_library()
    if hasattr(lib, "cuemu_sync_all"):
        lib.cuemu_sync_all()


class FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        drain()

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def record_event(self, ev=None):
        ev = ev or FakeEvent()
        ev.record()
        return ev


class FakeEvent:
    """everything is synchronous under emulation: an event is a host timestamp"""

    def __init__(self, *a, **k):
        self.t = None

    def record(self, stream=None):
        import time
        self.t = time.perf_counter()

    def synchronize(self):
        drain()

    def wait(self, stream=None):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return (other.t - self.t) * 1000.0


def patch():
    import torch

    def on_cpu(fn):
        def wrapper(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapper
    if getattr(torch, "_ws_emulated", False):
        return
    torch._ws_emulated = True
    for name in ("zeros", "empty", "ones", "full", "tensor", "frombuffer", "arange", "randint"):
        setattr(torch, name, on_cpu(getattr(torch, name)))
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()
    if os.environ.get("CUEMU_ASYNC"):
        # a device-to-host read is ordered behind the stream's work; so is a collective enqueued on that stream
        cpu = torch.Tensor.cpu

        def cpu_after_the_streams(self, *a, **k):
            drain()
            return cpu(self, *a, **k)
        torch.Tensor.cpu = cpu_after_the_streams
        import torch.distributed as dist

        def after_the_streams(fn):
            def wrapper(*a, **k):
                drain()
                return fn(*a, **k)
            return wrapper
        for name in ("all_gather_into_tensor", "all_gather", "all_reduce", "broadcast", "send", "recv", "batch_isend_irecv", "barrier"):
            if hasattr(dist, name):
                setattr(dist, name, after_the_streams(getattr(dist, name)))
    torch.cuda.synchronize = lambda *a, **k: drain()
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: max(1, int(os.environ.get("CUEMU_DEVICES", "1")))
    torch.cuda.set_device = lambda *a, **k: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.cuda.Stream = FakeStream
    torch.cuda.Event = FakeEvent
    torch.cuda.set_stream = lambda *a, **k: None
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.current_stream = lambda *a, **k: FakeStream()
