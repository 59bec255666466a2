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


# GPU tests written after the round's last B200 run (GPU budget spent): so far they have passed only on the kernel
# emulator (tests/host/cuemu), not on the hardware.  They run LAST, each under a hard time limit, so that under `pytest -x`
# a surprise in one of them cannot keep the hardware-verified tests (golden proofs, MSM / NTT parity) from running, and a
# hung kernel ends as a test failure instead of a hung box.  Nothing is skipped.
_EMULATOR_ONLY_SO_FAR = (
    "test_gpu_field.py::test_to_montgomery_takes_raw_values",
    "test_gpu_bls.py::test_to_montgomery_takes_raw_values",
    "test_gpu_msm.py::test_from_montgomery_methods_equal_the_reference_functions",
    "test_gpu_msm.py::test_back_to_back_async_multiexps_without_a_sync_between_them",
    "test_gpu_prover.py::test_degenerate_keys_and_witnesses",
    "test_gpu_prover.py::test_fft_reference_zeros_pattern",
    "test_gpu_prover.py::test_back_to_back_async_proofs_without_a_sync_between_them",
    "test_gpu_prover.py::test_int_mul_edge_grid_as_r_times_s",
    "test_gpu_prover.py::test_last_error_text_belongs_to_the_failing_thread",
    "test_gpu_prover.py::test_multi_handle_with_more_gpus_than_points",
    "test_node_addon_mock.py::test_prove_through_the_addon_equals_the_reference_proof",
    "test_node_addon_mock.py::test_overlapping_proofs_on_one_handle",
    "test_node_addon_mock.py::test_terminate_semantics",
    "test_node_addon_mock.py::test_multiexp_affine_calch_through_the_addon",
)


def _emulator_only_so_far(item):
    return any(key in item.nodeid for key in _EMULATOR_ONLY_SO_FAR)


def _collects_emulated_children(item):
    return "test_emulated_library.py" in item.nodeid and "children" in getattr(item, "fixturenames", ())


def pytest_collection_modifyitems(config, items):
    if os.environ.get("WS_TEST_EMULATED"):
        return
    # CPU suite: the emulated child runs (tests/test_emulated_library.py) are subprocesses -- started when the session starts
    # (fixture below), their results collected LAST, so they run beside the serial tests instead of in front of them
    waiting = [it for it in items if _collects_emulated_children(it)]
    if waiting:
        items[:] = [it for it in items if not _collects_emulated_children(it)] + waiting
    late = [it for it in items if _emulator_only_so_far(it)]
    if not late:
        return
    if config.pluginmanager.hasplugin("timeout"):
        for it in late:
            it.add_marker(pytest.mark.timeout(300, method="thread"))
    items[:] = [it for it in items if not _emulator_only_so_far(it)] + late


@pytest.fixture(scope="session", autouse=True)
def emulated_children_started_early(request):
    started = False
    if not os.environ.get("WS_TEST_EMULATED") and sum(_collects_emulated_children(it) for it in request.session.items) > 1:
        import test_emulated_library
        test_emulated_library.start_children()
        started = True
    yield
    if started:
        import test_emulated_library
        test_emulated_library.stop_children()


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
