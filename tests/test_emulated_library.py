"""The library's own CUDA kernels EXECUTED ON THE CPU (tests/host/cuemu) -- a selection of the GPU parity tests re-run without a GPU.

tests/host/cuemu/build_emu.py compiles the unchanged wasmsnark_b200/csrc/*.cu with g++ against a kernel emulator (one fiber
per CUDA thread; __syncthreads / __shfl_*_sync / __syncwarp as rendezvous points; device memory = host memory; a synchronous
CUDA runtime with CUEMU_DEVICES pretend devices).  A child pytest with WS_TEST_EMULATED=1 (tests/conftest.py) then runs the
ordinary `-m gpu` tests against that build: the launch geometry, shared-memory exchanges, scans, warp-cooperative chains,
NTT passes, the prover's stream choreography, the stateless key cache, the thread-safe contexts, the N-API addon and the
multi-GPU handle (four pretend GPUs) -- everything except the PTX carry chains of the field layer, which take their portable
host path here (tests/test_fp_asm_model.py and the B200 run cover those) and except performance.

This is test scaffolding, not a CPU fallback: the package never looks for the emulated library; only conftest's
WS_TEST_EMULATED hook hands it over.  The full GPU suite passes under emulation too (27 min, incl. the 2^20-constraint golden
proof); the selection below is sized for the CPU suite.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FIELD_MSM = [
    "tests/test_gpu_field.py",
    "tests/test_gpu_msm.py::test_small_sets_vs_oracle",
    "tests/test_gpu_msm.py::test_group_identities_from_reference_suite",
    "tests/test_gpu_msm.py::test_2p14_vs_reference_multiexp2",
    "tests/test_gpu_msm.py::test_window_table_msm",
]
PROVER = [
    "tests/test_gpu_prover.py::test_fft_golden_vectors",
    "tests/test_gpu_prover.py::test_fft_vs_oracle_all_radix_mixes",
    "tests/test_gpu_prover.py::test_fft_rejects_bad_sizes",
    "tests/test_gpu_prover.py::test_fft_interleave_identity[16]",
    "tests/test_gpu_prover.py::test_fft_in_place_pass_counts_and_variants[13]",
    "tests/test_gpu_prover.py::test_synth_small_calch_and_proofs",
    "tests/test_gpu_prover.py::test_stateless_call_key_cache",
    "tests/test_gpu_prover.py::test_random_rs_differs_and_bad_keys_rejected",
    "tests/test_gpu_prover.py::test_sharded_partials_emulated_on_one_gpu",
    "tests/test_gpu_prover.py::test_split_prover_entry_points_equal_the_single_call",
    "tests/test_gpu_prover.py::test_balanced_split_emulated_on_one_gpu[3]",
    "tests/test_gpu_prover.py::test_last_error_text_belongs_to_the_failing_thread",
    "tests/test_gpu_prover.py::test_degenerate_keys_and_witnesses",
    "tests/test_node_addon_mock.py",
]
OTHER_CURVES = [
    "tests/test_gpu_mnt.py::test_known_answer_and_small_sets",
    "tests/test_gpu_mnt.py::test_fq3_mul_1k",
    "tests/test_gpu_bls.py::test_g1_multiexp_small",
    "tests/test_gpu_bls.py::test_g2_multiexp_small",
    "tests/test_gpu_bls.py::test_field_ops",
    "tests/test_gpu_bls.py::test_fft_vs_oracle",
]
MULTI = ["tests/test_gpu_prover.py::test_multi_handle_on_the_visible_gpus"]


MULTI_SPARSE = ["tests/test_gpu_prover.py::test_multi_handle_with_more_gpus_than_points"]
GROUPS = {"field_msm": (FIELD_MSM, 1), "prover": (PROVER, 1), "other_curves": (OTHER_CURVES, 1), "multi4": (MULTI, 4), "multi8": (MULTI_SPARSE, 8)}


@pytest.fixture(scope="module")
def children():
    """the child pytest runs, started together (each is single-threaded; the emulated build is made once, first)"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host", "cuemu"))
    import build_emu
    build_emu.build()
    procs = {}
    for name, (selection, devices) in GROUPS.items():
        env = dict(os.environ, WS_TEST_EMULATED="1", CUEMU_DEVICES=str(devices))
        procs[name] = subprocess.Popen([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + selection, cwd=ROOT, env=env,
                                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    yield procs
    for p in procs.values():
        if p.poll() is None:
            p.kill()


def finished(children, name, timeout=1500):
    out, _ = children[name].communicate(timeout=timeout)
    assert children[name].returncode == 0, out[-3000:]
    assert " passed" in out and " failed" not in out, out[-3000:]
    return out


def test_field_and_msm_kernels_run_on_the_cpu_emulator(children):
    finished(children, "field_msm")


def test_ntt_prover_cache_threads_and_addon_on_the_cpu_emulator(children):
    finished(children, "prover")


def test_mnt6753_and_bls12381_kernels_on_the_cpu_emulator(children):
    finished(children, "other_curves")


def test_multi_gpu_handle_with_four_pretend_devices(children):
    """ws_multi_* (csrc/multi.cu) over four pretend GPUs: key shards per device, the rotating H owner, peer copies of the h
    slices and of the 768-byte records, assembly on device 0 -- bit-identical to the single-context proof."""
    assert "1 passed" in finished(children, "multi4")


def test_multi_gpu_handle_with_more_pretend_devices_than_points(children):
    assert "1 passed" in finished(children, "multi8")
