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
    "tests/test_gpu_prover.py::test_fft_reference_zeros_pattern",
    "tests/test_gpu_prover.py::test_fft_interleave_identity[16]",
    "tests/test_gpu_prover.py::test_fft_in_place_pass_counts_and_variants[13]",
    "tests/test_gpu_prover.py::test_synth_small_calch_and_proofs",
    "tests/test_gpu_prover.py::test_stateless_call_key_cache",
    "tests/test_gpu_prover.py::test_random_rs_differs_and_bad_keys_rejected",
    "tests/test_gpu_prover.py::test_sharded_partials_emulated_on_one_gpu",
    "tests/test_gpu_prover.py::test_split_prover_entry_points_equal_the_single_call",
    "tests/test_gpu_prover.py::test_balanced_split_emulated_on_one_gpu[3]",
    "tests/test_gpu_prover.py::test_back_to_back_async_proofs_without_a_sync_between_them",
    "tests/test_gpu_prover.py::test_last_error_text_belongs_to_the_failing_thread",
    "tests/test_gpu_prover.py::test_degenerate_keys_and_witnesses",
    "tests/test_gpu_prover.py::test_int_mul_edge_grid_as_r_times_s",
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
# the prover's stream choreography and the multi-GPU handle once more with the streams executed DEFERRED (cuemu.cpp, CUEMU_ASYNC):
# lazily (a consumer without a wait on its producer's event reads stale bytes) and eagerly (a buffer reused before its reader was
# ordered is read overwritten) -- proofs must not change
PROVER_STREAMS = [t for t in PROVER if "test_fft" not in t and "degenerate" not in t and "int_mul" not in t] + \
    ["tests/test_gpu_prover.py::test_context_is_thread_safe", "tests/test_gpu_msm.py::test_back_to_back_async_multiexps_without_a_sync_between_them"]
GROUPS = {"field_msm": (FIELD_MSM, 1, ""), "prover": (PROVER, 1, ""), "other_curves": (OTHER_CURVES, 1, ""), "multi4": (MULTI, 4, ""), "multi8": (MULTI_SPARSE, 8, ""),
          "prover_lazy": (PROVER_STREAMS, 1, "lazy"), "prover_eager": (PROVER_STREAMS, 1, "eager"), "multi4_lazy": (MULTI, 4, "lazy"), "multi4_eager": (MULTI, 4, "eager")}


_PROCS = {}


def start_children():
    """the child pytest runs, started together (each is single-threaded; the emulated build is made once, first).  Called at
    the START of the session (tests/conftest.py) so that they run beside the serial tests of the CPU suite; the tests below
    only collect the results"""
    if _PROCS:
        return _PROCS
    sys.path.insert(0, os.path.join(ROOT, "tests", "host", "cuemu"))
    import build_emu
    build_emu.build()
    for name, (selection, devices, streams) in GROUPS.items():
        env = dict(os.environ, WS_TEST_EMULATED="1", CUEMU_DEVICES=str(devices), CUEMU_ASYNC=streams)
        _PROCS[name] = subprocess.Popen([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + selection, cwd=ROOT, env=env,
                                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return _PROCS


def stop_children():
    for p in _PROCS.values():
        if p.poll() is None:
            p.kill()
    _PROCS.clear()


@pytest.fixture(scope="module")
def children():
    yield start_children()
    stop_children()


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


@pytest.mark.parametrize("group", ["prover_lazy", "prover_eager", "multi4_lazy", "multi4_eager"])
def test_stream_choreography_under_deferred_execution(children, group):
    """The prover (side streams for the witness MSMs and the (r, s) products, the H chain on the caller's stream, the joins in
    front of the assembly, the stateless call's key cache and staging ring, four host threads on one context) and the
    multi-GPU handle (peer copies of the h slices and records, cross-device events, the rotating H owner) with every stream
    executed deferred: lazily and eagerly.  Same proofs as always -- the streams are ordered by stream order and events, not
    by timing."""
    assert " passed" in finished(children, group)


# ---- the torchrun form of the multi-GPU prover (wasmsnark_b200/parallel.py: ShardedProver) on gloo + emulated kernels ------------
def _sharded_worker(rank, world, port, split, q):
    """one rank = one process = one emulated GPU: key shard, witness MSMs, H on the rotating owner, send / recv of the h
    slices, all_gather of the 768-byte records, assembly -- the code path bench.py --gpus N times, with gloo for NCCL"""
    import json
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CUEMU_DEVICES="1")
        sys.path.insert(0, os.path.join(ROOT, "tests", "host", "cuemu"))
        import build_emu
        import emulated_torch
        from wasmsnark_b200 import _lib
        _lib.load_library(path=build_emu.build())
        emulated_torch.patch()
        import torch
        import torch.distributed as dist
        from wasmsnark_b200 import build_bn128
        from wasmsnark_b200.parallel import ShardedProver
        dist.init_process_group("gloo", rank=rank, world_size=world)
        gold = os.path.join(ROOT, "tests", "golden")
        pk = open(os.path.join(gold, "synth_small.pkey.bin"), "rb").read()
        wit = open(os.path.join(gold, "synth_small.witness.bin"), "rb").read()
        g = json.load(open(os.path.join(gold, "synth_small.golden.json")))
        bn = build_bn128(0)
        sp = ShardedProver(bn, pk, rank, world, split=split)
        h_wit = torch.frombuffer(bytearray(wit), dtype=torch.uint8)
        d_wit = torch.zeros(len(wit), dtype=torch.uint8, device="cuda")
        d_rs = torch.frombuffer(bytearray(bytes.fromhex(g["r"]) + bytes.fromhex(g["s"])), dtype=torch.uint8).cuda()
        d_out = torch.zeros(384, dtype=torch.uint8, device="cuda")
        stream = torch.cuda.Stream()
        proofs = []
        for _ in range(world + 1):                      # every rank takes a turn as H owner
            owner = sp.next_owner()
            d_wit.zero_()
            sp.upload_witness(d_wit, h_wit, owner, stream)          # non-owners get only the ranges they need
            out = bytes(sp.prove_dev(d_wit, d_rs, d_out, stream, owner=owner).numpy())
            ints = [str(int.from_bytes(out[32 * i:32 * i + 32], "little")) for i in range(12)]
            proofs.append({"pi_a": ints[0:3], "pi_b": [ints[3:5], ints[5:7], ints[7:9]], "pi_c": ints[9:12]})
        sp.free()
        bn.terminate()
        dist.destroy_process_group()
        q.put((rank, all(p == g["proof_rs"] for p in proofs), len(proofs)))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:] + repr(e)))


@pytest.mark.parametrize("world,split", [(2, "balanced"), (3, "uniform")])
def test_sharded_prover_over_gloo_with_emulated_kernels(world, split):
    """parallel.ShardedProver end to end without a GPU: `world` processes over gloo, each with the emulated library; every
    rank must end up with the reference's golden proof for every H owner."""
    import random
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "host", "cuemu"))
    import build_emu
    build_emu.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + random.randrange(2000)
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, split, q)) for r in range(world)]
    [p.start() for p in procs]
    try:
        got = sorted(q.get(timeout=600) for _ in range(world))
    finally:
        [p.join(timeout=60) for p in procs]
        [p.kill() for p in procs if p.is_alive()]
    assert got == [(r, True, world + 1) for r in range(world)], got


def test_driver_smoke_hook_on_the_cpu_emulator():
    """__graft_entry__.smoke() (what the driver runs on cuda:0 before the bench) through tests/host/cuemu/run_emulated.py: its own
    logic is checked here; on the B200 box it runs against the real library."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host", "cuemu", "run_emulated.py"), os.path.join(ROOT, "__graft_entry__.py"), "smoke"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "smoke ok" in r.stdout, (r.stdout + r.stderr)[-2000:]


@pytest.mark.parametrize("order,streams", [("", ""), ("reverse", ""), ("shuffle", ""), ("", "lazy"), ("", "eager"), ("", "random")])
def test_emulator_self_test(order, streams):
    """The emulator itself on kernels with known answers (tests/host/cuemu/selftest.cpp), so that a pass of the suites above
    means something: a shuffle hands over the value of a lane that takes part and POISON for a lane that does not (one that
    returned earlier, or beyond the last thread of a partial warp -- undefined data on the GPU, where the emulator used to
    return the lane's stale slot); barriers order a shared-memory exchange for every thread order; launch configurations the
    hardware refuses are refused with cudaErrorInvalidConfiguration.  Streams (CUEMU_ASYNC, deferred execution): two programs
    whose streams are NOT ordered by events -- a consumer without a wait on its producer, a buffer refilled before its reader
    was ordered -- give the right bytes synchronously, stale bytes under `lazy` resp. overwritten bytes under `eager`; their
    repaired versions give the right bytes in every mode; a launch's arguments are those of the moment of the launch."""
    here = os.path.join(ROOT, "tests", "host", "cuemu")
    exe = os.path.join(here, "_build", "selftest")
    deps = [os.path.join(here, f) for f in ("selftest.cpp", "cuemu.cpp", "cuda_runtime.h")]
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-w", "-D__CUDACC__=1", "-I" + here, "-include", os.path.join(here, "cuda_runtime.h"),
                               deps[0], deps[1], "-lpthread", "-o", exe + ".tmp"])
        os.replace(exe + ".tmp", exe)
    if streams == "random":        # single steps of randomly chosen streams: the repaired programs are right for every seed, the
        seen = set()               # mis-ordered ones go wrong for some (seeds 1..8 show each defect at least once)
        for seed in range(1, 9):
            r = subprocess.run([exe], env=dict(os.environ, CUEMU_ASYNC="random", CUEMU_SEED=str(seed)), capture_output=True, text=True, timeout=120)
            assert r.returncode == 0, r.stderr[-2000:]
            got = dict((k, int(v)) for k, v in (line.split() for line in r.stdout.splitlines()))
            assert (got["ordered_consumer_sees"], got["ordered_reader_sees"], got["launch_arguments_kept"], got["exchange_good"]) == (1, 5, 1, 576), (seed, got)
            seen.add((got["unordered_consumer_sees"], got["unordered_reader_sees"]))
        assert any(c == 0 for c, _ in seen) and any(r == 9 for _, r in seen), seen
        return
    r = subprocess.run([exe], env=dict(os.environ, CUEMU_ORDER=order, CUEMU_ASYNC=streams), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    got = dict((k, int(v)) for k, v in (line.split() for line in r.stdout.splitlines()))
    assert got == {"shuffle_good": 24, "shuffle_poisoned": 16, "shuffle_stale": 0, "exchange_good": 576, "empty_grid_error": 9,
                   "too_many_threads_error": 9, "error_cleared": 0, "launch_arguments_kept": 1,
                   "unordered_consumer_sees": 0 if streams == "lazy" else 1, "unordered_reader_sees": 9 if streams == "eager" else 5,
                   "ordered_consumer_sees": 1, "ordered_reader_sees": 5}, got
