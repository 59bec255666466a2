"""N > 1 host logic on CPU: world_size-2 gloo group.  Each rank takes its contiguous point range
(shard_range = the reference's worker split, src/bn128.js:353-361), computes its partial sum with the
oracle standing in for the GPU kernel, the 96-byte partial records are exchanged with the same
all_gather helper the NCCL path uses, and the combined point must equal the unsharded sum."""
import os
import random

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bn128 as O
from wasmsnark_b200.parallel import all_gather_records, balanced_ranges, scatter_slices, section_weights, shard_range, witness_hull


def test_shard_range_matches_reference_split():
    for n in (0, 1, 7, 8, 66232, 1 << 20):
        for w in (1, 2, 3, 8):
            ranges = [shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0
            for (lo, cnt), (lo2, _) in zip(ranges, ranges[1:]):
                assert lo + cnt == lo2
            assert ranges[-1][0] + ranges[-1][1] == n
            assert all(cnt == n // w for _, cnt in ranges[:-1])       # floor(n/W) each, last takes the rest
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, scalars, points, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, cnt = shard_range(len(scalars), rank, world)
    part = O.multiexp(O.Fq, scalars[lo:lo + cnt], points[lo:lo + cnt])
    rec = torch.frombuffer(bytearray(O.g1_to_bytes_m(part)), dtype=torch.uint8)
    allrec = all_gather_records(rec, world)
    assert allrec.numel() == 96 * world
    acc = O.pt_zero(O.Fq)
    raw = bytes(allrec.numpy())
    for r in range(world):
        x, y, z = (O.from_mont(int.from_bytes(raw[96 * r + 32 * k:96 * r + 32 * k + 32], "little")) for k in range(3))
        acc = O.pt_add(O.Fq, acc, (x, y, z))
    q.put((rank, O.pt_affine(O.Fq, acc)))
    dist.destroy_process_group()


def test_two_rank_sharded_sum_equals_unsharded():
    rnd = random.Random(9)
    n = 37
    base = (O.G1_GEN[0], O.G1_GEN[1], 1)
    points = [O.pt_times_scalar(O.Fq, base, rnd.randrange(1, O.R_ORDER)) for _ in range(n)]
    scalars = [rnd.getrandbits(256) for _ in range(n)]
    want = O.pt_affine(O.Fq, O.multiexp(O.Fq, scalars, points))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scalars, points, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(r for r, _ in got) == [0, 1]
    assert all(res == want for _, res in got)


def _scatter_worker(rank, world, port, n_elems, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    full = torch.arange(n_elems * 32, dtype=torch.int64).remainder(251).to(torch.uint8)          # what the H owner computed
    for owner in range(world):                    # the owner rotates from proof to proof
        buf = full.clone() if rank == owner else torch.full((n_elems * 32,), 255, dtype=torch.uint8)
        scatter_slices(buf, 32, n_elems, owner, rank, world)
        lo, cnt = shard_range(n_elems, rank, world)
        ok &= bool(torch.equal(buf[lo * 32:(lo + cnt) * 32], full[lo * 32:(lo + cnt) * 32]))
        if rank != owner:                         # nothing outside the slice was written
            rest = torch.cat([buf[:lo * 32], buf[(lo + cnt) * 32:]])
            ok &= bool((rest == 255).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_h_slices_reach_every_rank_from_a_rotating_owner():
    """The H hand-over of the sharded prover (parallel.scatter_slices) on a world_size-2 gloo group, uneven split."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + random.randrange(2000)
    procs = [ctx.Process(target=_scatter_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(got) == [(0, True), (1, True)]


def test_witness_hull_covers_both_point_ranges():
    # (v_lo, v_cnt, c_lo incl. the nPublic + 1 offset, c_cnt, h_lo, h_cnt)
    assert witness_hull((100, 50, 130, 40, 0, 0)) == (100, 70)
    assert witness_hull((0, 10, 59, 5, 0, 0)) == (0, 64)
    assert witness_hull((0, 0, 0, 0, 0, 0)) == (0, 0)


def test_balanced_split_partitions_every_section_and_matches_the_c_library():
    """parallel.balanced_ranges: every point of every section belongs to exactly one rank, pieces are contiguous, the
    weighted load is even, and the C library (ws_balanced_ranges, what ws_multi_* uses) cuts identically."""
    import ctypes
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_synth_key
    from wasmsnark_b200 import _lib
    pk, _ = gen_synth_key.make(5000, 7, 4096, seed=3)
    weights, sizes = section_weights(pk)
    assert weights["B2"] > weights["A"] and weights["B1"] < 100            # the key's B points are 30 % infinity
    lib = _lib.load_library()
    for world in (1, 2, 3, 4, 8):
        rows = [balanced_ranges(weights, sizes, r, world) for r in range(world)]
        out = (ctypes.c_uint32 * (10 * world))()
        assert lib.ws_balanced_ranges(_lib.as_ptr(pk), len(pk), world, out) == 0
        assert [list(out[10 * r:10 * r + 10]) for r in range(world)] == rows
        loads = []
        for i, sec in enumerate(("A", "B1", "B2", "C", "H")):
            pos = 0
            for r in range(world):
                lo, cnt = rows[r][2 * i], rows[r][2 * i + 1]
                if cnt:
                    assert lo == pos
                    pos += cnt
            assert pos == sizes[sec]
        for r in range(world):
            loads.append(sum(rows[r][2 * i + 1] * weights[sec] for i, sec in enumerate(("A", "B1", "B2", "C", "H"))))
        assert max(loads) - min(loads) <= 2 * max(weights.values())
    assert lib.ws_balanced_ranges(_lib.as_ptr(pk[:100]), 100, 2, (ctypes.c_uint32 * 20)()) != 0
