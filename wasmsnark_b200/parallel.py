"""Multi-GPU plumbing, one process per GPU (torchrun), torch.distributed / NCCL for the exchange steps.

The reference parallelises a proof like this (/root/reference/src/bn128.js):
  * every multiexp is split over the W workers by contiguous point ranges -- floor(n / W) points each, the last
    worker takes the remainder -- and the W partial points are added on the main thread (:353-383);
  * CALC_H (constructLC + five FFTs) is ONE task on ONE worker (:569-578); the H multiexp is queued when it
    resolves (:607-622).
The same structure is used across ranks here:
  * `shard_range` is the reference's split; each rank keeps only its ranges of A, B1, B2, C, hExps in HBM;
  * per proof ONE rank -- the H owner, rotating from proof to proof so that a pipeline of proofs loads all ranks
    equally -- computes h and sends every other rank its slice h[h_lo, h_lo + h_cnt) (NCCL send/recv over NVLink,
    4 MiB per rank at 2^20 constraints / 8 GPUs);
  * every rank runs its five partial MSMs; one all_gather of the 768-byte partial records; every rank adds the
    records and runs the final assembly (elliptic-curve addition is not an NCCL reduction operator, so
    all_gather + on-device add is the single-collective equivalent of an all-reduce of partial sums).
The one-process form of the same split (peer copies instead of NCCL) is ws_multi_* in the C ABI.
"""
import ctypes


def shard_range(n, rank, world):
    """(first index, count) of rank's contiguous range -- src/bn128.js:355-361."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    per = n // world
    lo = per * rank
    cnt = per if rank < world - 1 else n - per * (world - 1)
    return lo, cnt


def all_gather_records(record, world, group=None):
    """record: 1-D uint8 tensor (same length on every rank) -> 1-D tensor of world * len bytes,
    rank-major.  Works on the gloo backend (CPU tensors) and NCCL (CUDA tensors)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return record
    out = torch.empty(world * record.numel(), dtype=torch.uint8, device=record.device)
    if record.is_cuda:
        dist.all_gather_into_tensor(out, record, group=group)
    else:
        parts = [torch.empty_like(record) for _ in range(world)]
        dist.all_gather(parts, record, group=group)
        out = torch.cat(parts)
    return out


def scatter_slices(buf, elem_bytes, n_elems, owner, rank, world, group=None, ranges=None):
    """The H hand-over: `buf` is a 1-D uint8 tensor of n_elems * elem_bytes bytes on every rank; on `owner` it holds
    the whole array, afterwards every rank holds ITS slice at the slice's natural offset (the rest of `buf` is
    untouched on non-owners).  ranges: [(lo, cnt)] per rank (default: shard_range); ranks with an empty slice take no
    part.  One send per receiving rank from the owner, one receive on each of them; works on gloo (CPU tensors) and NCCL
    (CUDA tensors, ordered on the current stream)."""
    import torch.distributed as dist
    if world == 1:
        return
    if ranges is None:
        ranges = [shard_range(n_elems, r, world) for r in range(world)]
    ops = []
    if rank == owner:
        for r in range(world):
            if r == owner:
                continue
            lo, cnt = ranges[r]
            if cnt:
                ops.append(dist.P2POp(dist.isend, buf[lo * elem_bytes:(lo + cnt) * elem_bytes], r, group))
    else:
        lo, cnt = ranges[rank]
        if cnt:
            ops.append(dist.P2POp(dist.irecv, buf[lo * elem_bytes:(lo + cnt) * elem_bytes], owner, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()          # NCCL: orders the current stream after the transfer (does not block the host)


def witness_hull(key_ranges):
    """(first element, count) of the witness range a non-owner rank needs: the hull of its A/B1/B2 range and its C range
    (ws_bn128_pkey_shard_info); empty ranges do not count."""
    v_lo, v_cnt, c_lo, c_cnt = key_ranges[:4]
    spans = [(lo, lo + cnt) for lo, cnt in ((v_lo, v_cnt), (c_lo, c_cnt)) if cnt]
    if not spans:
        return 0, 0
    lo = min(a for a, _ in spans)
    return lo, max(b for _, b in spans) - lo


# ---- balanced split --------------------------------------------------------------------------------------------------
# The reference gives every worker 1/W of EVERY multiexp (src/bn128.js:353-361).  On GPUs a rank then runs five small
# MSMs per proof, each with its own bucket-reduction tail, and small point sets force narrower windows (more additions per
# point).  The sums are associative, so any partition of the points gives the same proof: here the five point sections
# are laid end to end, weighted by their cost per point, and cut into `world` equal pieces -- a rank gets one or two LARGE
# contiguous pieces (e.g. a third of B2, or the tail of A plus the head of B1).  Measured on B200s (2^20 constraints,
# proofs/s, balanced vs uniform): 101 vs 91 on 2 GPUs, 145 vs 159 on 4, 210 vs 282 on 8 -- the static weights do not
# capture the latency-bound G2 reduction tails, so with many ranks the ones holding B2 pieces lag; "auto" therefore uses
# the balanced split on 2 ranks only.
SECTION_ORDER = ("B2", "A", "B1", "C", "H")             # order along the line; index into ranges is A, B1, B2, C, H
_SEC_INDEX = {"A": 0, "B1": 1, "B2": 2, "C": 3, "H": 4}
BASE_WEIGHT = {"A": 100, "B1": 100, "B2": 280, "C": 100, "H": 120}    # G2 addition ~2.8 G1 additions; h is full-width


def section_weights(pkey, sample_step=16):
    """Integer cost per point of each section: the base weight scaled by the share of finite points (x != 0; points at
    infinity cost nothing, src/build_multiexp.js:335-349), estimated on every `sample_step`-th point so that every rank
    derives the same numbers from the same key bytes."""
    import struct
    import numpy as np
    hdr = struct.unpack_from("<10I", pkey, 0)
    n_vars, n_public, domain = hdr[0], hdr[1], hdr[2]
    off = {"A": hdr[5], "B1": hdr[6], "B2": hdr[7], "C": hdr[8], "H": hdr[9]}
    size = {"A": n_vars, "B1": n_vars, "B2": n_vars, "C": n_vars - n_public - 1, "H": domain}
    out = {}
    buf = np.frombuffer(pkey, dtype=np.uint8)
    for sec in SECTION_ORDER:
        psize, n = (128 if sec == "B2" else 64), size[sec]
        xbytes = 64 if sec == "B2" else 32
        idx = np.arange(0, n, sample_step, dtype=np.int64)
        if len(idx) == 0:
            out[sec] = BASE_WEIGHT[sec]
            continue
        rows = buf[off[sec]:off[sec] + n * psize].reshape(n, psize)[idx, :xbytes]
        finite = int(np.count_nonzero(rows.any(axis=1)))
        out[sec] = max(1, BASE_WEIGHT[sec] * max(finite, 1) // len(idx))
    return out, size


def balanced_ranges(weights, sizes, rank, world):
    """[lo, cnt] x 5 in section order A, B1, B2, C, H for `rank`: element i of a section with weight w occupies
    [off + i w, off + (i + 1) w) on the line; a rank owns the elements whose start falls into its 1/world of the line."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    total = sum(weights[s] * sizes[s] for s in SECTION_ORDER)
    start, end = rank * total // world, (rank + 1) * total // world
    out = [0] * 10
    off = 0
    for sec in SECTION_ORDER:
        w, n = weights[sec], sizes[sec]
        lo = min(n, max(0, -(-(start - off) // w)))
        hi = min(n, max(0, -(-(end - off) // w)))
        i = _SEC_INDEX[sec]
        out[2 * i], out[2 * i + 1] = (lo, hi - lo) if hi > lo else (0, 0)
        off += w * n
    return out


class ShardedProver:
    """Groth16 prover over the ranks of a torch.distributed group (NCCL), structured like the reference's worker pool:
    point sections range-sharded, H computed by one rank per proof (rotating).  Every rank ends up with the same
    proof, bit-identical to the single-GPU one.  All device work of a proof is enqueued on ONE torch.cuda.Stream given
    by the caller (the library launches on it, NCCL orders itself against it): no host synchronisation inside."""

    def __init__(self, bn, pkey_bytes, rank, world, group=None, split="auto"):
        """split: "uniform" (1/world of every section, the reference's worker split), "balanced" (cost-weighted contiguous
        pieces, see balanced_ranges) or "auto" = balanced on 2 ranks, uniform otherwise (what measured fastest: with many
        ranks the exactness of the uniform split beats the larger pieces of the balanced one)."""
        import torch
        from ._lib import as_ptr
        self.bn, self.rank, self.world, self.group = bn, rank, world, group
        self.key = ctypes.c_void_p()
        if split == "auto":
            split = "balanced" if world == 2 else "uniform"
        self.split = split if world > 1 else "uniform"
        if self.split == "balanced":
            weights, sizes = section_weights(pkey_bytes)
            self.all_ranges = [balanced_ranges(weights, sizes, r, world) for r in range(world)]
            arr = (ctypes.c_uint32 * 10)(*self.all_ranges[rank])
            bn.ctx.check(bn.lib.ws_bn128_pkey_load_ranges(bn.ctx.h, as_ptr(pkey_bytes), len(pkey_bytes), arr, ctypes.byref(self.key)))
            self.h_ranges = [(rg[8], rg[9]) for rg in self.all_ranges]
        else:
            bn.ctx.check(bn.lib.ws_bn128_pkey_load_shard(bn.ctx.h, as_ptr(pkey_bytes), len(pkey_bytes), rank, world, ctypes.byref(self.key)))
            self.h_ranges = None
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        bn.lib.ws_bn128_pkey_info(self.key, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        self.n_vars, self.n_public, self.domain = a.value, b.value, c.value
        rg = (ctypes.c_uint32 * 6)()
        bn.lib.ws_bn128_pkey_shard_info(self.key, rg)
        self.ranges = tuple(rg)
        dev = torch.device("cuda", bn.ctx.device)
        if self.h_ranges is None:
            self.h_ranges = [shard_range(self.domain, r, world) for r in range(world)]
        assert tuple(self.h_ranges[rank]) == (self.ranges[4], self.ranges[5])
        self.d_h = torch.zeros(self.domain * 32, dtype=torch.uint8, device=dev)
        self.d_part = torch.zeros(768, dtype=torch.uint8, device=dev)
        self.d_all = torch.zeros(768 * world, dtype=torch.uint8, device=dev)
        self.count = 0

    def next_owner(self):
        """H owner of the next proof: the same sequence on every rank."""
        o = self.count % self.world
        self.count += 1
        return o

    def upload_witness(self, d_witness, h_witness, owner, stream):
        """Host -> device copy of what this rank needs of the witness for a proof owned by `owner`: everything on the
        owner (the LC step reads every signal), the hull of its point ranges elsewhere.  h_witness: pinned uint8 tensor."""
        import torch
        with torch.cuda.stream(stream):
            if self.rank == owner:
                d_witness.copy_(h_witness, non_blocking=True)
                return h_witness.numel()
            lo, cnt = witness_hull(self.ranges)
            if cnt:
                d_witness[lo * 32:(lo + cnt) * 32].copy_(h_witness[lo * 32:(lo + cnt) * 32], non_blocking=True)
            return cnt * 32

    def prove_dev(self, d_witness, d_rs, d_out, stream, owner=None):
        """d_witness / d_rs / d_out: CUDA uint8 tensors (nVars*32, 64, 384 bytes); stream: the torch.cuda.Stream all of
        the proof's work is ordered on (required).  owner: the rank that computes H (default: rotates)."""
        import torch
        if stream is None or not hasattr(stream, "cuda_stream"):
            raise ValueError("prove_dev needs a torch.cuda.Stream")
        if owner is None:
            owner = self.next_owner()
        bn, h, lib = self.bn, self.bn.ctx.h, self.bn.lib
        st = ctypes.c_void_p(stream.cuda_stream)
        h_lo, h_cnt = self.ranges[4], self.ranges[5]
        with torch.cuda.stream(stream):
            bn.ctx.check(lib.ws_bn128_groth16_prepare_dev(h, self.key, d_rs.data_ptr(), st))
            bn.ctx.check(lib.ws_bn128_groth16_witness_msms_dev(h, self.key, d_witness.data_ptr(), self.d_part.data_ptr(), st))
            if self.rank == owner:
                bn.ctx.check(lib.ws_bn128_groth16_h_dev(h, self.key, d_witness.data_ptr(), self.d_h.data_ptr(), st))
            scatter_slices(self.d_h, 32, self.domain, owner, self.rank, self.world, self.group, ranges=self.h_ranges)
            bn.ctx.check(lib.ws_bn128_groth16_h_msm_dev(h, self.key, self.d_h.data_ptr() + 32 * h_lo, self.d_part.data_ptr(), st))
            src = self.d_part
            if self.world > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self.d_all, self.d_part, group=self.group)
                src = self.d_all
            bn.ctx.check(lib.ws_bn128_groth16_finish_dev(h, self.key, src.data_ptr(), self.world, d_rs.data_ptr(), d_out.data_ptr(), st))
        return d_out

    def free(self):
        if self.key:
            self.bn.lib.ws_bn128_pkey_free(self.key)
            self.key = None
