"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange step.

The reference parallelises every multiexp by giving each worker a contiguous point range --
floor(n / W) points each, the last worker takes the remainder -- and adding the W partial points on
the main thread (/root/reference/src/bn128.js:353-383).  The same split is used across ranks here;
the serial add loop becomes one all_gather of fixed-size partial records followed by a tiny on-device
combine (elliptic-curve addition is not an NCCL reduction operator, so all_gather + local reduce is
the single-collective equivalent of an all-reduce).
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


class ShardedProver:
    """Groth16 prover with the proving key's point sections range-sharded over the ranks of a
    torch.distributed group (NCCL).  Every rank ends up with the same proof."""

    def __init__(self, bn, pkey_bytes, rank, world, group=None):
        from ._lib import as_ptr
        self.bn, self.rank, self.world, self.group = bn, rank, world, group
        self.key = ctypes.c_void_p()
        bn.ctx.check(bn.lib.ws_bn128_pkey_load_shard(bn.ctx.h, as_ptr(pkey_bytes), len(pkey_bytes), rank, world, ctypes.byref(self.key)))

    def prove_dev(self, d_witness, d_rs, d_out, stream_ptr):
        """d_witness / d_rs / d_out: CUDA uint8 tensors (nVars*32, 64, 384 bytes)."""
        import torch
        bn = self.bn
        part = torch.empty(768, dtype=torch.uint8, device=d_witness.device)
        bn.ctx.check(bn.lib.ws_bn128_groth16_prepare_dev(bn.ctx.h, self.key, d_rs.data_ptr(), stream_ptr))
        bn.ctx.check(bn.lib.ws_bn128_groth16_partials_dev(bn.ctx.h, self.key, d_witness.data_ptr(), part.data_ptr(), stream_ptr))
        allp = all_gather_records(part, self.world, self.group)
        bn.ctx.check(bn.lib.ws_bn128_groth16_finish_dev(bn.ctx.h, self.key, allp.data_ptr(), self.world, d_rs.data_ptr(), d_out.data_ptr(), stream_ptr))
        return d_out

    def free(self):
        if self.key:
            self.bn.lib.ws_bn128_pkey_free(self.key)
            self.key = None
