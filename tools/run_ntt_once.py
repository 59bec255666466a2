"""Profiling driver: a few forward + inverse BN128 Fr transforms (for ncu captures of the NTT passes).
usage: python tools/run_ntt_once.py [log2 n = 22] [reps = 2] [variant = 0]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wasmsnark_b200 import build_bn128  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 1 << lg
bn = build_bn128(0)
bn.lib.ws_set_ntt_variant(bn.ctx.h, variant)
x = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
x[:, 31] &= 0x1F
x0 = x.clone()
bn.lib.ws_set_profiling(bn.ctx.h, 1)
for _ in range(reps):
    bn.ctx.check(bn.lib.ws_bn128_fr_fft_dev(bn.ctx.h, x.data_ptr(), n, 0, 0, None))
    fwd = bn.ctx.timings()
    bn.ctx.check(bn.lib.ws_bn128_fr_fft_dev(bn.ctx.h, x.data_ptr(), n, 0, 1, None))
    inv = bn.ctx.timings()
torch.cuda.synchronize()
print("variant %d" % variant, "ntt 2^%d: forward %.3f ms, inverse %.3f ms, round trip bit-exact: %s" % (lg, sum(v for _, v in fwd), sum(v for _, v in inv), bool(torch.equal(x, x0))))
