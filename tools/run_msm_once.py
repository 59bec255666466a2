"""Profiling driver: a few G1 MSMs of 2^k pairs on device-resident inputs (for ncu / stage timing).
usage: python tools/run_msm_once.py [log2n] [reps] [window_bits] [task_cap] [dist]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
from wasmsnark_b200 import build_bn128  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dist = sys.argv[5] if len(sys.argv) > 5 else "fr"
seg = int(sys.argv[6]) if len(sys.argv) > 6 else 0
variant = int(sys.argv[7]) if len(sys.argv) > 7 else 0
group = int(sys.argv[8]) if len(sys.argv) > 8 else 1
n = 1 << log2n
scalars, points, P0, D = bench.make_inputs(n, dist=dist)
d_s = torch.frombuffer(bytearray(scalars), dtype=torch.uint8).cuda()
d_p = torch.frombuffer(bytearray(points), dtype=torch.uint8).cuda()
d_out = torch.zeros(256, dtype=torch.uint8, device="cuda")
bn = build_bn128(0)
bn.lib.ws_set_profiling(bn.ctx.h, 1)
bn.lib.ws_set_msm_task_cap(bn.ctx.h, cap)
bn.lib.ws_set_msm_seg_len(bn.ctx.h, seg)
bn.lib.ws_set_msm_acc_variant(bn.ctx.h, variant)
tot = {}
for r in range(reps):
    torch.cuda.synchronize()
    bn.ctx.check(bn.lib.ws_bn128_g1_multiexp_dev(bn.ctx.h, d_s.data_ptr(), d_p.data_ptr(), n, d_out.data_ptr(), wb, None))
    t = bn.ctx.timings()
    if r:
        for k, v in t:
            tot[k] = tot.get(k, 0) + v
print("n=2^%d c=%d cap=%d dist=%s L=%d" % (log2n, wb, cap, dist, seg), " ".join("%s=%.3f" % (k, v / max(1, reps - 1)) for k, v in tot.items()), "total=%.3f" % (sum(tot.values()) / max(1, reps - 1)))
