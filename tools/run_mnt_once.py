"""Profiling driver: MNT6753 G1 MSM on 2^k pairs (1024 distinct points tiled), stage times from the library's timer.
usage: python tools/run_mnt_once.py [log2 n = 18] [reps = 2]"""
import ctypes
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import mnt6753 as MN  # noqa: E402  (test infrastructure: builds valid curve points)
from wasmsnark_b200 import build_bn128  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, distinct = 1 << lg, 1024
rnd = random.Random(0x6753)
P, D = MN.times_scalar(MN.G1_GEN, rnd.randrange(MN.R_ORDER)), MN.times_scalar(MN.G1_GEN, rnd.randrange(MN.R_ORDER))
base = []
for _ in range(distinct):
    base.append(MN.affine_bytes_m(P))
    P = MN.add(P, D)
pts = torch.frombuffer(bytearray(b"".join(base)), dtype=torch.uint8).view(distinct, 192).cuda().repeat(n // distinct, 1).contiguous()
sc_np = np.random.default_rng(1).integers(0, 256, size=(n, 96), dtype=np.uint8)
sc_np[:, 94:] = 0
sc = torch.from_numpy(sc_np).cuda()
out = torch.zeros(288, dtype=torch.uint8, device="cuda")
bn = build_bn128(0)
bn.lib.ws_set_profiling(bn.ctx.h, 1)
for _ in range(reps):
    bn.ctx.check(bn.lib.ws_mnt6753_g1_multiexp_dev(bn.ctx.h, sc.data_ptr(), pts.data_ptr(), n, out.data_ptr(), 0, None))
    t = bn.ctx.timings()
torch.cuda.synchronize()
print("mnt6753 g1 msm 2^%d:" % lg, " ".join("%s=%.3f" % kv for kv in t), "total=%.3f" % sum(v for _, v in t))
