"""Profiling driver: a few Groth16 proofs on a synthetic key (for ncu launch lists / kernel captures).
usage: python tools/run_prove_once.py [log2 constraints] [reps] [tables 0/1] [accumulate variant] [h priority 0/1] [back-to-back loops]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import gen_synth_key  # noqa: E402
from wasmsnark_b200 import build_bn128  # noqa: E402
from wasmsnark_b200._lib import as_ptr  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tables = int(sys.argv[3]) if len(sys.argv) > 3 else 1
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hprio = int(sys.argv[5]) if len(sys.argv) > 5 else 0
loops = int(sys.argv[6]) if len(sys.argv) > 6 else 10
pk, wit = gen_synth_key.make(1 << lg, 58, 1 << lg, seed=0xB204)
bn = build_bn128(0)
key = ctypes.c_void_p()
bn.lib.ws_set_key_tables(bn.ctx.h, tables)
bn.lib.ws_set_msm_acc_variant(bn.ctx.h, variant)
bn.lib.ws_set_h_priority(bn.ctx.h, hprio)
import time
t0 = time.perf_counter()
bn.ctx.check(bn.lib.ws_bn128_pkey_load(bn.ctx.h, as_ptr(pk), len(pk), ctypes.byref(key)))
print('key load %.3f s (tables=%d)' % (time.perf_counter() - t0, tables))
d_w = torch.frombuffer(bytearray(wit), dtype=torch.uint8).cuda()
d_rs = torch.frombuffer(bytearray(bytes(range(1, 65))), dtype=torch.uint8).cuda()
d_out = torch.zeros(384, dtype=torch.uint8, device="cuda")
bn.lib.ws_set_profiling(bn.ctx.h, 1)
for r in range(reps):
    torch.cuda.synchronize()
    bn.ctx.check(bn.lib.ws_bn128_groth16_prove_dev(bn.ctx.h, key, d_w.data_ptr(), d_rs.data_ptr(), d_out.data_ptr(), None))
    t = bn.ctx.timings()
    torch.cuda.synchronize()
bn.lib.ws_set_profiling(bn.ctx.h, 0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for r in range(loops):
    bn.ctx.check(bn.lib.ws_bn128_groth16_prove_dev(bn.ctx.h, key, d_w.data_ptr(), d_rs.data_ptr(), d_out.data_ptr(), None))
torch.cuda.synchronize()
if loops:
    print("h_priority=%d: %.3f ms per proof (%d back to back on the context stream, host clock)" % (hprio, (time.perf_counter() - t0) * 1000 / loops, loops))
proof = bytes(d_out.cpu().numpy()).hex()
print("proof sha:", __import__("hashlib").sha256(bytes.fromhex(proof)).hexdigest()[:16])
print("prove 2^%d variant %d:" % (lg, variant), " ".join("%s=%.3f" % kv for kv in t), "total=%.3f" % sum(v for _, v in t))
