"""Timing driver for the one-process multi-GPU handle (ws_multi_*): blocking stateless proofs of the synthetic key.
usage: python tools/run_multi_once.py [log2 constraints = 20] [ngpus = 0 (all)] [reps = 10] [split: -1 default, 0 uniform, 1 balanced]"""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_synth_key  # noqa: E402
from wasmsnark_b200 import build_bn128, build_bn128_multi  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
pk, wit = gen_synth_key.make(1 << lg, 58, 1 << lg, seed=0xB204)
rs = bytes(range(1, 65))
split = int(sys.argv[4]) if len(sys.argv) > 4 else -1
m = build_bn128_multi(ng)
if split >= 0:
    m.lib.ws_multi_set_split(m.h, split)
t0 = time.perf_counter()
p = m.groth16GenProof(wit, pk, r=rs[:32], s=rs[32:], raw=True)
first = time.perf_counter() - t0
for _ in range(3):
    m.groth16GenProof(wit, pk, r=rs[:32], s=rs[32:], raw=True)
t0 = time.perf_counter()
for _ in range(reps):
    q = m.groth16GenProof(wit, pk, r=rs[:32], s=rs[32:], raw=True)
dt = (time.perf_counter() - t0) / reps
tok = m.load_proving_key(pk)
for _ in range(3):
    m.groth16GenProof(wit, tok, r=rs[:32], s=rs[32:], raw=True)
t0 = time.perf_counter()
for _ in range(reps):
    q2 = m.groth16GenProof(wit, tok, r=rs[:32], s=rs[32:], raw=True)
dt_res = (time.perf_counter() - t0) / reps
print("  resident key: %.2f ms per blocking call (%.1f proofs/s), same bytes: %s" % (1000 * dt_res, 1 / dt_res, q2 == p))
m.terminate()
bn = build_bn128(0)
key = bn.load_proving_key(pk)
want = bn.groth16GenProof(wit, key, r=rs[:32], s=rs[32:], raw=True)
key.free()
bn.terminate()
print("split %d" % split, "ws_multi 2^%d on %d GPU(s): first call %.2f s, then %.2f ms per blocking stateless call (%.1f proofs/s); same bytes as the single-GPU proof: %s; sha %s"
      % (lg, m.ngpus, first, 1000 * dt, 1 / dt, p == want == q, hashlib.sha256(b"".join(p)).hexdigest()[:16]))
