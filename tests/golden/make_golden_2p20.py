#!/usr/bin/env python3
"""make_golden_2p20.py -- TEST INFRASTRUCTURE.  Produces tests/golden/synth_2p<k>.golden.json: the proof the
REFERENCE's own code (its WASM image compiled natively, oracle/_ref, driven by oracle/ref_wasm.RefBn128Pool exactly as
src/bn128.js:580-720 drives it) computes for the synthetic 2^k-constraint key of tools/gen_synth_key.py
(make(2^k, 58, 2^k, seed=0xB204) -- BASELINE.json's headline config for k = 20) with the fixed (r, s) that bench.py
and tests/test_gpu_prover.py use.  The key itself (501 MB at k = 20) is not stored: it is regenerated from the seed
wherever the check runs, and its sha256 is pinned here together with the proof.

    python tests/golden/make_golden_2p20.py [k=20] [threads=all]

Runs where /root/reference is mounted (the build container); about 2-3 minutes at k = 20 on 8 cores."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

RS = bytes(range(1, 65))        # r = bytes 1..32, s = bytes 33..64 (raw LE, as bench.py injects them)
SEED = 0xB204
N_PUBLIC = 58


def pages_for(pkey, witness):
    import struct
    hdr = struct.unpack_from("<10I", pkey, 0)
    n_signals, _, domain, pA, pB, pPA = hdr[0], hdr[1], hdr[2], hdr[3], hdr[4], hdr[5]
    need = (1 << 20) + len(witness) + (pB - pA) + (pPA - pB) + n_signals * 32 + domain * 32 * 6
    return max(5000, need // 65536 + 256)


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else len(os.sched_getaffinity(0))
    import gen_synth_key
    from oracle import ref_wasm
    t0 = time.perf_counter()
    pk, wit = gen_synth_key.make(1 << k, N_PUBLIC, 1 << k, seed=SEED)
    t_gen = time.perf_counter() - t0
    pool = ref_wasm.RefBn128Pool(concurrency=threads, pages=pages_for(pk, wit))
    t0 = time.perf_counter()
    proof = pool.groth16GenProof(wit, pk, r=RS[:32], s=RS[32:])
    t_ref = time.perf_counter() - t0
    pool.close()
    out = {
        "what": "reference groth16GenProof (src/bn128.js:580-720 on the reference's compiled WASM image) for gen_synth_key.make(2^%d, %d, 2^%d, seed=0x%X)" % (k, N_PUBLIC, k, SEED),
        "log2_constraints": k, "n_public": N_PUBLIC, "seed": SEED, "r": RS[:32].hex(), "s": RS[32:].hex(),
        "pkey_sha256": hashlib.sha256(pk).hexdigest(), "pkey_bytes": len(pk),
        "witness_sha256": hashlib.sha256(wit).hexdigest(),
        "proof": proof,
        "reference_seconds": round(t_ref, 2), "reference_threads": threads, "keygen_seconds": round(t_gen, 2),
    }
    path = os.path.join(HERE, "synth_2p%d.golden.json" % k)
    json.dump(out, open(path, "w"), indent=1)
    print("wrote %s: reference proof in %.1f s on %d threads (keygen %.1f s)" % (path, t_ref, threads, t_gen))


if __name__ == "__main__":
    main()
