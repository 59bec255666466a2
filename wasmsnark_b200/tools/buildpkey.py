#!/usr/bin/env python3
"""snarkjs proving_key.json -> proving_key.bin, byte-for-byte what the reference's
tools/buildpkey.js writes (/root/reference/tools/buildpkey.js:50-183):

  10 x u32 header (nVars, nPublic, domainSize, pPolsA, pPolsB, pPointsA, pPointsB1, pPointsB2, pPointsC, pPointsHExps);
  alfa1, beta1, delta1 (G1), beta2, delta2 (G2); polsA, polsB as per-signal lists
  {u32 ncoefs; {u32 idx; coef}}; A, B1 (G1) and B2 (G2) for every signal; C for signals nPublic+1..; hExps[domainSize].
All field elements 8 x u32 little-endian in Montgomery form (R = 2^256).

    python -m wasmsnark_b200.tools.buildpkey -i proving_key.json -o proving_key.bin
"""
import argparse
import json
import struct

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _mq(v):
    return (int(v) << 256) % Q          # toMontgomeryQ, buildpkey.js:57-60


def _mr(v):
    return (int(v) << 256) % R          # toMontgomeryR, buildpkey.js:62-65


def _b(x):
    return x.to_bytes(32, "little")     # writeBigInt, buildpkey.js:50-55


def _g1(p):
    return _b(_mq(p[0])) + _b(_mq(p[1]))                                     # writePoint :67-70


def _g2(p):
    return _b(_mq(p[0][0])) + _b(_mq(p[0][1])) + _b(_mq(p[1][0])) + _b(_mq(p[1][1]))   # writePoint2 :72-77


def _pol(p):
    # writeTransformedPolynomial :79-89.  Object.keys() of a JS object yields integer-like keys in
    # ascending numeric order, whatever the order in the JSON text.
    keys = sorted(p.keys(), key=int)
    out = [struct.pack("<I", len(keys))]
    for k in keys:
        out.append(struct.pack("<I", int(k)) + _b(_mr(p[k])))
    return b"".join(out)


def build(pk):
    n_vars, n_public, domain = pk["nVars"], pk["nPublic"], pk["domainSize"]
    fixed = _g1(pk["vk_alfa_1"]) + _g1(pk["vk_beta_1"]) + _g1(pk["vk_delta_1"]) + _g2(pk["vk_beta_2"]) + _g2(pk["vk_delta_2"])
    pols_a = b"".join(_pol(pk["polsA"][i]) for i in range(n_vars))
    pols_b = b"".join(_pol(pk["polsB"][i]) for i in range(n_vars))
    a = b"".join(_g1(pk["A"][i]) for i in range(n_vars))
    b1 = b"".join(_g1(pk["B1"][i]) for i in range(n_vars))
    b2 = b"".join(_g2(pk["B2"][i]) for i in range(n_vars))
    c = b"".join(_g1(pk["C"][i]) for i in range(n_public + 1, n_vars))
    h = b"".join(_g1(pk["hExps"][i]) for i in range(domain))
    p_pols_a = 40 + len(fixed)
    p_pols_b = p_pols_a + len(pols_a)
    p_a = p_pols_b + len(pols_b)
    p_b1 = p_a + len(a)
    p_b2 = p_b1 + len(b1)
    p_c = p_b2 + len(b2)
    p_h = p_c + len(c)
    hdr = struct.pack("<10I", n_vars, n_public, domain, p_pols_a, p_pols_b, p_a, p_b1, p_b2, p_c, p_h)
    return b"".join((hdr, fixed, pols_a, pols_b, a, b1, b2, c, h))


def main():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("-i", "--input", default="proving_key.json")
    ap.add_argument("-o", "--output", default="proving_key.bin")
    args = ap.parse_args()
    with open(args.input) as f:
        pk = json.load(f)
    with open(args.output, "wb") as f:
        f.write(build(pk))


if __name__ == "__main__":
    main()
