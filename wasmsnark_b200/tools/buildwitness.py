#!/usr/bin/env python3
"""snarkjs witness.json -> witness.bin exactly as the reference's tools/buildwitness.js does
(/root/reference/tools/buildwitness.js:36-69): nVars x 32 bytes, little-endian, NOT Montgomery, no header.

    python -m wasmsnark_b200.tools.buildwitness -i witness.json -o witness.bin
"""
import argparse
import json


def build(witness):
    return b"".join(int(w).to_bytes(32, "little") for w in witness)


def main():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("-i", "--input", default="witness.json")
    ap.add_argument("-o", "--output", default="witness.bin")
    args = ap.parse_args()
    with open(args.input) as f:
        w = json.load(f)
    with open(args.output, "wb") as f:
        f.write(build(w))


if __name__ == "__main__":
    main()
