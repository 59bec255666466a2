#!/usr/bin/env node
/* baseline/run_reference_node.js -- times the UNMODIFIED iden3/wasmsnark (its V8/WASM build, Web-Worker threads) on the
 * same metric / config as bench.py.  NOT MEASURED IN THE BUILD OR GPU IMAGE: neither has Node.js (SURVEY.md 8d);
 * the CPU baseline that bench.py reports is the reference's own WASM image translated to C (oracle/_ref) instead.
 * Run where Node >= 12 and a wasmsnark checkout exist:
 *
 *     node baseline/run_reference_node.js /path/to/wasmsnark [log2 pairs = 14] [proofs = 3]
 *
 * Prints one JSON line per measurement:
 *   g1 multiexp : Bn128.g1_multiexp(scalars, points) -- src/bn128.js:353-383 (workers x g1m_multiexp2, w = 7) on
 *                 n = 2^k pairs cut from the example key's A points + witness (tiled), pairs/s
 *   prove       : Bn128.groth16GenProof(witness, provingKey) on example/bn128 -- src/bn128.js:580-720, proofs/s
 */
"use strict";
const fs = require("fs");
const path = require("path");
const os = require("os");

async function main() {
    const ref = path.resolve(process.argv[2] || ".");
    const log2 = parseInt(process.argv[3] || "14", 10);
    const proofs = parseInt(process.argv[4] || "3", 10);
    const { buildBn128 } = require(path.join(ref, "index.js"));
    const toAB = (b) => b.buffer.slice(b.byteOffset, b.byteOffset + b.byteLength);
    const witness = toAB(fs.readFileSync(path.join(ref, "example", "bn128", "witness.bin")));
    const pkey = toAB(fs.readFileSync(path.join(ref, "example", "bn128", "proving_key.bin")));
    const bn128 = await buildBn128();

    // ---- G1 multiexp on n = 2^log2 pairs (points: the key's A section, tiled; scalars: the witness, tiled)
    const hdr = new Uint32Array(pkey, 0, 10);
    const nVars = hdr[0], pPointsA = hdr[5];
    const n = 1 << log2;
    const scalars = new Uint8Array(n * 32), points = new Uint8Array(n * 64);
    const w8 = new Uint8Array(witness), k8 = new Uint8Array(pkey);
    for (let i = 0; i < n; i++) {
        const j = i % nVars;
        scalars.set(w8.subarray(j * 32, j * 32 + 32), i * 32);
        points.set(k8.subarray(pPointsA + j * 64, pPointsA + j * 64 + 64), i * 64);
    }
    await bn128.g1_multiexp(scalars.buffer.slice(0, 32 * 64), points.buffer.slice(0, 64 * 64));     // warm the workers
    let t0 = process.hrtime.bigint();
    await bn128.g1_multiexp(scalars.buffer, points.buffer);
    let dt = Number(process.hrtime.bigint() - t0) / 1e9;
    console.log(JSON.stringify({ impl: "reference-node", metric: "BN128 G1-MSM scalar-point pairs/sec", n, value: n / dt, unit: "pairs/s",
        seconds: dt, cores: os.cpus().length, node: process.version }));

    // ---- full proofs on the example key
    await bn128.groth16GenProof(witness.slice(0), pkey.slice(0));                                  // warm-up
    t0 = process.hrtime.bigint();
    for (let i = 0; i < proofs; i++) await bn128.groth16GenProof(witness.slice(0), pkey.slice(0));
    dt = Number(process.hrtime.bigint() - t0) / 1e9;
    console.log(JSON.stringify({ impl: "reference-node", metric: "Groth16 proofs/sec (BN128, example/bn128 key, 2^16 constraints)", value: proofs / dt,
        unit: "proofs/s", ms_per_proof: 1000 * dt / proofs, cores: os.cpus().length, node: process.version }));
    bn128.terminate();
}

main().catch((e) => { console.error(e); process.exit(1); });
