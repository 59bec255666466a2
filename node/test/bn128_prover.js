/* node/test/bn128_prover.js -- the reference's prover test (test/bn128_prover.js:23-80) against this package, for a machine
 * with Node.js:  WASMSNARK_REF=/path/to/wasmsnark npx mocha test     (needs a B200 and the built addon).
 * NOT RUN in the build image (no Node.js); the addon's C code is executed there by tests/test_node_addon_mock.py and the
 * same golden comparisons run through the C ABI in tests/test_gpu_prover.py.
 *
 *   - "It should do a basic point doubling G1": g1_multiexp on the first 1 / 2 / n points equals the reference's result
 *     after affine normalisation (test/bn128_prover.js:25-49 compares g1_multiexp with a naive sum the same way);
 *   - "It should do a basic groth16 proof": groth16GenProof(signals, provingKey) is accepted by groth16Verify
 *     (test/bn128_prover.js:51-80), and with r = s = 0 equals example/bn128/proof_good0.json string for string.
 */
"use strict";
const assert = require("assert");
const fs = require("fs");
const path = require("path");
const { buildBn128 } = require("../index.js");

const REF = process.env.WASMSNARK_REF || path.join(__dirname, "..", "..", "..", "wasmsnark");
const toAB = (b) => b.buffer.slice(b.byteOffset, b.byteOffset + b.byteLength);

describe("Basic tests for groth16 proof generator (wasmsnark_b200)", function () {
    this.timeout(100000);
    let bn128, signals, provingKey;
    before(async () => {
        bn128 = await buildBn128();
        signals = toAB(fs.readFileSync(path.join(REF, "test", "data", "witness.bin")));
        provingKey = toAB(fs.readFileSync(path.join(REF, "test", "data", "proving_key.bin")));
    });
    after(() => bn128.terminate());

    it("g1_multiexp of n points equals the sum of its two halves", async () => {
        const hdr = new Uint32Array(provingKey, 0, 10);
        const nSignals = hdr[0], pPointsA = hdr[5];
        const pts = provingKey.slice(pPointsA, pPointsA + nSignals * 64);
        const n = 1000, h = 437;
        const whole = bn128.g1_affine(await bn128.g1_multiexp(signals.slice(0, n * 32), pts.slice(0, n * 64)));
        const lo = await bn128.g1_multiexp(signals.slice(0, h * 32), pts.slice(0, h * 64));
        const hi = await bn128.g1_multiexp(signals.slice(h * 32, n * 32), pts.slice(h * 64, n * 64));
        // lo + hi through a 2-point multiexp with scalar 1 each (both affine, Montgomery, as key points are)
        const one = new Uint8Array(64); one[0] = 1; one[32] = 1;
        const two = new Uint8Array(128);
        two.set(new Uint8Array(bn128.g1_affine(lo)).subarray(0, 64), 0);
        two.set(new Uint8Array(bn128.g1_affine(hi)).subarray(0, 64), 64);
        const sum = bn128.g1_affine(await bn128.g1_multiexp(one.buffer, two.buffer));
        assert.deepStrictEqual(new Uint8Array(sum), new Uint8Array(whole));
    });

    it("It should do a basic groth16 proof", async () => {
        const proof = await bn128.groth16GenProof(signals, provingKey);
        const vk = JSON.parse(fs.readFileSync(path.join(REF, "test", "data", "verification_key.json"), "utf8"));
        const pub = JSON.parse(fs.readFileSync(path.join(REF, "test", "data", "public.json"), "utf8"));
        assert.strictEqual(await bn128.groth16Verify(vk, pub, proof), true);
        pub[0] = (BigInt(pub[0]) + 1n).toString();
        assert.strictEqual(await bn128.groth16Verify(vk, pub, proof), false);
    });

    it("r = s = 0 reproduces example/bn128/proof_good0.json", async () => {
        const wit = toAB(fs.readFileSync(path.join(REF, "example", "bn128", "witness.bin")));
        const pk = toAB(fs.readFileSync(path.join(REF, "example", "bn128", "proving_key.bin")));
        const zero = new Uint8Array(32).buffer;
        const proof = await bn128.groth16GenProof(wit, pk, zero, zero);
        const good = JSON.parse(fs.readFileSync(path.join(REF, "example", "bn128", "proof_good0.json"), "utf8"));
        for (const k of ["pi_a", "pi_b", "pi_c"]) assert.deepStrictEqual(proof[k], good[k]);
    });

    it("overlapping proofs on one instance, then terminate()", async () => {
        const ps = await Promise.all([0, 1, 2, 3].map(() => bn128.groth16GenProof(signals, provingKey)));
        assert.strictEqual(new Set(ps.map((p) => p.pi_a[0])).size, 4);      // fresh (r, s) each
    });
});
