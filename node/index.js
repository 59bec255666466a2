/* node/index.js -- drop-in for the proving half of wasmsnark's npm surface, on libwasmsnark_b200.so.
 *
 *   const { buildBn128 } = require("./node");           // index.js:20-22 of the reference
 *   const bn128 = await buildBn128();
 *   const proof = await bn128.groth16GenProof(witness, provingKey);     // src/bn128.js:580-720
 *
 * and the browser-global convention of main_bn128.js:27-41:
 *   groth16GenProof(witness, provingKey, cb)  ->  cb(err, proof) or a Promise; alias genZKSnarkProof (README.md:28).
 * The returned proof has the reference's shape: decimal strings, pi_a / pi_c = [x, y, "1"],
 * pi_b = [[x0, x1], [y0, y1], ["1", "0"]] (src/bn128.js:714-718).  Not runnable in the build image (no Node).
 */
"use strict";
const native = require("./wasmsnark_b200.node");

function bin2int(buf, off) {            // src/bn128.js:319-327, with BigInt instead of big-integer
    const v = new Uint32Array(buf, off, 8);
    let acc = 0n;
    for (let i = 7; i >= 0; i--) acc = (acc << 32n) | BigInt(v[i]);
    return acc.toString();
}
const bin2g1 = (b) => [bin2int(b, 0), bin2int(b, 32), bin2int(b, 64)];                                   // :329-335
const bin2g2 = (b) => [[bin2int(b, 0), bin2int(b, 32)], [bin2int(b, 64), bin2int(b, 96)], [bin2int(b, 128), bin2int(b, 160)]];   // :336-351

function int2bin(v) {                   // setInt, src/bn128.js:455-470: decimal string / BigInt -> 32 bytes LE (plain form)
    let x = BigInt(v);
    if (x < 0n || x >> 256n) throw new RangeError("field element out of range");
    const out = new Uint8Array(32);
    for (let i = 0; i < 32; i++) { out[i] = Number(x & 0xffn); x >>= 8n; }
    return out;
}
const cat = (parts) => { const o = new Uint8Array(parts.reduce((n, p) => n + p.length, 0)); let k = 0; for (const p of parts) { o.set(p, k); k += p.length; } return o; };
const g1bin = (e) => cat([int2bin(e[0]), int2bin(e[1])]);                                                // setG1Affine :503-507
const g2bin = (e) => cat([int2bin(e[0][0]), int2bin(e[0][1]), int2bin(e[1][0]), int2bin(e[1][1])]);      // setG2Affine :509-513

// g?m_fromMontgomery (src/build_curve_jacobian_a0.js:258-277): every 32-byte Fq coordinate x -> x R^-1 mod q, nothing else
const BN128_Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583n;
const BN128_RINV = 20988524275117001072002809824448087578619730785600314334253784976379291040311n;      // (2^256)^-1 mod q
function fromMontgomery(p, size) {
    const b = p instanceof ArrayBuffer ? new Uint8Array(p) : new Uint8Array(p.buffer, p.byteOffset, p.byteLength);
    if (b.length === 0 || b.length % size) throw new RangeError("whole " + size + "-byte points expected");
    const out = new Uint8Array(b.length);
    for (let o = 0; o < b.length; o += 32) {
        let x = 0n;
        for (let i = 31; i >= 0; i--) x = (x << 8n) | BigInt(b[o + i]);
        out.set(int2bin((x * BN128_RINV) % BN128_Q), o);
    }
    return out.buffer;
}

class Bn128 {
    // opts: a device index, or { device, gpus } -- gpus > 1 (0 = all) spreads every proof / multiexp over that many GPUs
    constructor(opts) {
        const o = typeof opts === "object" && opts !== null ? opts : { device: opts };
        this.h = native.create(o.device || 0, o.gpus === undefined ? 1 : o.gpus);
    }
    groth16GenProof(signals, pkey, r, s) {
        return new Promise((resolve, reject) => {
            native.prove(this.h, signals, pkey, r || null, s || null, (err, p) => {
                if (err) return reject(err);
                resolve({ pi_a: bin2g1(p.pi_a), pi_b: bin2g2(p.pi_b), pi_c: bin2g1(p.pi_c) });
            });
        });
    }
    async g1_multiexp(scalars, points) { return native.g1Multiexp(this.h, scalars, points); }
    async g2_multiexp(scalars, points) { return native.g2Multiexp(this.h, scalars, points); }
    g1_affine(p) { return native.g1Affine(this.h, p); }
    g2_affine(p) { return native.g2Affine(this.h, p); }
    g1_fromMontgomery(p) { return fromMontgomery(p, 96); }             // src/bn128.js:429-433
    g2_fromMontgomery(p) { return fromMontgomery(p, 192); }            // :435-439
    async calcH(signals, polsA, polsB, nSignals, domainSize) { return native.calcH(this.h, signals, polsA, polsB, nSignals, domainSize); }
    async groth16Verify(verificationKey, input, proof) {                // src/bn128.js:722-786
        if (input === undefined || input === null) input = [];
        else if (!Array.isArray(input)) input = [input];
        let vk, ic, ins, pr;
        try {
            vk = cat([g1bin(verificationKey.vk_alfa_1), g2bin(verificationKey.vk_beta_2), g2bin(verificationKey.vk_gamma_2), g2bin(verificationKey.vk_delta_2)]);
            ic = cat(verificationKey.IC.map(g1bin));
            ins = cat(input.map(int2bin));
            pr = cat([g1bin(proof.pi_a), g2bin(proof.pi_b), g1bin(proof.pi_c)]);
        } catch (e) { if (e instanceof RangeError) return false; throw e; }
        return native.verify(vk.buffer, ic.buffer, ins.buffer, pr.buffer);
    }
    bin2g1(b) { return bin2g1(b); }
    bin2g2(b) { return bin2g2(b); }
    terminate() { native.destroy(this.h); this.h = null; }
}

async function buildBn128(opts) { return new Bn128(opts); }

let shared = null;
function groth16GenProof(witness, provingKey, cb) {
    if (!shared) shared = new Bn128(0);
    const p = shared.groth16GenProof(witness, provingKey);
    if (cb) p.then((proof) => cb(null, proof), (err) => cb(err));
    else return p;
}

// snarkjs proof.json as shipped in example/bn128/proof_good.json
const proofJson = (p) => JSON.stringify({ pi_a: p.pi_a, pi_b: p.pi_b, pi_c: p.pi_c, protocol: "groth" }, null, 1);

module.exports = { buildBn128, groth16GenProof, genZKSnarkProof: groth16GenProof, proofJson };
