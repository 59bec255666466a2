/* wasmsnark_b200.h -- C ABI of libwasmsnark_b200.so: the B200 (sm_100a) replacement for the
 * data-parallel proving path of iden3/wasmsnark.
 *
 * Every entry point below states the reference interface it replaces (paths relative to the
 * reference checkout).  In the reference these are WebAssembly exports taking i32 offsets into
 * one linear memory, driven by the JS host src/bn128.js; here they take plain host (or,
 * for the *_dev variants, device) pointers and sizes.  All buffers use the reference's byte
 * layouts unchanged: field elements are 8 (BN128) or 24 (MNT6753) 32-bit little-endian limbs;
 * key points / coefficients are in Montgomery form (R = 2^256 resp. 2^768), witness scalars
 * are plain and are consumed as raw 256-bit integers (never reduced).
 *
 * Return value: 0 on success, a negative WS_E_* code otherwise (ws_last_error() gives text).
 * The proving path never falls back to a CPU implementation: without a usable CUDA device
 * ws_create fails with WS_E_CUDA.  Only the Groth16 verifier / pairing entry points are host code, by design
 * (four pairings are not data-parallel work); they take no ws_ctx.
 *
 * Threading: a ws_ctx owns one CUDA device, one stream set and one workspace and is thread-safe:
 * every entry point holds the context's mutex for the duration of the call, so blocking calls from
 * several host threads queue up like the reference's actionQueue (src/bn128.js:272-300), and the
 * asynchronous *_dev calls -- which share the context's workspace -- are additionally chained on the
 * device (work enqueued on another stream than the previous call's first waits for it).  Use one
 * ctx per stream of work that should run concurrently.  Different ctx objects are independent.
 */
#ifndef WASMSNARK_B200_H
#define WASMSNARK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ws_ctx ws_ctx;
typedef struct ws_pkey ws_pkey;   /* a proving key resident in HBM (see ws_bn128_pkey_load) */

enum {
    WS_OK = 0,
    WS_E_ARG = -1,     /* null pointer, bad length, bad header, misaligned device pointer */
    WS_E_SIZE = -2,    /* FFT size not a power of two or > 2^28 (reference traps: src/build_fft.js:137-154) */
    WS_E_CUDA = -3,    /* CUDA runtime error (no device, out of memory, launch failure) */
    WS_E_FIELD = -4    /* unknown field id */
};

/* field ids for the batch field operations */
enum {
    WS_BN128_FQ = 0,   /* f1m_*  : src/bn128/build_bn128.js:33 (q, :19)  */
    WS_BN128_FR = 1,   /* frm_*  : src/bn128/build_bn128.js:34 (r, :20)  */
    WS_MNT6753_FQ = 2, /* f1m_* of the MNT6753 image: src/mnt6753/build_mnt6753.js:19 */
    WS_BLS12381_FQ = 3, /* f1m_* of the BLS12-381 module: src/bls12381/build_bls12381.js:20,33 (48-byte elements) */
    WS_BLS12381_FR = 4  /* frm_* of the BLS12-381 module: src/bls12381/build_bls12381.js:21,34 */
};

/* ---- context -------------------------------------------------------------------------------
 * Replaces buildBn128() (src/bn128.js:173-265): instead of 8 workers with a 320 MB WASM memory
 * each, one CUDA context on `device` with a growable HBM workspace.                           */
int ws_create(int device, ws_ctx** out);
void ws_destroy(ws_ctx* ctx);                       /* Bn128.terminate(), src/bn128.js:562-566 */
/* Text of the CALLING THREAD's last failed call on this context (valid until that thread's next failure); a thread that
 * never failed on it reads the context's most recent failure.  So several threads sharing a context each see their own
 * message, whatever the others did in between.  ws_multi_last_error behaves the same way. */
const char* ws_last_error(const ws_ctx* ctx);
const char* ws_version(void);

/* ---- field arithmetic (config C1 / C5) -----------------------------------------------------
 * out[i] = x[i] (op) y[i] for n elements; x, y, out are n*32 bytes (n*96 for MNT6753, n*48 for BLS12-381 Fq).
 * Replaces f1m_mul / frm_mul        src/build_f1m.js:235-436
 *          f1m_add, f1m_sub         src/build_f1m.js:67-101
 *          f1m_square               src/build_f1m.js:439-737      (y ignored)
 *          f1m_toMontgomery / fromMontgomery  src/build_f1m.js:749-770   (y ignored)
 *          f1m_inverse              src/build_f1m.js:772-783      (y ignored)
 * and the batch forms fft_toMontgomeryN / fromMontgomeryN / mulN  src/build_fft.js:418-547.
 * Operands are field elements (< p) -- except for WS_OP_TOMONT, which like f1m_toMontgomery takes ANY number that fits
 * the element's bytes and returns the canonical x R mod p (the witness enters the prover that way, src/bn128.js:134). */
enum { WS_OP_MUL = 0, WS_OP_ADD = 1, WS_OP_SUB = 2, WS_OP_SQUARE = 3, WS_OP_TOMONT = 4, WS_OP_FROMMONT = 5, WS_OP_INVERSE = 6 };
int ws_fp_batch(ws_ctx* ctx, int field, int op, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n);

/* Fq2 (BN128, u^2 = -1) batch multiply: replaces f2m_mul src/build_f2m.js:127-163; 64-byte elements. */
int ws_bn128_fq2_mul_batch(ws_ctx* ctx, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n);

/* ---- multi-scalar multiplication (config C2) -------------------------------------------------
 * Replaces Bn128.g1_multiexp(scalars, points)  src/bn128.js:353-383
 *          -> worker G1_MULTIEXP               src/bn128.js:102-113
 *          -> g1m_multiexp2(ps, pp, n, 7, pr)  src/build_multiexp.js:651-744
 * scalars: n*32 bytes raw LE; points: n*64 bytes affine Montgomery (x == 0 means infinity,
 * src/build_multiexp.js:335-349).  out: 96 bytes X|Y|Z Montgomery, an UN-NORMALISED Jacobian triple as
 * in the reference (callers normalise with g1_affine, src/bn128.js:706); which representative of
 * the point comes back is unspecified there and here -- compare after ws_bn128_g1_affine, the way
 * the reference's own test does (test/bn128_prover.js:31-46).                                     */
int ws_bn128_g1_multiexp(ws_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[96]);

/* Replaces Bn128.g2_multiexp  src/bn128.js:385-415 -> G2_MULTIEXP :114-125 -> g2m_multiexp
 * src/build_multiexp.js:498-580.  points: n*128 bytes (x.c0|x.c1|y.c0|y.c1); out: 192 bytes.   */
int ws_bn128_g2_multiexp(ws_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[192]);

/* Replaces Bn128.g1_affine / g2_affine (src/bn128.js:417-427 -> g?m_affine,
 * src/build_curve_jacobian_a0.js:424-457) and, with from_montgomery != 0, the following
 * g?_fromMontgomery (src/bn128.js:429-439): `count` Jacobian triples in, (x, y, 1) out, infinity
 * becomes (0, 1, 0).  96-byte (G1) / 192-byte (G2) records.                                      */
int ws_bn128_g1_affine(ws_ctx* ctx, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery);
int ws_bn128_g2_affine(ws_ctx* ctx, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery);

/* Device-resident variants: d_scalars / d_points / d_out are device pointers (16-byte aligned)
 * on the ctx's device, work is enqueued on `stream` (a cudaStream_t, may be NULL = ctx stream)
 * and the call returns without synchronising.  d_out receives the same 96 / 192 bytes.
 * window_bits = 0 picks the window size from n.                                                */
int ws_bn128_g1_multiexp_dev(ws_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, void* d_out, int window_bits, void* stream);
int ws_bn128_g2_multiexp_dev(ws_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, void* d_out, int window_bits, void* stream);

/* Window tables for point sets that stay in HBM across many MSMs (what resident proving keys use internally):
 * build_table writes ws_msm_table_windows(table_c) * n affine points, table[w * n + i] = 2^(table_c * w) * P_i;
 * multiexp_table then feeds every digit window into one shared bucket set -- wider windows, no doublings.
 * table_c in 8..22; result format as ws_bn128_g?_multiexp_dev.                                             */
int ws_msm_table_windows(int table_c);
int ws_bn128_g1_build_table_dev(ws_ctx* ctx, const void* d_points, size_t n, int table_c, void* d_table, void* stream);
int ws_bn128_g2_build_table_dev(ws_ctx* ctx, const void* d_points, size_t n, int table_c, void* d_table, void* stream);
int ws_bn128_g1_multiexp_table_dev(ws_ctx* ctx, const void* d_scalars, const void* d_table, size_t n, int table_c, void* d_out, void* stream);
int ws_bn128_g2_multiexp_table_dev(ws_ctx* ctx, const void* d_scalars, const void* d_table, size_t n, int table_c, void* d_out, void* stream);

/* Multi-GPU combine: d_parts holds `count` partial sums in the 96-byte (G1) / 192-byte (G2) result
 * format above (e.g. all-gathered over NCCL from the ranks that each took a contiguous point range,
 * as the reference's workers do, src/bn128.js:353-372); d_out receives their normalised sum.
 * Replaces the serial g1m_add loop over worker results, src/bn128.js:376-380 / :408-412.          */
int ws_bn128_g1_sum_dev(ws_ctx* ctx, const void* d_parts, int count, void* d_out, void* stream);
int ws_bn128_g2_sum_dev(ws_ctx* ctx, const void* d_parts, int count, void* d_out, void* stream);

/* ---- MNT6753 wide-limb path (config C5): 24-limb (768-bit) Montgomery field -------------------------
 * ws_fp_batch(field = WS_MNT6753_FQ) covers f1m_* of the MNT6753 image.
 * Fq3 = Fq[u]/(u^3 - 11) product: replaces f3m_mul  src/build_f3m.js:125-189 (non-residue 11,
 * src/mnt6753/build_mnt6753.js:35-56); elements c0|c1|c2, 288 bytes.                                  */
int ws_mnt6753_fq3_mul_batch(ws_ctx* ctx, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n);
/* G1 multi-scalar multiplication on y^2 = x^3 + 11x + b (curve module src/build_curve_projective.js with
 * mulByA = 11, src/mnt6753/build_mnt6753.js:90-109).  The reference has NO multiexp for this curve; the
 * result equals sum_i g1m_timesScalar(P_i, s_i) folded with g1m_add.  scalars: n*96 bytes raw LE; points:
 * n*192 bytes affine Montgomery x|y (x == 0 => infinity, as in the BN128 path); out: 288 bytes, a
 * homogeneous projective triple X|Y|Z (x = X/Z, y = Y/Z) as the reference's g1m_* use; infinity = Z 0.  */
int ws_mnt6753_g1_multiexp(ws_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[288]);
int ws_mnt6753_g1_multiexp_dev(ws_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, void* d_out, int window_bits, void* stream);
/* g1m_affine (src/build_curve_projective.js:624-652) [+ g1m_fromMontgomery]: projective triples -> (x, y, 1). */
int ws_mnt6753_g1_affine(ws_ctx* ctx, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery);

/* ---- BLS12-381: the same kernels instantiated for the reference's bls12381 module --------
 * src/bls12381/build_bls12381.js:20-43 builds f1m (q, 6 x 64 bit), frm (r), g1m / g2m (a = 0 Jacobian curves,
 * src/build_curve_jacobian_a0.js), g1m_multiexp / g2m_multiexp (src/build_multiexp.js), fft over frm
 * (src/build_fft.js) and f2m = f1m[u]/(u^2+1) (src/build_f2m.js).  Conventions as for BN128: 32-byte raw LE scalars,
 * affine Montgomery points (G1 x|y = 96 bytes, G2 = 192 bytes; x == 0 => infinity), Jacobian results X|Y|Z (G1 144,
 * G2 288 bytes; compare after *_affine), fft in place on 32-byte Montgomery Fr elements, n a power of two <= 2^28.
 * ws_fp_batch(field = WS_BLS12381_FQ / WS_BLS12381_FR) covers f1m_* / frm_*.  The reference ships no compiled image
 * and no host for this module, so these are checked against the big-integer oracle only. */
int ws_bls12381_fq2_mul_batch(ws_ctx* ctx, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n);
int ws_bls12381_g1_multiexp(ws_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[144]);
int ws_bls12381_g1_multiexp_dev(ws_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, void* d_out, int window_bits, void* stream);
int ws_bls12381_g1_affine(ws_ctx* ctx, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery);
int ws_bls12381_g2_multiexp(ws_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[288]);
int ws_bls12381_g2_multiexp_dev(ws_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, void* d_out, int window_bits, void* stream);
int ws_bls12381_g2_affine(ws_ctx* ctx, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery);
int ws_bls12381_fr_fft(ws_ctx* ctx, uint8_t* px, size_t n, int odd);
int ws_bls12381_fr_ifft(ws_ctx* ctx, uint8_t* px, size_t n, int odd);
int ws_bls12381_fr_fft_dev(ws_ctx* ctx, void* d_px, size_t n, int odd, int inverse, void* stream);

/* ---- NTT over Fr (config C3) -------------------------------------------------------------------
 * Replaces fft_fft(px, n, odd) / fft_ifft(px, n, odd)   src/build_fft.js:159-221 (+ __rawfft :223-372,
 * __finalInverse :550-648): in place on n 32-byte Montgomery Fr elements, natural order in and out.
 * odd != 0 evaluates on the odd 2n-th roots of unity (twiddle starts at w_2m).  n must be a power of
 * two <= 2^28, else WS_E_SIZE (the reference traps, src/build_fft.js:137-154).                     */
int ws_bn128_fr_fft(ws_ctx* ctx, uint8_t* px, size_t n, int odd);
int ws_bn128_fr_ifft(ws_ctx* ctx, uint8_t* px, size_t n, int odd);
/* device-resident form: d_px on the ctx's device, enqueued on `stream`, no synchronisation */
int ws_bn128_fr_fft_dev(ws_ctx* ctx, void* d_px, size_t n, int odd, int inverse, void* stream);

/* ---- H polynomial (worker command CALC_H) ----------------------------------------------------------
 * Replaces Bn128.calcH(signals, polsA, polsB, nSignals, domainSize)  src/bn128.js:569-578 -> :126-166.
 * signals: nSignals*32 bytes plain LE; polsA / polsB: the key's sections (per signal: u32 ncoefs, then
 * {u32 idx, 32-byte Montgomery coef} x ncoefs; tools/buildpkey.js:79-89); out: domainSize*32 bytes,
 * the coefficients n..2n-1 of A*B, out of Montgomery form.                                          */
int ws_bn128_calc_h(ws_ctx* ctx, const uint8_t* signals, const uint8_t* polsA, size_t lenA, const uint8_t* polsB, size_t lenB,
                    uint32_t n_signals, uint32_t domain_size, uint8_t* out);

/* ---- Groth16 prover (config C4) -------------------------------------------------------------------
 * Replaces Bn128.groth16GenProof(signals, pkey)  src/bn128.js:580-720 (public surface main_bn128.js:27-41).
 * witness: nVars*32 bytes plain LE (tools/buildwitness.js:36-69); pkey: proving_key.bin
 * (tools/buildpkey.js:130-183).  r, s: 32 raw bytes each; NULL draws them from the OS CSPRNG like
 * the reference (src/bn128.js:642-664); all-zero r and s is the reference's debug mode (:666-668)
 * pinned by example/bn128/proof_good0.json.
 * Outputs are what the reference formats into decimal strings (src/bn128.js:714-718): affine and out of
 * Montgomery form -- pi_a, pi_c: x|y|z (96 bytes), pi_b: x.c0|x.c1|y.c0|y.c1|z.c0|z.c1 (192 bytes);
 * z is 1 (or the point is (0,1,0) at infinity).                                                      */
int ws_bn128_groth16_prove(ws_ctx* ctx, const uint8_t* witness, size_t witness_len, const uint8_t* pkey, size_t pkey_len,
                           const uint8_t* r, const uint8_t* s, uint8_t pi_a[96], uint8_t pi_b[192], uint8_t pi_c[96]);

/* ---- Groth16 verifier and pairing, on the HOST (no ws_ctx, no GPU: four pairings are not data-parallel work).
 * Replaces Bn128.groth16Verify src/bn128.js:722-786 -> bn128_pairingEq4 src/bn128/build_bn128.js:1287-1318, and the
 * bn128_pairing export (:1320-1346; prepareG1/G2 :265-575, millerLoop :756-836, finalExponentiation :943-1262).
 * Every field element crosses as a PLAIN (non-Montgomery) 32-byte little-endian integer, i.e. the decimal strings of
 * verification_key.json / proof.json / public.json; G1 = x|y, G2 = x.c0|x.c1|y.c0|y.c1 (the JSON's [[x0,x1],[y0,y1]]),
 * all-zero coordinates = point at infinity.
 * ws_bn128_groth16_verify: ic = IC[0..n_ic-1] (64 B each), n_ic must be n_inputs + 1 (else WS_E_ARG); *valid = 1 iff
 * e(A,B) e(-vk_x,gamma2) e(-C,delta2) e(-alfa1,beta2) = 1.  An input >= r gives *valid = 0 like the reference (:772);
 * a coordinate >= q or a point off its curve also gives *valid = 0 (the reference does not check and computes an
 * unspecified value).
 * ws_bn128_pairing: out = e(P, Q) as 12 plain Fq values in the reference's ftm order (c0.a0, c0.a1, c0.a2, c1.a0,
 * c1.a1, c1.a2, each re|im) -- bit-identical to bn128_pairing + ftm_fromMontgomery.  WS_E_ARG for malformed points.
 * The Miller loop walks the twist point in affine coordinates, which has no formula for a vertical line (T = +-Q in an
 * addition step): impossible for Q in the order-r subgroup, possible for a twist point outside it.  Such a step is
 * detected (never fed into the shared batch inversion): ws_bn128_pairing returns WS_E_ARG, the verifier *valid = 0. */
int ws_bn128_groth16_verify(const uint8_t vk_alfa_1[64], const uint8_t vk_beta_2[128], const uint8_t vk_gamma_2[128],
                            const uint8_t vk_delta_2[128], const uint8_t* ic, size_t n_ic, const uint8_t* inputs, size_t n_inputs,
                            const uint8_t pi_a[64], const uint8_t pi_b[128], const uint8_t pi_c[64], int* valid);
int ws_bn128_pairing(const uint8_t p[64], const uint8_t q[128], uint8_t out[384]);

/* The stateless call above keeps the parsed / uploaded key of the previous call (one entry per ctx) and reuses it
 * when the next call's key buffer has the same length and 64-bit content fingerprint (the buffer is hashed on
 * every call, ~10 GB/s); a different key replaces the entry.  on = 0 disables the cache and drops the entry. */
int ws_set_key_cache(ws_ctx* ctx, int on);

/* Resident proving key: parses and uploads proving_key.bin once (sections to HBM, the two polynomial
 * sections regrouped per row); the reference re-slices and re-posts the key on every proof
 * (src/bn128.js:592-604).  The handle belongs to `ctx`.                                              */
int ws_bn128_pkey_load(ws_ctx* ctx, const uint8_t* pkey, size_t pkey_len, ws_pkey** out);
/* Resident keys additionally get window tables in HBM (2^(c w) * P_i for every point section, c up to 20 bits:
 * about 13x the key's point bytes): all digit windows of an MSM then share one bucket set and the result needs no
 * doublings.  on = 0 disables them for keys loaded afterwards (same results, table-free MSM).  Default on.       */
int ws_set_key_tables(ws_ctx* ctx, int on);
/* Tuning hook: on != 0 runs the prover's one dependent chain (constructLC -> five NTTs -> H MSM; the reference's
 * CALC_H task and the multiexp that waits for it, src/bn128.js:607-622) on an internal high-priority stream, forked
 * from and joined to the caller's stream, so the four witness MSMs on the side streams do not stretch it.  Same
 * results.  Measured at 2^20 constraints: the chain ends 3x earlier but the proof takes the same 21.5 ms (the GPU
 * is throughput-bound), so the default is off. */
int ws_set_h_priority(ws_ctx* ctx, int on);
void ws_bn128_pkey_free(ws_pkey* key);
int ws_bn128_pkey_info(const ws_pkey* key, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size);
int ws_bn128_groth16_prove_key(ws_ctx* ctx, const ws_pkey* key, const uint8_t* witness, size_t witness_len,
                               const uint8_t* r, const uint8_t* s, uint8_t pi_a[96], uint8_t pi_b[192], uint8_t pi_c[96]);
/* d_witness: nVars*32 bytes on the device; d_rs: r|s, 64 bytes on the device; d_out: 384 bytes
 * (pi_a | pi_b | pi_c) on the device; enqueued on `stream`, no synchronisation.                      */
int ws_bn128_groth16_prove_dev(ws_ctx* ctx, const ws_pkey* key, const void* d_witness, const void* d_rs, void* d_out, void* stream);

/* ---- multi-GPU prover (one process / one ctx per GPU) --------------------------------------------------
 * The reference shards every multiexp over its workers by contiguous point ranges and adds the partial
 * points (src/bn128.js:353-383).  Same split here across ranks: ws_bn128_pkey_load_shard keeps only rank's
 * ranges of A, B1, B2, C, hExps in HBM (plus the polynomial sections: H is computed on every rank);
 * ws_bn128_groth16_partials_dev writes the rank's five partial sums (768 bytes: XYZZ records A|B1|C|H|B2);
 * the caller all-gathers the `world` records (NCCL) and every rank finishes with
 * ws_bn128_groth16_finish_dev, which adds them and runs the final assembly.                             */
#define WS_PARTIALS_BYTES 768
int ws_bn128_pkey_load_shard(ws_ctx* ctx, const uint8_t* pkey, size_t pkey_len, int rank, int world, ws_pkey** out);
int ws_bn128_groth16_partials_dev(ws_ctx* ctx, const ws_pkey* key, const void* d_witness, void* d_partials, void* stream);
/* The same, split so that H (the reference's single CALC_H task, src/bn128.js:569-578) is computed by ONE rank per
 * proof and handed to the others -- what strong scaling over the GPUs of a box needs (every rank repeating the five
 * NTTs caps the speed-up):
 *   ws_bn128_groth16_witness_msms_dev  starts the rank's A, B1, B2, C partial MSMs on the context's side streams, forked
 *        from `stream` at this point; d_witness is an nVars*32-byte device buffer of which only the ranges
 *        ws_bn128_pkey_shard_info reports need to hold data.  Writes slots A, B1, C, B2 of d_partials.
 *   ws_bn128_groth16_h_dev             h = coefficients n..2n-1 of A*B (domainSize*32 bytes, plain form) from the WHOLE
 *        witness, on `stream` -- run by the proof's H owner only; the caller then ships h[h_lo, h_lo + h_cnt) to every
 *        rank (NCCL send/recv, or a peer copy inside one process: see ws_multi_* below).
 *   ws_bn128_groth16_h_msm_dev         the rank's H partial MSM over its slice d_h_shard (h_cnt*32 bytes), on `stream`,
 *        then joins the side streams: afterwards d_partials is complete in stream order (all_gather + finish as above).
 * ws_bn128_pkey_shard_info: range = {v_lo, v_cnt, c_lo, c_cnt, h_lo, h_cnt} -- witness element ranges [v_lo, v_lo+v_cnt)
 * (A/B1/B2; their hull if they differ) and [c_lo, c_lo+c_cnt) (C; c_lo already includes the nPublic+1 offset), h range
 * [h_lo, h_lo+h_cnt).                                                                                                */
int ws_bn128_groth16_witness_msms_dev(ws_ctx* ctx, const ws_pkey* key, const void* d_witness, void* d_partials, void* stream);
int ws_bn128_groth16_h_dev(ws_ctx* ctx, const ws_pkey* key, const void* d_witness, void* d_h, void* stream);
int ws_bn128_groth16_h_msm_dev(ws_ctx* ctx, const ws_pkey* key, const void* d_h_shard, void* d_partials, void* stream);
int ws_bn128_pkey_shard_info(const ws_pkey* key, uint32_t range[6]);
/* A partial key with an explicit element range per point section: ranges = {lo, cnt} x 5 in the order A, B1, B2, C, H (A, B1,
 * B2 index the signals, C the signals after the public ones, H the domain).  Any partition of the points over the ranks
 * gives the same proof (the sums are associative); wasmsnark_b200/parallel.py (balanced_ranges) and ws_multi_* use it to
 * give a rank one or two LARGE contiguous pieces -- cut from the five sections laid end to end, weighted by cost per
 * point -- instead of 1/N of every section: fewer, larger MSMs per rank (wider windows, fewer reduction tails).
 * Accepted by the partials / split entry points above, like a ws_bn128_pkey_load_shard key.                          */
int ws_bn128_pkey_load_ranges(ws_ctx* ctx, const uint8_t* pkey, size_t pkey_len, const uint32_t ranges[10], ws_pkey** out);
/* The balanced partition of a key over n ranks (host code, no GPU): out receives n x 10 values, rank-major, each row in the
 * layout ws_bn128_pkey_load_ranges takes.  WS_E_ARG for a malformed key header or n outside 1..16.                   */
int ws_balanced_ranges(const uint8_t* pkey, size_t pkey_len, int n, uint32_t* out);
/* optional, before partials: starts the products that depend only on (r, s) and the key's fixed points
 * (r*delta1, s*delta1, (r*s)*delta1, s*delta2; src/bn128.js:672-702) on a side stream so that they overlap
 * the multi-scalar multiplications; finish_dev picks the result up (or computes it itself if not prepared). */
int ws_bn128_groth16_prepare_dev(ws_ctx* ctx, const ws_pkey* key, const void* d_rs, void* stream);
int ws_bn128_groth16_finish_dev(ws_ctx* ctx, const ws_pkey* key, const void* d_all_partials, int world, const void* d_rs, void* d_out, void* stream);

/* ---- all the GPUs of one box behind one handle (one process, one blocking call) -------------------------------
 * Replaces buildBn128() + the worker pool as a whole (src/bn128.js:173-265): ws_multi_create(ngpus) opens one ws_ctx
 * per GPU (ngpus <= 0: every visible device) with peer access between them; the calls below are the reference's
 * Bn128.groth16GenProof / g1_multiexp / g2_multiexp with their worker split (:353-415, :580-720) mapped onto the
 * GPUs: contiguous point ranges per GPU, H (the single CALC_H task, :569-578) computed by one GPU per proof -- rotating
 * from proof to proof -- and shipped slice-wise to the others with peer copies over NVLink, the partial points copied
 * to GPU 0, added and assembled there.  Host buffers, blocking, thread-safe (calls on one handle are serialised).
 * ws_multi_bn128_groth16_prove is stateless like ws_bn128_groth16_prove: the key's shards stay on the GPUs between
 * calls and are reused when the caller's key buffer has the same length and content fingerprint (hashed on the host
 * while the GPUs already work on the proof).  Results are bit-identical to the single-GPU calls.              */
typedef struct ws_multi ws_multi;
int ws_multi_create(int ngpus, ws_multi** out);
void ws_multi_destroy(ws_multi* m);
int ws_multi_gpus(const ws_multi* m);
/* Point partition for keys loaded afterwards: 0 = uniform (1/N of every section, the reference's worker split,
 * src/bn128.js:353-361; the default), 1 = balanced (see ws_bn128_pkey_load_ranges).  Same proofs. */
int ws_multi_set_split(ws_multi* m, int mode);
const char* ws_multi_last_error(const ws_multi* m);
int ws_multi_bn128_groth16_prove(ws_multi* m, const uint8_t* witness, size_t witness_len, const uint8_t* pkey, size_t pkey_len,
                                 const uint8_t* r, const uint8_t* s, uint8_t pi_a[96], uint8_t pi_b[192], uint8_t pi_c[96]);
/* Resident form (no per-call fingerprint of the key bytes -- at 8 GPUs hashing a 500 MB key takes longer than the proof):
 * ws_multi_bn128_pkey_load makes `pkey` the handle's resident key (its shards replace the previous key's);
 * ws_multi_bn128_groth16_prove_resident proves with it.                                                              */
int ws_multi_bn128_pkey_load(ws_multi* m, const uint8_t* pkey, size_t pkey_len);
int ws_multi_bn128_groth16_prove_resident(ws_multi* m, const uint8_t* witness, size_t witness_len, const uint8_t* r, const uint8_t* s,
                                          uint8_t pi_a[96], uint8_t pi_b[192], uint8_t pi_c[96]);
int ws_multi_bn128_g1_multiexp(ws_multi* m, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[96]);
int ws_multi_bn128_g2_multiexp(ws_multi* m, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[192]);

/* Number of this library's CUDA kernels launched on the context so far (MSM, NTT and prover paths). */
unsigned long long ws_kernel_launches(const ws_ctx* ctx);

/* Stage timing on/off (CUDA events recorded between the pipeline's kernels on the launch stream). */
int ws_set_profiling(ws_ctx* ctx, int on);
/* Tuning/test hook: force the per-task entry cap of the MSM accumulate kernel (0 = automatic). */
int ws_set_msm_task_cap(ws_ctx* ctx, unsigned cap);
/* Tuning/test hook: buckets per segment in the bucket-reduction step (power of two, default 32). */
int ws_set_msm_seg_len(ws_ctx* ctx, unsigned len);
/* Tuning/test hook: occupancy variant (launch bounds) of the bucket-accumulation kernel; 0 = default.  Bits 0..3 select the
 * variant for G1-type curves (1: 128 threads x 5 blocks, 2: 128 x 4, 3: 192 x 3), bits 4..7 for G2-type curves (1: 64
 * threads x 3 blocks, 2: 128 x 3).  The environment variables WS_MSM_ACC_VARIANT, WS_MSM_SEG_LEN, WS_MSM_TASK_CAP and
 * WS_NTT_VARIANT preset the corresponding hooks of every context a process creates (A/B runs of whole programs). */
int ws_set_msm_acc_variant(ws_ctx* ctx, int variant);
/* Tuning/test hook: NTT pass variant -- 0 (default) / 6: radix-16 staged passes, 8 blocks/SM; 5: the same at 6 blocks/SM;
 * 7: radix-64 passes, 4 blocks/SM (round 1); 1 / 2: radix-64 at 5 / 6 blocks/SM; 3 / 4: radix-64 with a warp-local
 * exchange (4 / 5 blocks/SM).  Same results. */
int ws_set_ntt_variant(ws_ctx* ctx, int variant);

/* Elapsed device time (ms, CUDA events on the launch stream) of the kernels enqueued by the last
 * *_dev / host call on this ctx, split per pipeline stage; names[i] are static strings.
 * Returns the number of stages written (<= cap).  Synchronises the ctx stream.                  */
int ws_last_timings(ws_ctx* ctx, const char** names, float* ms, int cap);

#ifdef __cplusplus
}
#endif
#endif /* WASMSNARK_B200_H */
