// multi.cu -- all the GPUs of one box behind ONE handle and ONE blocking call (ws_multi_*).
//
// Replaces the reference's worker pool as a whole: buildBn128() spawns W workers and Bn128.groth16GenProof /
// g1_multiexp / g2_multiexp split every multiexp over them by contiguous point ranges and add the partial points
// (/root/reference/src/bn128.js:173-265, :353-415, :580-720), with CALC_H as one task on one worker (:569-578).
// Here one host process drives one ws_ctx per GPU through the per-GPU C ABI (this file is a pure client of
// include/wasmsnark_b200.h plus CUDA peer copies):
//   * key sections are range-sharded over the GPUs once (ws_bn128_pkey_load_shard), cached per handle;
//   * per proof, ONE GPU (rotating from proof to proof) computes H and ships each GPU its slice h[h_lo, h_lo + h_cnt)
//     with cudaMemcpyPeerAsync over NVLink; every GPU runs its five partial MSMs;
//   * the 768-byte partial records are peer-copied to GPU 0, which adds them and runs the final assembly.
// Cross-device ordering uses CUDA events (cudaStreamWaitEvent works across devices of one process); no NCCL is needed
// inside one process.  The torchrun / NCCL form of the same split lives in wasmsnark_b200/parallel.py.
#include <cuda_runtime.h>
#include <sys/random.h>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wasmsnark_b200.h"
#include "host_io.cuh"
#include "tls_error.cuh"

#define WS_API __attribute__((visibility("default")))

namespace {
constexpr int MAXG = 16;

unsigned long long fingerprint(const uint8_t* p, size_t n) {      // same construction as the per-context key cache
    auto one = [](const uint8_t* q, size_t m) {
        unsigned long long h[4] = {0x243F6A8885A308D3ull ^ m, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
        const unsigned long long M = 0x9E3779B97F4A7C15ull;
        size_t i = 0;
        for (; i + 32 <= m; i += 32)
            for (int k = 0; k < 4; k++) {
                unsigned long long w;
                memcpy(&w, q + i + 8 * k, 8);
                h[k] = (h[k] ^ w) * M;
                h[k] = (h[k] << 29) | (h[k] >> 35);
            }
        for (; i < m; i++) h[i & 3] = (h[i & 3] ^ q[i]) * M;
        unsigned long long r = h[0];
        for (int k = 1; k < 4; k++) r = (r ^ (h[k] + (r << 6) + (r >> 2))) * M;
        return r ^ (r >> 31);
    };
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<size_t>(std::min<unsigned>(hw ? hw : 1, 16), n / ((size_t)8 << 20));
    if (nt <= 1) return one(p, n);
    std::vector<unsigned long long> part(nt);
    std::vector<std::thread> th;
    size_t per = ((n + nt - 1) / nt + 31) & ~(size_t)31;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
            part[t] = one(p + lo, hi - lo);
        });
    for (auto& x : th) x.join();
    return one((const uint8_t*)part.data(), part.size() * sizeof(part[0])) ^ n;
}
}  // namespace

struct ws_multi {
    int n = 0;
    int dev[MAXG] = {};
    ws_ctx* ctx[MAXG] = {};
    cudaStream_t st[MAXG] = {};            // one work stream per GPU (the library launches on it)
    cudaEvent_t ev_h[MAXG] = {};           // H ready on its owner (an event is recorded on a stream of its own device)
    cudaEvent_t ev_part[MAXG] = {};        // GPU g's partial record has landed on GPU 0
    ws::PinnedStager stager[MAXG];
    // proving-key shards of the last key seen (stateless call: keyed by length + content fingerprint)
    ws_pkey* key[MAXG] = {};
    unsigned long long key_fp = 0;
    size_t key_len = 0;
    uint32_t n_vars = 0, n_public = 0, domain = 0;
    uint32_t range[MAXG][6] = {};
    // per-GPU device buffers
    uint8_t* d_wit[MAXG] = {};             // nVars * 32 (only the needed ranges are filled on non-owners)
    uint8_t* d_small[MAXG] = {};           // r|s (64) | partial record (768) | proof (384) | all records (768 * n)
    uint8_t* d_h[MAXG] = {};               // domain * 32: full h on the owner, the slice (at its natural offset) elsewhere
    size_t wit_cap = 0, h_cap = 0;
    // multiexp scratch
    uint8_t* d_ms[MAXG] = {};
    size_t ms_cap[MAXG] = {};
    unsigned long long proofs = 0;         // rotates the H owner
    int balanced_from = 1 << 30;           // GPU count from which the balanced split replaces the uniform one: never by default
                                           // (ws_multi_set_split(m, 1) turns it on; a single blocking proof measured 9.5 ms uniform
                                           // vs 11.0 ms balanced on 8 GPUs -- the uniform split balances exactly)
    std::mutex mu;
    std::string err;
    int fail(int code, const std::string& m) {
        err = m;
        ws::tls_error_set(this, err);      // see ctx.cuh: the failing thread reads its own message back
        return code;
    }
    int fail_ctx(int g, int code) {
        err = std::string("gpu ") + std::to_string(dev[g]) + ": " + ws_last_error(ctx[g]);
        ws::tls_error_set(this, err);
        return code;
    }
};

namespace {
constexpr size_t OFF_RS = 0, OFF_PART = 256, OFF_PROOF = 1280, OFF_ALL = 2048;

void free_keys(ws_multi* m) {
    for (int g = 0; g < m->n; g++)
        if (m->key[g]) {
            ws_bn128_pkey_free(m->key[g]);
            m->key[g] = nullptr;
        }
    m->key_len = 0;
}

// Balanced split (the same construction as wasmsnark_b200/parallel.py: balanced_ranges): the five point sections laid end
// to end in the order B2, A, B1, C, H, each point weighted by its cost (G2 ~2.8 G1 additions, h is full-width, points at
// infinity -- x == 0 -- are free: their share is sampled on every 16th point), cut into n equal pieces; a GPU owns the
// points whose start falls into its piece -- one or two large contiguous pieces instead of 1/n of every section.
// ranges: [g][10] = {lo, cnt} in section order A, B1, B2, C, H.  Returns false for a malformed header.
bool balanced_ranges(const uint8_t* pkey, size_t plen, int n, uint32_t (*ranges)[10]) {
    if (plen < 488) return false;
    uint32_t h[10];
    memcpy(h, pkey, 40);
    const uint64_t nv = h[0], npub = h[1], dom = h[2];
    if (npub + 1 > nv) return false;
    // line order B2, A, B1, C, H -> index into the A, B1, B2, C, H layout
    const int sec[5] = {2, 0, 1, 3, 4};
    const uint64_t off[5] = {h[7], h[5], h[6], h[8], h[9]}, size[5] = {nv, nv, nv, nv - npub - 1, dom};
    const uint64_t psize[5] = {128, 64, 64, 64, 64}, xbytes[5] = {64, 32, 32, 32, 32}, base[5] = {280, 100, 100, 100, 120};
    uint64_t w[5], total = 0;
    for (int k = 0; k < 5; k++) {
        if (off[k] + size[k] * psize[k] > plen) return false;
        uint64_t samples = 0, finite = 0;
        for (uint64_t i = 0; i < size[k]; i += 16) {
            const uint8_t* x = pkey + off[k] + i * psize[k];
            bool nz = false;
            for (uint64_t b = 0; b < xbytes[k] && !nz; b++) nz = x[b] != 0;
            samples++;
            finite += nz;
        }
        w[k] = samples ? std::max<uint64_t>(1, base[k] * std::max<uint64_t>(finite, 1) / samples) : base[k];
        total += w[k] * size[k];
    }
    for (int g = 0; g < n; g++) {
        const uint64_t start = (uint64_t)g * total / n, end = (uint64_t)(g + 1) * total / n;
        uint64_t o = 0;
        for (int k = 0; k < 5; k++) {
            auto cut = [&](uint64_t pos) -> uint64_t {       // first element whose start is >= pos
                if (pos <= o) return 0;
                return std::min<uint64_t>(size[k], (pos - o + w[k] - 1) / w[k]);
            };
            const uint64_t lo = cut(start), hi = cut(end);
            ranges[g][2 * sec[k]] = hi > lo ? (uint32_t)lo : 0;
            ranges[g][2 * sec[k] + 1] = hi > lo ? (uint32_t)(hi - lo) : 0;
            o += w[k] * size[k];
        }
    }
    return true;
}

// Loads the key's shards (one host thread per GPU: the row regrouping of the polynomial sections is host work).
int load_keys(ws_multi* m, const uint8_t* pkey, size_t plen, unsigned long long fp) {
    free_keys(m);
    std::vector<int> rc(m->n, 0);
    std::vector<std::thread> th;
    uint32_t ranges[MAXG][10];
    const bool balanced = m->n >= m->balanced_from && balanced_ranges(pkey, plen, m->n, ranges);
    for (int g = 0; g < m->n; g++)
        th.emplace_back([&, g] {
            rc[g] = balanced ? ws_bn128_pkey_load_ranges(m->ctx[g], pkey, plen, ranges[g], &m->key[g]) : ws_bn128_pkey_load_shard(m->ctx[g], pkey, plen, g, m->n, &m->key[g]);
        });
    for (auto& x : th) x.join();
    for (int g = 0; g < m->n; g++)
        if (rc[g]) {
            int code = rc[g];
            m->fail_ctx(g, code);
            free_keys(m);
            return code;
        }
    ws_bn128_pkey_info(m->key[0], &m->n_vars, &m->n_public, &m->domain);
    for (int g = 0; g < m->n; g++) ws_bn128_pkey_shard_info(m->key[g], m->range[g]);
    size_t wb = (size_t)m->n_vars * 32, hb = (size_t)m->domain * 32;
    for (int g = 0; g < m->n; g++) {
        cudaSetDevice(m->dev[g]);
        if (wb > m->wit_cap || !m->d_wit[g]) {
            cudaFree(m->d_wit[g]);
            m->d_wit[g] = nullptr;
            if (cudaMalloc(&m->d_wit[g], wb + 256) != cudaSuccess) {
                m->d_wit[g] = nullptr;
                free_keys(m);          // no resident key without its buffers (prove_resident checks key[0] only)
                return m->fail(WS_E_CUDA, "witness buffer");
            }
        }
        if (hb > m->h_cap || !m->d_h[g]) {
            cudaFree(m->d_h[g]);
            m->d_h[g] = nullptr;
            if (cudaMalloc(&m->d_h[g], hb + 256) != cudaSuccess) {
                m->d_h[g] = nullptr;
                free_keys(m);
                return m->fail(WS_E_CUDA, "h buffer");
            }
        }
    }
    m->wit_cap = std::max(m->wit_cap, wb);
    m->h_cap = std::max(m->h_cap, hb);
    m->key_fp = fp;
    m->key_len = plen;
    return WS_OK;
}

// Enqueues one proof over all GPUs; the result lands in GPU 0's proof slot, ordered on st[0].
int enqueue_proof(ws_multi* m, const uint8_t* witness, const uint8_t* rs) {
    const int n = m->n;
    const int owner = (int)(m->proofs++ % (unsigned long long)n);
    // 1. inputs + the four witness MSMs everywhere (the owner needs the whole witness for the LC step); one host thread
    //    per GPU, so that the staging copies of the (pageable) witness run side by side
    std::vector<int> rc1(n, WS_OK);
    std::vector<std::string> err1(n);
    auto stage1 = [&](int g) {
        cudaSetDevice(m->dev[g]);
        const uint32_t* r = m->range[g];
        if (g == owner) {
            if (m->stager[g].upload(m->d_wit[g], witness, (size_t)m->n_vars * 32, m->st[g], m->dev[g]) != cudaSuccess) {
                rc1[g] = WS_E_CUDA;
                err1[g] = "witness upload";
                return;
            }
        } else {
            uint32_t lo = 0xFFFFFFFFu, hi = 0;
            if (r[1]) lo = std::min(lo, r[0]), hi = std::max(hi, r[0] + r[1]);
            if (r[3]) lo = std::min(lo, r[2]), hi = std::max(hi, r[2] + r[3]);
            if (hi > lo && m->stager[g].upload(m->d_wit[g] + (size_t)lo * 32, witness + (size_t)lo * 32, (size_t)(hi - lo) * 32, m->st[g], m->dev[g]) != cudaSuccess) {
                rc1[g] = WS_E_CUDA;
                err1[g] = "witness upload";
                return;
            }
        }
        int rc = WS_OK;
        if (g == 0) {
            if (cudaMemcpyAsync(m->d_small[0] + OFF_RS, rs, 64, cudaMemcpyHostToDevice, m->st[0]) != cudaSuccess) {
                rc1[g] = WS_E_CUDA;
                err1[g] = "r, s upload";
                return;
            }
            rc = ws_bn128_groth16_prepare_dev(m->ctx[0], m->key[0], m->d_small[0] + OFF_RS, m->st[0]);
        }
        if (!rc) rc = ws_bn128_groth16_witness_msms_dev(m->ctx[g], m->key[g], m->d_wit[g], m->d_small[g] + OFF_PART, m->st[g]);
        if (rc) {
            rc1[g] = rc;
            err1[g] = std::string("gpu ") + std::to_string(m->dev[g]) + ": " + ws_last_error(m->ctx[g]);
        }
    };
    {
        std::vector<std::thread> th;
        for (int g = 1; g < n; g++) th.emplace_back(stage1, g);
        stage1(0);
        for (auto& x : th) x.join();
    }
    for (int g = 0; g < n; g++)
        if (rc1[g]) return m->fail(rc1[g], err1[g]);
    // 2. H on its owner
    cudaSetDevice(m->dev[owner]);
    int rc = ws_bn128_groth16_h_dev(m->ctx[owner], m->key[owner], m->d_wit[owner], m->d_h[owner], m->st[owner]);
    if (rc) return m->fail_ctx(owner, rc);
    cudaEventRecord(m->ev_h[owner], m->st[owner]);
    // 3. every GPU: its slice of h (peer copy over NVLink), the H MSM, and the record to GPU 0
    for (int g = 0; g < n; g++) {
        cudaSetDevice(m->dev[g]);
        const uint32_t* r = m->range[g];
        uint8_t* hs = m->d_h[g] + (size_t)r[4] * 32;
        if (g != owner && r[5]) {
            cudaStreamWaitEvent(m->st[g], m->ev_h[owner], 0);
            if (r[5] && cudaMemcpyPeerAsync(hs, m->dev[g], m->d_h[owner] + (size_t)r[4] * 32, m->dev[owner], (size_t)r[5] * 32, m->st[g]) != cudaSuccess)
                return m->fail(WS_E_CUDA, "peer copy of h");
        }
        if ((rc = ws_bn128_groth16_h_msm_dev(m->ctx[g], m->key[g], hs, m->d_small[g] + OFF_PART, m->st[g]))) return m->fail_ctx(g, rc);
        if (cudaMemcpyPeerAsync(m->d_small[0] + OFF_ALL + (size_t)g * WS_PARTIALS_BYTES, m->dev[0], m->d_small[g] + OFF_PART, m->dev[g], WS_PARTIALS_BYTES, m->st[g]) != cudaSuccess)
            return m->fail(WS_E_CUDA, "peer copy of the partial record");
        cudaEventRecord(m->ev_part[g], m->st[g]);
    }
    // 4. GPU 0 adds the records and assembles (src/bn128.js:642-712)
    cudaSetDevice(m->dev[0]);
    for (int g = 1; g < n; g++) cudaStreamWaitEvent(m->st[0], m->ev_part[g], 0);
    rc = ws_bn128_groth16_finish_dev(m->ctx[0], m->key[0], m->d_small[0] + OFF_ALL, n, m->d_small[0] + OFF_RS, m->d_small[0] + OFF_PROOF, m->st[0]);
    if (rc) return m->fail_ctx(0, rc);
    // the next proof's owner must not overwrite d_h / d_wit before this proof's consumers are done: the callers
    // (collect_proof, or drain after a failed enqueue) synchronise every GPU's stream before the handle is used again
    return WS_OK;
}

// waits for everything enqueued on the handle's streams and clears a pending CUDA error
void drain(ws_multi* m) {
    for (int g = 0; g < m->n; g++) {
        cudaSetDevice(m->dev[g]);
        cudaStreamSynchronize(m->st[g]);
    }
    cudaGetLastError();
}

int collect_proof(ws_multi* m, uint8_t* pi_a, uint8_t* pi_b, uint8_t* pi_c) {
    cudaSetDevice(m->dev[0]);
    uint8_t outb[384];
    if (cudaMemcpyAsync(outb, m->d_small[0] + OFF_PROOF, 384, cudaMemcpyDeviceToHost, m->st[0]) != cudaSuccess) return m->fail(WS_E_CUDA, "proof download");
    bool ok = true;
    for (int g = 0; g < m->n; g++) {
        cudaSetDevice(m->dev[g]);
        ok &= cudaStreamSynchronize(m->st[g]) == cudaSuccess;
    }
    if (!ok || cudaGetLastError() != cudaSuccess) return m->fail(WS_E_CUDA, std::string("prove: ") + cudaGetErrorString(cudaGetLastError()));
    memcpy(pi_a, outb, 96);
    memcpy(pi_b, outb + 96, 192);
    memcpy(pi_c, outb + 288, 96);
    return WS_OK;
}

template <int G>
int multiexp(ws_multi* m, const uint8_t* scalars, const uint8_t* points, size_t npts, uint8_t* out) {
    constexpr size_t PS = G == 1 ? 64 : 128, OS = G == 1 ? 96 : 192;
    if (!m || !out || (npts && (!scalars || !points))) return m ? m->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    const int n = m->n;
    const size_t per = npts / n;
    for (int g = 0; g < n; g++) {
        cudaSetDevice(m->dev[g]);
        const size_t lo = per * g, cnt = g < n - 1 ? per : npts - per * (n - 1);     // src/bn128.js:355-361
        const size_t need = ((cnt * 32 + 255) & ~(size_t)255) + ((cnt * PS + 255) & ~(size_t)255) + 512;
        if (need > m->ms_cap[g]) {
            cudaStreamSynchronize(m->st[g]);
            cudaFree(m->d_ms[g]);
            m->d_ms[g] = nullptr;
            m->ms_cap[g] = 0;
            if (cudaMalloc(&m->d_ms[g], need + need / 8) != cudaSuccess) return m->fail(WS_E_CUDA, "multiexp input buffer");
            m->ms_cap[g] = need + need / 8;
        }
        uint8_t* ds = m->d_ms[g];
        uint8_t* dp = ds + ((cnt * 32 + 255) & ~(size_t)255);
        uint8_t* dout = dp + ((cnt * PS + 255) & ~(size_t)255);
        if (m->stager[g].upload(ds, scalars + lo * 32, cnt * 32, m->st[g], m->dev[g]) != cudaSuccess ||
            m->stager[g].upload(dp, points + lo * PS, cnt * PS, m->st[g], m->dev[g]) != cudaSuccess)
            return m->fail(WS_E_CUDA, "multiexp upload");
        int rc = G == 1 ? ws_bn128_g1_multiexp_dev(m->ctx[g], ds, dp, cnt, dout, 0, m->st[g]) : ws_bn128_g2_multiexp_dev(m->ctx[g], ds, dp, cnt, dout, 0, m->st[g]);
        if (rc) return m->fail_ctx(g, rc);
        if (cudaMemcpyPeerAsync(m->d_small[0] + OFF_ALL + (size_t)g * OS, m->dev[0], dout, m->dev[g], OS, m->st[g]) != cudaSuccess) return m->fail(WS_E_CUDA, "peer copy of the partial sum");
        cudaEventRecord(m->ev_part[g], m->st[g]);
    }
    cudaSetDevice(m->dev[0]);
    for (int g = 1; g < n; g++) cudaStreamWaitEvent(m->st[0], m->ev_part[g], 0);
    uint8_t* dsum = m->d_small[0] + OFF_PROOF;
    int rc = G == 1 ? ws_bn128_g1_sum_dev(m->ctx[0], m->d_small[0] + OFF_ALL, n, dsum, m->st[0]) : ws_bn128_g2_sum_dev(m->ctx[0], m->d_small[0] + OFF_ALL, n, dsum, m->st[0]);
    if (rc) return m->fail_ctx(0, rc);
    uint8_t outb[192];
    cudaMemcpyAsync(outb, dsum, OS, cudaMemcpyDeviceToHost, m->st[0]);
    bool ok = true;
    for (int g = 0; g < n; g++) {
        cudaSetDevice(m->dev[g]);
        ok &= cudaStreamSynchronize(m->st[g]) == cudaSuccess;
    }
    if (!ok || cudaGetLastError() != cudaSuccess) return m->fail(WS_E_CUDA, "multiexp");
    memcpy(out, outb, OS);
    return WS_OK;
}
}  // namespace

extern "C" {

WS_API int ws_multi_create(int ngpus, ws_multi** out) {
    if (!out) return WS_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return WS_E_CUDA;
    if (ngpus <= 0) ngpus = ndev;
    if (ngpus > ndev || ngpus > MAXG) return WS_E_ARG;
    ws_multi* m = new ws_multi();
    m->n = ngpus;
    bool ok = true;
    for (int g = 0; g < ngpus && ok; g++) {
        m->dev[g] = g;
        ok = ws_create(g, &m->ctx[g]) == WS_OK && cudaSetDevice(g) == cudaSuccess && cudaStreamCreateWithFlags(&m->st[g], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&m->ev_part[g], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&m->ev_h[g], cudaEventDisableTiming) == cudaSuccess &&
             cudaMalloc(&m->d_small[g], OFF_ALL + (size_t)WS_PARTIALS_BYTES * MAXG + 256) == cudaSuccess;
        for (int p = 0; p < ngpus && ok; p++) {                 // direct NVLink copies between every pair
            if (p == g) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g, p);
            if (can && cudaDeviceEnablePeerAccess(p, 0) != cudaSuccess) cudaGetLastError();    // already enabled is fine
        }
    }
    if (!ok) {
        ws_multi_destroy(m);
        return WS_E_CUDA;
    }
    *out = m;
    return WS_OK;
}

WS_API void ws_multi_destroy(ws_multi* m) {
    if (!m) return;
    for (int g = 0; g < m->n; g++) {
        if (!m->ctx[g]) continue;
        cudaSetDevice(m->dev[g]);
        if (m->st[g]) cudaStreamSynchronize(m->st[g]);
    }
    free_keys(m);
    for (int g = 0; g < m->n; g++) {
        cudaSetDevice(m->dev[g]);
        cudaFree(m->d_wit[g]);
        cudaFree(m->d_small[g]);
        cudaFree(m->d_h[g]);
        cudaFree(m->d_ms[g]);
        m->stager[g].destroy();
        if (m->ev_part[g]) cudaEventDestroy(m->ev_part[g]);
        if (m->ev_h[g]) cudaEventDestroy(m->ev_h[g]);
        if (m->st[g]) cudaStreamDestroy(m->st[g]);
        if (m->ctx[g]) ws_destroy(m->ctx[g]);
    }
    ws::tls_error_forget(m);
    delete m;
}

// The balanced partition itself (pure host code, no device): out = n x 10 values.  Exposed so that other hosts of the
// per-GPU C ABI (one process per GPU) can use exactly the partition ws_multi_* uses, and for the CPU tests.
WS_API int ws_balanced_ranges(const uint8_t* pkey, size_t plen, int n, uint32_t* out) {
    if (!pkey || !out || n < 1 || n > MAXG) return WS_E_ARG;
    uint32_t ranges[MAXG][10];
    if (!balanced_ranges(pkey, plen, n, ranges)) return WS_E_ARG;
    memcpy(out, ranges, (size_t)n * 10 * sizeof(uint32_t));
    return WS_OK;
}

WS_API int ws_multi_gpus(const ws_multi* m) { return m ? m->n : 0; }
// mode 0: uniform split (1/n of every section, the reference's worker split); 1: balanced split; applies to keys loaded afterwards
WS_API int ws_multi_set_split(ws_multi* m, int mode) {
    if (!m) return WS_E_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    m->balanced_from = mode ? 2 : 1 << 30;
    return WS_OK;
}
WS_API const char* ws_multi_last_error(const ws_multi* m) {
    if (!m) return "null handle";
    const ws::TlsError& e = ws::tls_error();
    return e.owner == m ? e.text.c_str() : m->err.c_str();
}

WS_API int ws_multi_bn128_g1_multiexp(ws_multi* m, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[96]) { return multiexp<1>(m, scalars, points, n, out); }
WS_API int ws_multi_bn128_g2_multiexp(ws_multi* m, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[192]) { return multiexp<2>(m, scalars, points, n, out); }

// Resident form: the key's shards are loaded once (replacing whatever key the handle held) and proofs then take only the
// witness -- no per-call fingerprint of the key bytes, which at 8 GPUs costs more host time than the proof takes.
WS_API int ws_multi_bn128_pkey_load(ws_multi* m, const uint8_t* pkey, size_t plen) {
    if (!m || !pkey) return m ? m->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    for (int g = 0; g < m->n; g++) {
        cudaSetDevice(m->dev[g]);
        cudaStreamSynchronize(m->st[g]);
    }
    return load_keys(m, pkey, plen, fingerprint(pkey, plen));
}

WS_API int ws_multi_bn128_groth16_prove_resident(ws_multi* m, const uint8_t* witness, size_t wlen, const uint8_t* r, const uint8_t* s, uint8_t pi_a[96],
                                                 uint8_t pi_b[192], uint8_t pi_c[96]) {
    if (!m || !witness || !pi_a || !pi_b || !pi_c) return m ? m->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    if (!m->key[0]) return m->fail(WS_E_ARG, "no resident proving key: call ws_multi_bn128_pkey_load first");
    if (wlen < (size_t)m->n_vars * 32) return m->fail(WS_E_ARG, "witness shorter than nVars * 32 bytes");
    uint8_t rs[64];
    if (r) memcpy(rs, r, 32);
    else if (getrandom(rs, 32, 0) != 32) return m->fail(WS_E_ARG, "getrandom failed");
    if (s) memcpy(rs + 32, s, 32);
    else if (getrandom(rs + 32, 32, 0) != 32) return m->fail(WS_E_ARG, "getrandom failed");
    int rc = enqueue_proof(m, witness, rs);
    if (rc) {
        drain(m);             // whatever part of the proof was enqueued must not overlap the next call's buffers
        return rc;
    }
    return collect_proof(m, pi_a, pi_b, pi_c);
}

WS_API int ws_multi_bn128_groth16_prove(ws_multi* m, const uint8_t* witness, size_t wlen, const uint8_t* pkey, size_t plen, const uint8_t* r, const uint8_t* s,
                                        uint8_t pi_a[96], uint8_t pi_b[192], uint8_t pi_c[96]) {
    if (!m || !witness || !pkey || !pi_a || !pi_b || !pi_c) return m ? m->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    uint8_t rs[64];
    if (r) memcpy(rs, r, 32);
    else if (getrandom(rs, 32, 0) != 32) return m->fail(WS_E_ARG, "getrandom failed");
    if (s) memcpy(rs + 32, s, 32);
    else if (getrandom(rs + 32, 32, 0) != 32) return m->fail(WS_E_ARG, "getrandom failed");
    // same-length key cached: start the proof with it and fingerprint the caller's buffer meanwhile (see prover.cu)
    bool optimistic = m->key[0] && m->key_len == plen && wlen >= (size_t)m->n_vars * 32;
    if (optimistic && enqueue_proof(m, witness, rs) != WS_OK) {
        drain(m);
        optimistic = false;
    }
    unsigned long long fp = fingerprint(pkey, plen);
    if (optimistic && fp == m->key_fp) return collect_proof(m, pi_a, pi_b, pi_c);
    drain(m);
    int rc = load_keys(m, pkey, plen, fp);
    if (rc) return rc;
    if (wlen < (size_t)m->n_vars * 32) return m->fail(WS_E_ARG, "witness shorter than nVars * 32 bytes");
    if ((rc = enqueue_proof(m, witness, rs))) {
        drain(m);
        return rc;
    }
    return collect_proof(m, pi_a, pi_b, pi_c);
}

}  // extern "C"
