/* wasmsnark_b200_addon.c -- thin Node.js N-API binding over libwasmsnark_b200.so.
 *
 * This is the binding a wasmsnark maintainer adds to make `groth16GenProof(witness, provingKey)`
 * (main_bn128.js:27-41; README name genZKSnarkProof) run on the GPU: no WASM, no workers; the JS side
 * (node/index.js) keeps the reference's surface and result formatting.  It cannot be built for Node in the build
 * image (no node_api.h, no node); it is written against the stable N-API (v6) and the C ABI in
 * include/wasmsnark_b200.h only, and is linked and executed there against a mock N-API runtime
 * (tests/host/napi_mock, tests/test_node_addon_mock.py).  Build where Node exists (or node-gyp: node/binding.gyp):
 *     cc -shared -fPIC -I$(node -p "require('node:path').dirname(process.execPath)+'/../include/node'") \
 *        -I../include wasmsnark_b200_addon.c -L../wasmsnark_b200 -lwasmsnark_b200 -o wasmsnark_b200.node
 *
 * Threading: prove() runs the blocking C call on a libuv worker thread and any number of proofs may be in flight on
 * one handle, as with the reference's promise API; that is safe because every C-ABI entry point holds its context's
 * mutex for the duration of the call (include/wasmsnark_b200.h, "Threading") -- overlapping calls on one handle queue
 * up in arrival order like the reference's actionQueue (src/bn128.js:282-300).
 *
 * Exports:
 *   create(device, ngpus) -> external handle                buildBn128()            src/bn128.js:173-265
 *        ngpus > 1 (or 0 = all): proofs and multiexps use every GPU of the box (ws_multi_*, the worker pool = the GPUs)
 *   destroy(handle)                                         Bn128.terminate()       src/bn128.js:562-566
 *        frees the GPU state at once (after the proofs in flight, if any); later calls on the handle throw
 *   prove(handle, witness, pkey, r|null, s|null, cb)        Bn128.groth16GenProof   src/bn128.js:580-720
 *        cb(err, {pi_a: ArrayBuffer(96), pi_b: ArrayBuffer(192), pi_c: ArrayBuffer(96)})   (async work)
 *   g1Multiexp(handle, scalars, points) -> ArrayBuffer(96)  Bn128.g1_multiexp       src/bn128.js:353-383
 *   g2Multiexp(handle, scalars, points) -> ArrayBuffer(192) Bn128.g2_multiexp       src/bn128.js:385-415
 *   g1Affine / g2Affine(handle, p) -> ArrayBuffer           Bn128.g1_affine/g2_affine src/bn128.js:417-427
 *   calcH(handle, signals, polsA, polsB, nSignals, domainSize) -> ArrayBuffer   Bn128.calcH  src/bn128.js:569-578
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>

#include "wasmsnark_b200.h"

#define CHECK(call)                                         \
    do {                                                    \
        if ((call) != napi_ok) {                            \
            napi_throw_error(env, NULL, "N-API call failed"); \
            return NULL;                                    \
        }                                                   \
    } while (0)

static int get_buffer(napi_env env, napi_value v, uint8_t** data, size_t* len) {
    bool is_ab = false, is_ta = false, is_buf = false;
    napi_is_arraybuffer(env, v, &is_ab);
    if (is_ab) return napi_get_arraybuffer_info(env, v, (void**)data, len) == napi_ok;
    napi_is_buffer(env, v, &is_buf);
    if (is_buf) return napi_get_buffer_info(env, v, (void**)data, len) == napi_ok;
    napi_is_typedarray(env, v, &is_ta);
    if (is_ta) {
        napi_typedarray_type t;
        size_t n, off;
        napi_value ab;
        if (napi_get_typedarray_info(env, v, &t, &n, (void**)data, &ab, &off) != napi_ok) return 0;
        size_t esz = (t == napi_uint32_array || t == napi_int32_array || t == napi_float32_array) ? 4 : (t == napi_uint16_array || t == napi_int16_array) ? 2 : (t >= napi_float64_array) ? 8 : 1;
        *len = n * esz;
        return 1;
    }
    return 0;
}

static napi_value make_ab(napi_env env, const uint8_t* src, size_t len) {
    void* dst;
    napi_value ab;
    if (napi_create_arraybuffer(env, len, &dst, &ab) != napi_ok) return NULL;
    memcpy(dst, src, len);
    return ab;
}

static napi_value throw_ws(napi_env env, ws_ctx* c, int rc) {
    (void)rc;
    napi_throw_error(env, "WASMSNARK_B200", c ? ws_last_error(c) : "wasmsnark_b200 error");
    return NULL;
}

/* one handle = one single-GPU context (always) + optionally the multi-GPU handle the heavy calls go through.
 * `inflight` / `closing` are touched on the JS thread only (Prove, prove_complete, Destroy all run there). */
typedef struct {
    ws_ctx* c;
    ws_multi* m;
    int inflight; /* proofs queued or running on worker threads */
    int closing;  /* destroy() was called while proofs were in flight: the last one to complete releases the GPU state */
} handle_t;

static void release_gpu_state(handle_t* h) {
    if (h->m) ws_multi_destroy(h->m);
    if (h->c) ws_destroy(h->c);
    h->m = NULL;
    h->c = NULL;
}

static void finalize_ctx(napi_env env, void* data, void* hint) {
    (void)env;
    (void)hint;
    handle_t* h = (handle_t*)data;
    release_gpu_state(h);
    free(h);
}

static napi_value Create(napi_env env, napi_callback_info info) {
    size_t argc = 2;
    napi_value argv[2];
    CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    int32_t device = 0, ngpus = 1;
    if (argc >= 1) napi_get_value_int32(env, argv[0], &device);
    if (argc >= 2) napi_get_value_int32(env, argv[1], &ngpus);
    handle_t* h = (handle_t*)calloc(1, sizeof *h);
    if (!h || ws_create(device, &h->c) || (ngpus != 1 && ws_multi_create(ngpus, &h->m))) {
        if (h && h->c) ws_destroy(h->c);
        free(h);
        napi_throw_error(env, "WASMSNARK_B200", "no usable CUDA device (ws_create / ws_multi_create failed); there is no CPU fallback");
        return NULL;
    }
    napi_value ext;
    CHECK(napi_create_external(env, h, finalize_ctx, NULL, &ext));
    return ext;
}

static handle_t* handle_arg(napi_env env, napi_value v) {
    void* p = NULL;
    return napi_get_value_external(env, v, &p) == napi_ok ? (handle_t*)p : NULL;
}
static ws_ctx* ctx_arg(napi_env env, napi_value v) {
    handle_t* h = handle_arg(env, v);
    return h && !h->closing ? h->c : NULL;
}

/* ---- prove: blocking C call wrapped in napi_async_work (the reference resolves a Promise from its workers) */
typedef struct {
    napi_async_work work;
    napi_ref cb, wit_ref, key_ref, handle_ref;
    handle_t* handle;
    ws_ctx* ctx;
    ws_multi* multi;
    uint8_t *wit, *key;
    size_t wit_len, key_len;
    uint8_t r[32], s[32];
    int have_r, have_s, rc;
    uint8_t pi_a[96], pi_b[192], pi_c[96];
    char err[256];
} prove_job;

static void prove_execute(napi_env env, void* data) {
    (void)env;
    prove_job* j = (prove_job*)data;
    const uint8_t *r = j->have_r ? j->r : NULL, *s = j->have_s ? j->s : NULL;
    if (j->multi) {
        j->rc = ws_multi_bn128_groth16_prove(j->multi, j->wit, j->wit_len, j->key, j->key_len, r, s, j->pi_a, j->pi_b, j->pi_c);
        if (j->rc) strncpy(j->err, ws_multi_last_error(j->multi), sizeof j->err - 1);
    } else {
        j->rc = ws_bn128_groth16_prove(j->ctx, j->wit, j->wit_len, j->key, j->key_len, r, s, j->pi_a, j->pi_b, j->pi_c);
        if (j->rc) strncpy(j->err, ws_last_error(j->ctx), sizeof j->err - 1);
    }
}

static void prove_complete(napi_env env, napi_status status, void* data) {
    prove_job* j = (prove_job*)data;
    napi_value cb, global, argv[2], result;
    napi_get_reference_value(env, j->cb, &cb);
    napi_get_global(env, &global);
    if (status != napi_ok || j->rc) {
        napi_value msg;
        napi_create_string_utf8(env, j->rc ? j->err : "async work cancelled", NAPI_AUTO_LENGTH, &msg);
        napi_create_error(env, NULL, msg, &argv[0]);
        napi_get_null(env, &argv[1]);
    } else {
        napi_value obj;
        napi_get_null(env, &argv[0]);
        napi_create_object(env, &obj);
        napi_set_named_property(env, obj, "pi_a", make_ab(env, j->pi_a, 96));
        napi_set_named_property(env, obj, "pi_b", make_ab(env, j->pi_b, 192));
        napi_set_named_property(env, obj, "pi_c", make_ab(env, j->pi_c, 96));
        argv[1] = obj;
    }
    if (--j->handle->inflight == 0 && j->handle->closing) release_gpu_state(j->handle);   /* terminate() during a proof */
    napi_call_function(env, global, cb, 2, argv, &result);
    napi_delete_reference(env, j->cb);
    napi_delete_reference(env, j->wit_ref);
    napi_delete_reference(env, j->key_ref);
    napi_delete_reference(env, j->handle_ref);
    napi_delete_async_work(env, j->work);
    free(j);
}

static napi_value Prove(napi_env env, napi_callback_info info) {
    size_t argc = 6;
    napi_value argv[6], name;
    CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    if (argc < 6) {
        napi_throw_type_error(env, NULL, "prove(handle, witness, pkey, r, s, cb)");
        return NULL;
    }
    prove_job* j = (prove_job*)calloc(1, sizeof *j);
    handle_t* hd = handle_arg(env, argv[0]);
    if (!j || !hd || !hd->c || hd->closing) {
        free(j);
        napi_throw_type_error(env, NULL, "prove: not a live handle (terminated, or not made by create())");
        return NULL;
    }
    j->handle = hd;
    j->ctx = hd->c;
    j->multi = hd->m;
    uint8_t* p;
    size_t l;
    if (!get_buffer(env, argv[1], &j->wit, &j->wit_len) || !get_buffer(env, argv[2], &j->key, &j->key_len)) {
        free(j);
        napi_throw_type_error(env, NULL, "witness and provingKey must be ArrayBuffers");
        return NULL;
    }
    if (get_buffer(env, argv[3], &p, &l) && l >= 32) { memcpy(j->r, p, 32); j->have_r = 1; }
    if (get_buffer(env, argv[4], &p, &l) && l >= 32) { memcpy(j->s, p, 32); j->have_s = 1; }
    /* keep the input buffers (and the handle) alive until the work completes */
    napi_create_reference(env, argv[0], 1, &j->handle_ref);
    napi_create_reference(env, argv[1], 1, &j->wit_ref);
    napi_create_reference(env, argv[2], 1, &j->key_ref);
    napi_create_reference(env, argv[5], 1, &j->cb);
    napi_create_string_utf8(env, "wasmsnark_b200.prove", NAPI_AUTO_LENGTH, &name);
    napi_status st = napi_create_async_work(env, NULL, name, prove_execute, prove_complete, j, &j->work);
    if (st == napi_ok && (st = napi_queue_async_work(env, j->work)) != napi_ok) napi_delete_async_work(env, j->work);
    if (st != napi_ok) {
        napi_delete_reference(env, j->cb);
        napi_delete_reference(env, j->wit_ref);
        napi_delete_reference(env, j->key_ref);
        napi_delete_reference(env, j->handle_ref);
        free(j);
        napi_throw_error(env, NULL, "could not queue the proof (napi_create_async_work / napi_queue_async_work)");
        return NULL;
    }
    hd->inflight++;
    return NULL;
}

/* ---- synchronous helpers (each a single blocking C call) */
static napi_value Multiexp(napi_env env, napi_callback_info info, int g2) {
    size_t argc = 3;
    napi_value argv[3];
    CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    handle_t* hd = handle_arg(env, argv[0]);
    ws_ctx* c = hd && !hd->closing ? hd->c : NULL;
    uint8_t *sc, *pt;
    size_t sl, pl;
    if (!c || !get_buffer(env, argv[1], &sc, &sl) || !get_buffer(env, argv[2], &pt, &pl) || sl % 32 || pl < (sl / 32) * (g2 ? 128 : 64)) {
        napi_throw_type_error(env, NULL, "multiexp(handle, scalars, points)");
        return NULL;
    }
    uint8_t out[192];
    int rc;
    if (hd->m) {
        rc = g2 ? ws_multi_bn128_g2_multiexp(hd->m, sc, pt, sl / 32, out) : ws_multi_bn128_g1_multiexp(hd->m, sc, pt, sl / 32, out);
        if (rc) {
            napi_throw_error(env, "WASMSNARK_B200", ws_multi_last_error(hd->m));
            return NULL;
        }
    } else {
        rc = g2 ? ws_bn128_g2_multiexp(c, sc, pt, sl / 32, out) : ws_bn128_g1_multiexp(c, sc, pt, sl / 32, out);
        if (rc) return throw_ws(env, c, rc);
    }
    return make_ab(env, out, g2 ? 192 : 96);
}
static napi_value G1Multiexp(napi_env env, napi_callback_info info) { return Multiexp(env, info, 0); }
static napi_value G2Multiexp(napi_env env, napi_callback_info info) { return Multiexp(env, info, 1); }

static napi_value Affine(napi_env env, napi_callback_info info, int g2) {
    size_t argc = 2;
    napi_value argv[2];
    CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    ws_ctx* c = ctx_arg(env, argv[0]);
    uint8_t* p;
    size_t l, rec = g2 ? 192 : 96;
    if (!c || !get_buffer(env, argv[1], &p, &l) || l < rec) {
        napi_throw_type_error(env, NULL, "affine(handle, point)");
        return NULL;
    }
    uint8_t out[192];
    int rc = g2 ? ws_bn128_g2_affine(c, p, out, 1, 0) : ws_bn128_g1_affine(c, p, out, 1, 0);
    if (rc) return throw_ws(env, c, rc);
    return make_ab(env, out, rec);
}
static napi_value G1Affine(napi_env env, napi_callback_info info) { return Affine(env, info, 0); }
static napi_value G2Affine(napi_env env, napi_callback_info info) { return Affine(env, info, 1); }

static napi_value CalcH(napi_env env, napi_callback_info info) {
    size_t argc = 6;
    napi_value argv[6];
    CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    ws_ctx* c = ctx_arg(env, argv[0]);
    uint8_t *sg, *pa, *pb;
    size_t sl, al, bl;
    uint32_t n_signals = 0, domain = 0;
    if (!c || !get_buffer(env, argv[1], &sg, &sl) || !get_buffer(env, argv[2], &pa, &al) || !get_buffer(env, argv[3], &pb, &bl) ||
        napi_get_value_uint32(env, argv[4], &n_signals) != napi_ok || napi_get_value_uint32(env, argv[5], &domain) != napi_ok || sl < (size_t)n_signals * 32) {
        napi_throw_type_error(env, NULL, "calcH(handle, signals, polsA, polsB, nSignals, domainSize)");
        return NULL;
    }
    void* dst;
    napi_value ab;
    CHECK(napi_create_arraybuffer(env, (size_t)domain * 32, &dst, &ab));
    int rc = ws_bn128_calc_h(c, sg, pa, al, pb, bl, n_signals, domain, (uint8_t*)dst);
    if (rc) return throw_ws(env, c, rc);
    return ab;
}

/* verify(vk, ic, inputs, proof): vk = alfa1|beta2|gamma2|delta2 (448 B), ic = (nPublic+1) * 64 B, inputs = nPublic * 32 B,
 * proof = pi_a|pi_b|pi_c (256 B), all plain little-endian (index.js packs the JSON decimals) -> boolean.
 * Replaces Bn128.groth16Verify, src/bn128.js:722-786; host code, ~10 ms, so it runs inline. */
static napi_value Verify(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4], res;
    CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    uint8_t *vk, *ic, *in, *pr;
    size_t vl, il, nl, pl;
    if (!get_buffer(env, argv[0], &vk, &vl) || !get_buffer(env, argv[1], &ic, &il) || !get_buffer(env, argv[2], &in, &nl) ||
        !get_buffer(env, argv[3], &pr, &pl) || vl != 448 || pl != 256 || il % 64 || nl % 32 || il / 64 != nl / 32 + 1) {
        napi_throw_type_error(env, NULL, "verify(vk448, ic, inputs, proof256)");
        return NULL;
    }
    int valid = 0;
    int rc = ws_bn128_groth16_verify(vk, vk + 64, vk + 192, vk + 320, ic, il / 64, nl ? in : NULL, nl / 32, pr, pr + 64, pr + 192, &valid);
    if (rc) {
        napi_throw_error(env, NULL, "ws_bn128_groth16_verify: bad arguments");
        return NULL;
    }
    CHECK(napi_get_boolean(env, valid == 1, &res));
    return res;
}

/* destroy(handle) = Bn128.terminate(): releases the contexts -- and with them the resident key, its window tables and all
 * workspaces, several GB of HBM -- NOW rather than at some later garbage collection; with proofs still in flight the last
 * one to complete does it.  The handle is dead afterwards (every call on it throws). */
static napi_value Destroy(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    handle_t* h = argc >= 1 ? handle_arg(env, argv[0]) : NULL;
    if (!h) {
        napi_throw_type_error(env, NULL, "destroy(handle)");
        return NULL;
    }
    if (h->inflight) h->closing = 1;
    else release_gpu_state(h);
    return NULL;
}

NAPI_MODULE_INIT() {
    napi_property_descriptor props[] = {
        {"create", NULL, Create, NULL, NULL, NULL, napi_default, NULL},
        {"destroy", NULL, Destroy, NULL, NULL, NULL, napi_default, NULL},
        {"prove", NULL, Prove, NULL, NULL, NULL, napi_default, NULL},
        {"g1Multiexp", NULL, G1Multiexp, NULL, NULL, NULL, napi_default, NULL},
        {"g2Multiexp", NULL, G2Multiexp, NULL, NULL, NULL, napi_default, NULL},
        {"g1Affine", NULL, G1Affine, NULL, NULL, NULL, napi_default, NULL},
        {"g2Affine", NULL, G2Affine, NULL, NULL, NULL, napi_default, NULL},
        {"calcH", NULL, CalcH, NULL, NULL, NULL, napi_default, NULL},
        {"verify", NULL, Verify, NULL, NULL, NULL, napi_default, NULL},
    };
    napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
    return exports;
}
