/* napi_mock.c -- a toy N-API runtime, TEST SCAFFOLDING ONLY (tests/test_node_addon_mock.py).
 *
 * The build image has no Node.js, so node/wasmsnark_b200_addon.c could only be type-checked.  This file implements the
 * N-API functions the addon calls (tests/host/napi_mock/node_api.h) over a minimal value model -- undefined / null /
 * boolean / number / string / object / function / external / ArrayBuffer / typed array / Buffer / Error -- so that the
 * addon can be linked into a shared object and its exports called from a test driver (Python, ctypes) the way
 * node/index.js calls them: argument validation, exceptions, the external handle's life cycle, and the asynchronous
 * prove() whose blocking C call runs on a worker thread (napi_async_work: execute on a pthread, complete on the
 * caller's thread, several works side by side).  It mimics documented N-API behaviour only as far as the addon relies on
 * it; it is not a JavaScript engine and proves nothing about V8.
 *
 * Values live until mock_env_free (no garbage collection); finalizers of externals run there, as at environment
 * teardown in Node.  One environment is used from one thread, apart from the execute callbacks.
 */
#include <node_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

enum { K_UNDEFINED, K_NULL, K_BOOL, K_NUMBER, K_STRING, K_OBJECT, K_FUNCTION, K_EXTERNAL, K_ARRAYBUFFER, K_TYPEDARRAY, K_BUFFER, K_ERROR };

typedef napi_value (*mock_host_fn)(napi_env env, size_t argc, napi_value* argv, void* data);

typedef struct prop {
    char* name;
    napi_value v;
    struct prop* next;
} prop;

struct napi_value__ {
    int kind;
    double num;
    char* str;             /* string contents / error message */
    char* code;            /* error code */
    int type_error;
    uint8_t* data;         /* ArrayBuffer / Buffer storage (owned), typed array view (not owned), external pointer */
    size_t len;            /* bytes (ArrayBuffer, Buffer) or elements (typed array) */
    int ta_type;
    size_t ta_off;
    napi_value ab;
    prop* props;
    napi_callback cb;      /* function made by the addon */
    void* cb_data;
    mock_host_fn host_fn;  /* function made by the test driver */
    void* host_data;
    napi_finalize fin;
    void* fin_hint;
    napi_value next_all;
};

struct napi_ref__ {
    napi_value v;
    uint32_t count;
};

struct napi_callback_info__ {
    size_t argc;
    napi_value* argv;
    napi_value this_arg;
    void* data;
};

struct napi_async_work__ {
    napi_async_execute_callback execute;
    napi_async_complete_callback complete;
    void* data;
    napi_env env;
    int queued, deleted;
    struct napi_async_work__* next;
};

struct napi_env__ {
    napi_value all;           /* every value ever made */
    napi_value undefined, null, global;
    napi_value pending;       /* pending exception */
    struct napi_async_work__* queue;
    int live_refs;
    int live_works;
};

static napi_value new_value(napi_env env, int kind) {
    napi_value v = (napi_value)calloc(1, sizeof *v);
    v->kind = kind;
    v->next_all = env->all;
    env->all = v;
    return v;
}

static char* dup_n(const char* s, size_t n) {
    if (n == NAPI_AUTO_LENGTH) n = strlen(s);
    char* d = (char*)malloc(n + 1);
    memcpy(d, s, n);
    d[n] = 0;
    return d;
}

static napi_value make_error(napi_env env, const char* code, const char* msg, int type_error) {
    napi_value e = new_value(env, K_ERROR);
    e->str = dup_n(msg ? msg : "", NAPI_AUTO_LENGTH);
    e->code = code ? dup_n(code, NAPI_AUTO_LENGTH) : NULL;
    e->type_error = type_error;
    return e;
}

/* ---------------------------------------------------------------- the N-API subset ---------------------------------- */
#define NEED(x) do { if (!(x)) return napi_invalid_arg; } while (0)

API napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
    NEED(env && info);
    if (argc) {
        if (argv)
            for (size_t i = 0; i < *argc; i++) argv[i] = i < info->argc ? info->argv[i] : env->undefined;   /* missing arguments read as undefined */
        *argc = info->argc;
    }
    if (this_arg) *this_arg = info->this_arg;
    if (data) *data = info->data;
    return napi_ok;
}

API napi_status napi_throw_error(napi_env env, const char* code, const char* msg) {
    NEED(env);
    if (env->pending) return napi_pending_exception;
    env->pending = make_error(env, code, msg, 0);
    return napi_ok;
}

API napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg) {
    NEED(env);
    if (env->pending) return napi_pending_exception;
    env->pending = make_error(env, code, msg, 1);
    return napi_ok;
}

static prop* find_prop(napi_value o, const char* name) {
    for (prop* p = o->props; p; p = p->next)
        if (!strcmp(p->name, name)) return p;
    return NULL;
}

API napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value) {
    NEED(env && object && utf8name && value);
    if (object->kind != K_OBJECT && object->kind != K_FUNCTION && object->kind != K_ERROR) return napi_object_expected;
    prop* p = find_prop(object, utf8name);
    if (!p) {
        p = (prop*)calloc(1, sizeof *p);
        p->name = dup_n(utf8name, NAPI_AUTO_LENGTH);
        p->next = object->props;
        object->props = p;
    }
    p->v = value;
    return napi_ok;
}

API napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref* result) {
    NEED(env && value && result);
    napi_ref r = (napi_ref)calloc(1, sizeof *r);
    r->v = value;
    r->count = initial_refcount;
    env->live_refs++;
    *result = r;
    return napi_ok;
}

API napi_status napi_delete_reference(napi_env env, napi_ref ref) {
    NEED(env && ref);
    env->live_refs--;
    free(ref);
    return napi_ok;
}

API napi_status napi_get_reference_value(napi_env env, napi_ref ref, napi_value* result) {
    NEED(env && ref && result);
    *result = ref->v;
    return napi_ok;
}

API napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result) {
    NEED(env && value && result);
    if (value->kind != K_NUMBER) return napi_number_expected;
    *result = (uint32_t)(int64_t)value->num;
    return napi_ok;
}

API napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result) {
    NEED(env && value && result);
    if (value->kind != K_NUMBER) return napi_number_expected;
    *result = (int32_t)(int64_t)value->num;
    return napi_ok;
}

API napi_status napi_get_null(napi_env env, napi_value* result) {
    NEED(env && result);
    *result = env->null;
    return napi_ok;
}

API napi_status napi_get_global(napi_env env, napi_value* result) {
    NEED(env && result);
    *result = env->global;
    return napi_ok;
}

API napi_status napi_get_boolean(napi_env env, bool value, napi_value* result) {
    NEED(env && result);
    napi_value v = new_value(env, K_BOOL);
    v->num = value ? 1 : 0;
    *result = v;
    return napi_ok;
}

API napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result) {
    NEED(env && str && result);
    napi_value v = new_value(env, K_STRING);
    v->str = dup_n(str, length);
    *result = v;
    return napi_ok;
}

API napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result) {
    NEED(env && result);
    napi_value v = new_value(env, K_ARRAYBUFFER);
    v->data = (uint8_t*)calloc(byte_length ? byte_length : 1, 1);      /* ArrayBuffers are zero-filled */
    if (!v->data) return napi_generic_failure;
    v->len = byte_length;
    if (data) *data = v->data;
    *result = v;
    return napi_ok;
}

API napi_status napi_create_object(napi_env env, napi_value* result) {
    NEED(env && result);
    *result = new_value(env, K_OBJECT);
    return napi_ok;
}

API napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value* result) {
    NEED(env && msg && result);
    if (msg->kind != K_STRING) return napi_string_expected;
    if (code && code->kind != K_STRING) return napi_string_expected;
    *result = make_error(env, code ? code->str : NULL, msg->str, 0);
    return napi_ok;
}

API napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result) {
    NEED(env && result);
    napi_value v = new_value(env, K_EXTERNAL);
    v->data = (uint8_t*)data;
    v->fin = finalize_cb;
    v->fin_hint = finalize_hint;
    *result = v;
    return napi_ok;
}

API napi_status napi_get_value_external(napi_env env, napi_value value, void** result) {
    NEED(env && value && result);
    if (value->kind != K_EXTERNAL) return napi_invalid_arg;
    *result = value->data;
    return napi_ok;
}

API napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name, napi_async_execute_callback execute,
                                       napi_async_complete_callback complete, void* data, napi_async_work* result) {
    (void)async_resource;
    NEED(env && execute && result && async_resource_name);
    napi_async_work w = (napi_async_work)calloc(1, sizeof *w);
    w->execute = execute;
    w->complete = complete;
    w->data = data;
    w->env = env;
    env->live_works++;
    *result = w;
    return napi_ok;
}

API napi_status napi_queue_async_work(napi_env env, napi_async_work work) {
    NEED(env && work);
    if (work->queued) return napi_generic_failure;
    work->queued = 1;
    work->next = NULL;
    napi_async_work* tail = &env->queue;
    while (*tail) tail = &(*tail)->next;
    *tail = work;
    return napi_ok;
}

API napi_status napi_delete_async_work(napi_env env, napi_async_work work) {
    NEED(env && work);
    work->deleted = 1;      /* freed by mock_run_async / mock_env_free: complete() deletes the work it runs inside of */
    env->live_works--;
    return napi_ok;
}

API napi_status napi_is_typedarray(napi_env env, napi_value value, bool* result) {
    NEED(env && value && result);
    *result = value->kind == K_TYPEDARRAY || value->kind == K_BUFFER;       /* a Node Buffer is a Uint8Array */
    return napi_ok;
}

API napi_status napi_is_buffer(napi_env env, napi_value value, bool* result) {
    NEED(env && value && result);
    *result = value->kind == K_BUFFER;
    return napi_ok;
}

API napi_status napi_is_arraybuffer(napi_env env, napi_value value, bool* result) {
    NEED(env && value && result);
    *result = value->kind == K_ARRAYBUFFER;
    return napi_ok;
}

API napi_status napi_get_typedarray_info(napi_env env, napi_value ta, napi_typedarray_type* type, size_t* length, void** data, napi_value* arraybuffer,
                                         size_t* byte_offset) {
    NEED(env && ta);
    if (ta->kind != K_TYPEDARRAY && ta->kind != K_BUFFER) return napi_invalid_arg;
    if (type) *type = ta->kind == K_BUFFER ? napi_uint8_array : (napi_typedarray_type)ta->ta_type;
    if (length) *length = ta->len;
    if (data) *data = ta->data;
    if (arraybuffer) *arraybuffer = ta->ab;
    if (byte_offset) *byte_offset = ta->ta_off;
    return napi_ok;
}

API napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length) {
    NEED(env && value);
    if (value->kind != K_BUFFER) return napi_invalid_arg;
    if (data) *data = value->data;
    if (length) *length = value->len;
    return napi_ok;
}

API napi_status napi_get_arraybuffer_info(napi_env env, napi_value ab, void** data, size_t* byte_length) {
    NEED(env && ab);
    if (ab->kind != K_ARRAYBUFFER) return napi_invalid_arg;
    if (data) *data = ab->data;
    if (byte_length) *byte_length = ab->len;
    return napi_ok;
}

static napi_value invoke(napi_env env, napi_value recv, napi_value fn, size_t argc, napi_value* argv) {
    if (fn->host_fn) return fn->host_fn(env, argc, argv, fn->host_data);
    struct napi_callback_info__ info = {argc, argv, recv, fn->cb_data};
    return fn->cb(env, &info);
}

API napi_status napi_call_function(napi_env env, napi_value recv, napi_value func, size_t argc, const napi_value* argv, napi_value* result) {
    NEED(env && func);
    if (env->pending) return napi_pending_exception;
    if (func->kind != K_FUNCTION) return napi_function_expected;
    napi_value r = invoke(env, recv, func, argc, (napi_value*)argv);
    if (env->pending) return napi_pending_exception;
    if (result) *result = r ? r : env->undefined;
    return napi_ok;
}

API napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor* properties) {
    NEED(env && object && (properties || !property_count));
    for (size_t i = 0; i < property_count; i++) {
        const napi_property_descriptor* d = &properties[i];
        if (!d->utf8name) return napi_name_expected;
        napi_value v = d->value;
        if (d->method) {
            v = new_value(env, K_FUNCTION);
            v->cb = d->method;
            v->cb_data = d->data;
        }
        if (!v) return napi_invalid_arg;
        napi_status st = napi_set_named_property(env, object, d->utf8name, v);
        if (st != napi_ok) return st;
    }
    return napi_ok;
}

/* ---------------------------------------------------------------- the test driver's side ---------------------------- */
API napi_env mock_env_new(void) {
    napi_env env = (napi_env)calloc(1, sizeof *env);
    env->undefined = new_value(env, K_UNDEFINED);
    env->null = new_value(env, K_NULL);
    env->global = new_value(env, K_OBJECT);
    return env;
}

/* environment teardown: finalizers of externals run (GPU contexts are released), everything is freed.  Returns the number
 * of references + async works the addon still holds (0 = it cleaned up after itself). */
API int mock_env_free(napi_env env) {
    int leaked = env->live_refs + env->live_works;
    for (napi_value v = env->all; v; v = v->next_all)
        if (v->kind == K_EXTERNAL && v->fin) v->fin(env, v->data, v->fin_hint);
    napi_value v = env->all;
    while (v) {
        napi_value nx = v->next_all;
        if (v->kind == K_ARRAYBUFFER || v->kind == K_BUFFER) free(v->data);
        free(v->str);
        free(v->code);
        for (prop* p = v->props; p;) {
            prop* pn = p->next;
            free(p->name);
            free(p);
            p = pn;
        }
        free(v);
        v = nx;
    }
    for (napi_async_work w = env->queue; w;) {
        napi_async_work wn = w->next;
        free(w);
        w = wn;
    }
    free(env);
    return leaked;
}

API napi_value mock_undefined(napi_env env) { return env->undefined; }
API napi_value mock_null(napi_env env) { return env->null; }
API napi_value mock_object(napi_env env) { return new_value(env, K_OBJECT); }
API napi_value mock_number(napi_env env, double x) {
    napi_value v = new_value(env, K_NUMBER);
    v->num = x;
    return v;
}
API napi_value mock_string(napi_env env, const char* s) {
    napi_value v = new_value(env, K_STRING);
    v->str = dup_n(s, NAPI_AUTO_LENGTH);
    return v;
}
API napi_value mock_arraybuffer(napi_env env, const void* src, size_t len) {      /* a copy, like new Uint8Array(bytes).buffer */
    void* d;
    napi_value v;
    napi_create_arraybuffer(env, len, &d, &v);
    if (len) memcpy(d, src, len);
    return v;
}
API napi_value mock_typedarray(napi_env env, int type, napi_value ab, size_t byte_offset, size_t elements) {
    napi_value v = new_value(env, K_TYPEDARRAY);
    v->ta_type = type;
    v->ab = ab;
    v->ta_off = byte_offset;
    v->data = ab->data + byte_offset;
    v->len = elements;
    return v;
}
API napi_value mock_buffer(napi_env env, const void* src, size_t len) {
    napi_value v = new_value(env, K_BUFFER);
    v->data = (uint8_t*)malloc(len ? len : 1);
    if (len) memcpy(v->data, src, len);
    v->len = len;
    return v;
}
API napi_value mock_function(napi_env env, mock_host_fn fn, void* data) {
    napi_value v = new_value(env, K_FUNCTION);
    v->host_fn = fn;
    v->host_data = data;
    return v;
}

API int mock_kind(napi_value v) { return v ? v->kind : -1; }
API double mock_num(napi_value v) { return v->num; }
API const char* mock_str(napi_value v) { return v->str; }
API const char* mock_code(napi_value v) { return v->code; }
API int mock_is_type_error(napi_value v) { return v->kind == K_ERROR && v->type_error; }
API const void* mock_data(napi_value v) { return v->data; }
API size_t mock_len(napi_value v) { return v->len; }
API napi_value mock_get(napi_value obj, const char* name) {
    prop* p = find_prop(obj, name);
    return p ? p->v : NULL;
}

/* calls a function value the way JS would: returns its result (undefined if it returned nothing), or NULL when it threw
 * -- the exception is then in mock_pending() until mock_clear_pending(). */
API napi_value mock_call(napi_env env, napi_value fn, size_t argc, napi_value* argv) {
    if (!fn || fn->kind != K_FUNCTION) return NULL;
    napi_value r = invoke(env, env->undefined, fn, argc, argv);
    if (env->pending) return NULL;
    return r ? r : env->undefined;
}
API napi_value mock_pending(napi_env env) { return env->pending; }
API void mock_clear_pending(napi_env env) { env->pending = NULL; }
API int mock_live_refs(napi_env env) { return env->live_refs; }
API int mock_queued(napi_env env) {
    int n = 0;
    for (napi_async_work w = env->queue; w; w = w->next) n++;
    return n;
}

static void* run_execute(void* p) {
    napi_async_work w = (napi_async_work)p;
    w->execute(w->env, w->data);
    return NULL;
}

/* the event loop's share: every queued work's execute() runs on its own thread, all at the same time (libuv's pool runs
 * four side by side), then the complete() callbacks run on the calling thread in queue order.  Returns the number run. */
API int mock_run_async(napi_env env) {
    napi_async_work q = env->queue;
    env->queue = NULL;
    int n = 0;
    for (napi_async_work w = q; w; w = w->next) n++;
    if (!n) return 0;
    pthread_t* th = (pthread_t*)calloc(n, sizeof *th);
    int i = 0;
    for (napi_async_work w = q; w; w = w->next) pthread_create(&th[i++], NULL, run_execute, w);
    for (i = 0; i < n; i++) pthread_join(th[i], NULL);
    free(th);
    for (napi_async_work w = q; w;) {
        napi_async_work nx = w->next;
        if (w->complete) w->complete(env, napi_ok, w->data);
        if (w->deleted) free(w);
        w = nx;
    }
    return n;
}
