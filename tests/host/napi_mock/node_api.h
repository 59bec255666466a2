/* Stand-in for Node's <node_api.h> (the N-API v6 subset node/wasmsnark_b200_addon.c uses) for an image without Node.js and
 * without network access to fetch the real header.  Signatures, enum values and the module-initialiser name follow the
 * public N-API documentation.  tests/test_abi.py type-checks the addon against it; tests/host/napi_mock/napi_mock.c
 * implements these functions over a toy value model so that the addon can be LINKED AND EXECUTED by
 * tests/test_node_addon_mock.py (a mock runtime, not Node: it proves the binding's own logic, not V8 behaviour). */
#ifndef WS_TEST_NODE_API_STUB_H
#define WS_TEST_NODE_API_STUB_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_ref__* napi_ref;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_async_work__* napi_async_work;
typedef enum {
    napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected, napi_function_expected, napi_number_expected,
    napi_boolean_expected, napi_array_expected, napi_generic_failure, napi_pending_exception, napi_cancelled, napi_escape_called_twice,
    napi_handle_scope_mismatch, napi_callback_scope_mismatch, napi_queue_full, napi_closing, napi_bigint_expected, napi_date_expected,
    napi_arraybuffer_expected, napi_detachable_arraybuffer_expected, napi_would_deadlock
} napi_status;
typedef enum { napi_default = 0, napi_writable = 1, napi_enumerable = 2, napi_configurable = 4, napi_static = 1024 } napi_property_attributes;
typedef enum {
    napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array, napi_uint32_array,
    napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);
typedef struct {
    const char* utf8name;
    napi_value name;
    napi_callback method, getter, setter;
    napi_value value;
    napi_property_attributes attributes;
    void* data;
} napi_property_descriptor;

#define NAPI_AUTO_LENGTH SIZE_MAX
/* the real macro also registers the module with node; the entry point it defines has this name and signature */
#define NAPI_MODULE_INIT() __attribute__((visibility("default"))) napi_value napi_register_module_v1(napi_env env, napi_value exports)

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref* result);
napi_status napi_delete_reference(napi_env env, napi_ref ref);
napi_status napi_get_reference_value(napi_env env, napi_ref ref, napi_value* result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_null(napi_env env, napi_value* result);
napi_status napi_get_global(napi_env env, napi_value* result);
napi_status napi_get_boolean(napi_env env, bool value, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name, napi_async_execute_callback execute,
                                   napi_async_complete_callback complete, void* data, napi_async_work* result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool* result);
napi_status napi_is_buffer(napi_env env, napi_value value, bool* result);
napi_status napi_is_arraybuffer(napi_env env, napi_value value, bool* result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length);
napi_status napi_get_arraybuffer_info(napi_env env, napi_value arraybuffer, void** data, size_t* byte_length);
napi_status napi_call_function(napi_env env, napi_value recv, napi_value func, size_t argc, const napi_value* argv, napi_value* result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor* properties);
#endif
