/* TEST INFRASTRUCTURE: a declarations-only stand-in for Node's <node_api.h> (absent from the build image) so that
 * host/napi/addon.cc can be syntax- and type-checked (tests/test_host_shim.py: g++ -fsyntax-only).  Signatures follow the
 * published N-API (Node-API version 8) [external]; nothing here is linked or shipped. */
#pragma once
#include <stddef.h>
#include <stdint.h>

extern "C" {
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_async_work__* napi_async_work;
typedef struct napi_threadsafe_function__* napi_threadsafe_function;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure, napi_closing } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array, napi_uint32_array,
               napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef enum { napi_tsfn_release, napi_tsfn_abort } napi_threadsafe_function_release_mode;
typedef enum { napi_tsfn_nonblocking, napi_tsfn_blocking } napi_threadsafe_function_call_mode;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);
typedef void (*napi_threadsafe_function_call_js)(napi_env env, napi_value js_callback, void* context, void* data);
typedef struct {
    const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value;
    napi_property_attributes attributes; void* data;
} napi_property_descriptor;
#define NAPI_AUTO_LENGTH SIZE_MAX
#define NAPI_MODULE(modname, regfunc) napi_value napi_stub_register_##modname(napi_env env, napi_value exports) { return regfunc(env, exports); }

napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_named_property(napi_env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_define_properties(napi_env, napi_value object, size_t n, const napi_property_descriptor* properties);
napi_status napi_create_object(napi_env, napi_value* result);
napi_status napi_create_string_utf8(napi_env, const char* str, size_t length, napi_value* result);
napi_status napi_create_double(napi_env, double value, napi_value* result);
napi_status napi_create_int32(napi_env, int32_t value, napi_value* result);
napi_status napi_create_error(napi_env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_create_promise(napi_env, napi_deferred* deferred, napi_value* promise);
napi_status napi_resolve_deferred(napi_env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value* result);
napi_status napi_get_typedarray_info(napi_env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data, napi_value* arraybuffer,
                                     size_t* byte_offset);
napi_status napi_create_external(napi_env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_value_external(napi_env, napi_value value, void** result);
napi_status napi_get_value_int32(napi_env, napi_value value, int32_t* result);
napi_status napi_get_value_double(napi_env, napi_value value, double* result);
napi_status napi_get_value_bool(napi_env, napi_value value, bool* result);
napi_status napi_get_value_bigint_uint64(napi_env, napi_value value, uint64_t* result, bool* lossless);
napi_status napi_get_value_string_utf8(napi_env, napi_value value, char* buf, size_t bufsize, size_t* result);
napi_status napi_coerce_to_bool(napi_env, napi_value value, napi_value* result);
napi_status napi_typeof(napi_env, napi_value value, napi_valuetype* result);
napi_status napi_get_undefined(napi_env, napi_value* result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_call_function(napi_env, napi_value recv, napi_value func, size_t argc, const napi_value* argv, napi_value* result);
napi_status napi_create_async_work(napi_env, napi_value async_resource, napi_value async_resource_name, napi_async_execute_callback execute,
                                   napi_async_complete_callback complete, void* data, napi_async_work* result);
napi_status napi_queue_async_work(napi_env, napi_async_work work);
napi_status napi_delete_async_work(napi_env, napi_async_work work);
napi_status napi_create_threadsafe_function(napi_env, napi_value func, napi_value async_resource, napi_value async_resource_name, size_t max_queue_size,
                                            size_t initial_thread_count, void* thread_finalize_data, napi_finalize thread_finalize_cb, void* context,
                                            napi_threadsafe_function_call_js call_js_cb, napi_threadsafe_function* result);
napi_status napi_call_threadsafe_function(napi_threadsafe_function func, void* data, napi_threadsafe_function_call_mode is_blocking);
napi_status napi_release_threadsafe_function(napi_threadsafe_function func, napi_threadsafe_function_release_mode mode);
}
