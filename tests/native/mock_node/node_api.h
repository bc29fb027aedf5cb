// A minimal stand-in for Node's <node_api.h>: just the declarations addon/binder_b200_napi.cc uses, with the
// signatures of the N-API documentation.  Node.js is absent from this image; compiling the addon against this
// header (tests/test_abi.py) catches drift between the addon and include/binder_b200.h.  Test infrastructure only.
#ifndef BB_MOCK_NODE_API_H
#define BB_MOCK_NODE_API_H
#include <stddef.h>
#include <stdint.h>
extern "C" {
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_generic_failure } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array, napi_uint32_array,
               napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef struct { const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value;
                 napi_property_attributes attributes; void* data; } napi_property_descriptor;
#define NAPI_AUTO_LENGTH SIZE_MAX
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char* buf, size_t bufsize, size_t* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value* result);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_get_value_bool(napi_env env, napi_value value, bool* result);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_bigint_uint64(napi_env env, napi_value value, uint64_t* result, bool* lossless);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t* result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor* properties);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_create_buffer(napi_env env, size_t length, void** data, napi_value* result);
}
#define NODE_GYP_MODULE_NAME binder_b200
#define NAPI_MODULE(modname, regfunc) extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports) { return regfunc(env, exports); }
#endif
