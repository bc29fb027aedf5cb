// N-API shim: binds libbinder_b200.so's C ABI (include/binder_b200.h) into binder's Node.js
// process.  NOT compiled in this repository's image (no node / node_api.h here); it is kept
// thin enough to be read against the header.  Build (on a host with Node >= 10):
//   g++ -std=c++17 -shared -fPIC -I$(node -p "process.execPath+'/../../include/node'") \
//       -I../include binder_b200_napi.cc -L../binder_b200 -lbinder_b200 -o binder_b200.node
//
// JS surface:
//   const bb = require('./binder_b200.node');
//   const zone   = bb.zoneBuild(snapshotBuffer, 'dc1.example.com');
//   const engine = bb.engineCreate({dnsDomain, datacenterName, recursion: true, device: 0});
//   bb.engineSwapZone(engine, zone);
//   const r = bb.resolveBatch(engine, pkts /*Buffer*/, pktOff /*Uint32Array n+1*/, seed /*BigInt*/);
//   // r = {out: Buffer, outOff: Uint32Array, outLen: Uint16Array, status: Uint8Array, miss: Uint32Array}
#include <node_api.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include "binder_b200.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, nullptr, #call); return nullptr; } } while (0)

static napi_value Throw(napi_env env, int err) {
    napi_throw_error(env, nullptr, err == BB_ERR_CUDA ? bb_last_cuda_error() : bb_strerror(err));
    return nullptr;
}

static napi_value ZoneBuild(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    void* data; size_t len; NAPI_OK(napi_get_buffer_info(env, argv[0], &data, &len));
    char dom[256]; size_t dl; NAPI_OK(napi_get_value_string_utf8(env, argv[1], dom, sizeof dom, &dl));
    int err = 0;
    bb_zone* z = bb_zone_build((const char*)data, len, dom, &err);
    if (!z) return Throw(env, err);
    napi_value ext;
    NAPI_OK(napi_create_external(env, z, [](napi_env, void* p, void*) { bb_zone_free((bb_zone*)p); }, nullptr, &ext));
    return ext;
}

static napi_value EngineCreate(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    char dom[256] = "", dc[256] = ""; size_t n; napi_value v; bool rec = false; int32_t dev = 0; uint32_t maxb = 1u << 16;
    NAPI_OK(napi_get_named_property(env, argv[0], "dnsDomain", &v)); NAPI_OK(napi_get_value_string_utf8(env, v, dom, sizeof dom, &n));
    if (napi_get_named_property(env, argv[0], "datacenterName", &v) == napi_ok) napi_get_value_string_utf8(env, v, dc, sizeof dc, &n);
    if (napi_get_named_property(env, argv[0], "recursion", &v) == napi_ok) napi_get_value_bool(env, v, &rec);
    if (napi_get_named_property(env, argv[0], "device", &v) == napi_ok) napi_get_value_int32(env, v, &dev);
    if (napi_get_named_property(env, argv[0], "maxBatch", &v) == napi_ok) napi_get_value_uint32(env, v, &maxb);
    bb_engine_opts o; memset(&o, 0, sizeof o);
    o.dns_domain = dom; o.datacenter_name = dc; o.recursion = rec; o.device = dev; o.max_batch = maxb;
    int err = 0;
    bb_engine* e = bb_engine_create(&o, &err);
    if (!e) return Throw(env, err);
    napi_value ext;
    NAPI_OK(napi_create_external(env, e, [](napi_env, void* p, void*) { bb_engine_destroy((bb_engine*)p); }, nullptr, &ext));
    return ext;
}

static napi_value EngineSwapZone(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    void *e, *z; NAPI_OK(napi_get_value_external(env, argv[0], &e)); NAPI_OK(napi_get_value_external(env, argv[1], &z));
    int rc = bb_engine_swap_zone((bb_engine*)e, (bb_zone*)z);
    if (rc != BB_OK) return Throw(env, rc);
    return nullptr;
}

// zoneApply(zone, delta: Buffer): watch events as JSON lines (bb_zone_apply)
static napi_value ZoneApply(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    void* z; NAPI_OK(napi_get_value_external(env, argv[0], &z));
    void* data; size_t len; NAPI_OK(napi_get_buffer_info(env, argv[1], &data, &len));
    int rc = bb_zone_apply((bb_zone*)z, (const char*)data, len);
    if (rc != BB_OK) return Throw(env, rc);
    return nullptr;
}

// engineApplyUpdate(engine, zone): ship what the deltas changed (bb_engine_apply_update)
static napi_value EngineApplyUpdate(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    void *e, *z; NAPI_OK(napi_get_value_external(env, argv[0], &e)); NAPI_OK(napi_get_value_external(env, argv[1], &z));
    int rc = bb_engine_apply_update((bb_engine*)e, (bb_zone*)z);
    if (rc != BB_OK) return Throw(env, rc);
    return nullptr;
}

// engineSetRecursionFilter(engine, regionDomain: string | null, dcNames: string[], ptrForwardable: boolean)
static napi_value EngineSetRecursionFilter(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    void* e; NAPI_OK(napi_get_value_external(env, argv[0], &e));
    napi_valuetype t; NAPI_OK(napi_typeof(env, argv[1], &t));
    if (t != napi_string) { int rc0 = bb_engine_set_recursion_filter((bb_engine*)e, nullptr, nullptr, 0, 0); return rc0 == BB_OK ? nullptr : Throw(env, rc0); }
    char dom[256]; size_t dl; NAPI_OK(napi_get_value_string_utf8(env, argv[1], dom, sizeof dom, &dl));
    uint32_t n = 0; NAPI_OK(napi_get_array_length(env, argv[2], &n));
    if (n > 16) return Throw(env, BB_ERR_ARG);
    char names[16][64]; const char* ptrs[16];
    for (uint32_t i = 0; i < n; i++) {
        napi_value v; size_t l; NAPI_OK(napi_get_element(env, argv[2], i, &v));
        NAPI_OK(napi_get_value_string_utf8(env, v, names[i], sizeof names[i], &l)); ptrs[i] = names[i];
    }
    bool ptr = false; napi_get_value_bool(env, argv[3], &ptr);
    int rc = bb_engine_set_recursion_filter((bb_engine*)e, dom, ptrs, n, ptr ? 1 : 0);
    if (rc != BB_OK) return Throw(env, rc);
    return nullptr;
}

// resolveBatch(engine, pkts: Buffer, pktOff: Uint32Array(n+1), seed: BigInt[, tcp: boolean]) -> result object
static napi_value ResolveBatch(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value argv[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    bool tcp = false;                              // optional 5th argument: the batch arrived over TCP
    if (argc > 4) napi_get_value_bool(env, argv[4], &tcp);
    void* e; NAPI_OK(napi_get_value_external(env, argv[0], &e));
    void* pk; size_t pklen; NAPI_OK(napi_get_buffer_info(env, argv[1], &pk, &pklen));
    napi_typedarray_type tt; size_t offn; void* offp; napi_value ab; size_t bo;
    NAPI_OK(napi_get_typedarray_info(env, argv[2], &tt, &offn, &offp, &ab, &bo));
    uint64_t seed = 0; bool lossless; napi_get_value_bigint_uint64(env, argv[3], &seed, &lossless);
    // pktOff: at least one entry, ascending, ending inside the pkts Buffer (the engine trusts only what it can check cheaply)
    if (tt != napi_uint32_array || offn < 1 || offn > (1u << 22) + 1) return Throw(env, BB_ERR_ARG);
    {
        const uint32_t* po = (const uint32_t*)offp;
        for (size_t i = 1; i < offn; i++) if (po[i] < po[i - 1]) return Throw(env, BB_ERR_ARG);
        if (po[offn - 1] > pklen) return Throw(env, BB_ERR_ARG);
    }
    const uint32_t n = (uint32_t)offn - 1;
    const uint32_t cap = n * (tcp ? 16384u : 1232u);
    void *out, *oo, *ol, *st, *ms; napi_value o_out, a_oo, a_ol, a_st, a_ms, t_oo, t_ol, t_st, t_ms;
    NAPI_OK(napi_create_buffer(env, cap, &out, &o_out));
    NAPI_OK(napi_create_arraybuffer(env, (n + 1) * 4, &oo, &a_oo)); NAPI_OK(napi_create_typedarray(env, napi_uint32_array, n + 1, a_oo, 0, &t_oo));
    NAPI_OK(napi_create_arraybuffer(env, n * 2, &ol, &a_ol));       NAPI_OK(napi_create_typedarray(env, napi_uint16_array, n, a_ol, 0, &t_ol));
    NAPI_OK(napi_create_arraybuffer(env, n, &st, &a_st));           NAPI_OK(napi_create_typedarray(env, napi_uint8_array, n, a_st, 0, &t_st));
    NAPI_OK(napi_create_arraybuffer(env, n * 4, &ms, &a_ms));
    uint32_t nmiss = 0;
    int rc = bb_resolve_batch_ex((bb_engine*)e, (const uint8_t*)pk, (const uint32_t*)offp, n, seed, 0, (uint8_t*)out, cap,
                                 (uint32_t*)oo, (uint16_t*)ol, (uint8_t*)st, (uint32_t*)ms, &nmiss, tcp ? BB_BATCH_TCP : 0u);
    if (rc != BB_OK) return Throw(env, rc);
    NAPI_OK(napi_create_typedarray(env, napi_uint32_array, nmiss, a_ms, 0, &t_ms));
    napi_value res; NAPI_OK(napi_create_object(env, &res));
    napi_set_named_property(env, res, "out", o_out);   napi_set_named_property(env, res, "outOff", t_oo);
    napi_set_named_property(env, res, "outLen", t_ol); napi_set_named_property(env, res, "status", t_st);
    napi_set_named_property(env, res, "miss", t_ms);
    return res;
}

static napi_value Init(napi_env env, napi_value exports) {
    napi_property_descriptor d[] = {
        { "zoneBuild", nullptr, ZoneBuild, nullptr, nullptr, nullptr, napi_default, nullptr },
        { "engineCreate", nullptr, EngineCreate, nullptr, nullptr, nullptr, napi_default, nullptr },
        { "engineSwapZone", nullptr, EngineSwapZone, nullptr, nullptr, nullptr, napi_default, nullptr },
        { "zoneApply", nullptr, ZoneApply, nullptr, nullptr, nullptr, napi_default, nullptr },
        { "engineApplyUpdate", nullptr, EngineApplyUpdate, nullptr, nullptr, nullptr, napi_default, nullptr },
        { "engineSetRecursionFilter", nullptr, EngineSetRecursionFilter, nullptr, nullptr, nullptr, napi_default, nullptr },
        { "resolveBatch", nullptr, ResolveBatch, nullptr, nullptr, nullptr, napi_default, nullptr },
    };
    napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
