// N-API shim: the JavaScript binding of include/olap_gpu.h.
//
// UNEXECUTED in this repository: the build image has no Node.js and no node_api.h
// (SURVEY.md F4).  It is shipped so a maintainer with Node 20 can `node-gyp rebuild` it
// (addon/binding.gyp) next to a built libolapgpu.so.  By design it holds no logic: it
// unpacks arguments, extracts TypedArray pointers, calls the C ABI and turns an error
// code into a thrown JS Error carrying olap_last_error() — the texts are the reference's
// own (`value length is invalid: a !== b`, `Unsupported aggregation method: x`, ...).
#include <node_api.h>

#include <vector>

#include "../include/olap_gpu.h"

#define NAPI_OK(call) \
    if ((call) != napi_ok) { napi_throw_error(env, nullptr, "N-API call failed: " #call); return nullptr; }

static napi_value throw_last(napi_env env) {
    napi_throw_error(env, nullptr, olap_last_error());
    return nullptr;
}
#define OLAP_CALL(expr) \
    if ((expr) != OLAP_OK) return throw_last(env)

static void finalize_store(napi_env env, void* data, void*) {
    olap_store* s = static_cast<olap_store*>(data);
    int64_t freed = -(olap_store_size(s) * (olap_store_has_status(s) ? 5 : 4));
    int64_t adjusted;
    napi_adjust_external_memory(env, freed, &adjusted);  // let V8's GC see HBM pressure
    olap_store_destroy(s);
}

static napi_value wrap_store(napi_env env, olap_store* s) {
    napi_value ext;
    NAPI_OK(napi_create_external(env, s, finalize_store, nullptr, &ext));
    int64_t adjusted;
    napi_adjust_external_memory(env, olap_store_size(s) * (olap_store_has_status(s) ? 5 : 4), &adjusted);
    return ext;
}

static olap_store* unwrap(napi_env env, napi_value v) {
    void* p = nullptr;
    napi_get_value_external(env, v, &p);
    return static_cast<olap_store*>(p);
}

struct Args {
    napi_value v[10];
    size_t n = 10;
};
#define GET_ARGS() \
    Args a;        \
    NAPI_OK(napi_get_cb_info(env, info, &a.n, a.v, nullptr, nullptr))

static std::vector<olap_store*> store_list(napi_env env, napi_value arr) {
    uint32_t n = 0;
    napi_get_array_length(env, arr, &n);
    std::vector<olap_store*> out(n);
    for (uint32_t i = 0; i < n; ++i) {
        napi_value e;
        napi_get_element(env, arr, i, &e);
        out[i] = unwrap(env, e);
    }
    return out;
}

static std::vector<int64_t> int64_list(napi_env env, napi_value arr) {
    uint32_t n = 0;
    napi_get_array_length(env, arr, &n);
    std::vector<int64_t> out(n);
    for (uint32_t i = 0; i < n; ++i) {
        napi_value e;
        napi_get_element(env, arr, i, &e);
        napi_get_value_int64(env, e, &out[i]);
    }
    return out;
}

// Array of Int32Array -> const int32_t* const*
static std::vector<const int32_t*> map_list(napi_env env, napi_value arr) {
    uint32_t n = 0;
    napi_get_array_length(env, arr, &n);
    std::vector<const int32_t*> out(n);
    for (uint32_t i = 0; i < n; ++i) {
        napi_value e;
        napi_get_element(env, arr, i, &e);
        napi_typedarray_type t;
        size_t len;
        void* data;
        napi_get_typedarray_info(env, e, &t, &len, &data, nullptr, nullptr);
        out[i] = static_cast<const int32_t*>(data);
    }
    return out;
}

static napi_value store_array_out(napi_env env, std::vector<olap_store*>& out) {
    napi_value arr;
    NAPI_OK(napi_create_array_with_length(env, out.size(), &arr));
    for (size_t i = 0; i < out.size(); ++i) napi_set_element(env, arr, i, wrap_store(env, out[i]));
    return arr;
}

// create(size, type, defaultKind, withStatus) -> external
static napi_value Create(napi_env env, napi_callback_info info) {
    GET_ARGS();
    int64_t size;
    int32_t type, kind;
    bool status;
    napi_get_value_int64(env, a.v[0], &size);
    napi_get_value_int32(env, a.v[1], &type);
    napi_get_value_int32(env, a.v[2], &kind);
    napi_get_value_bool(env, a.v[3], &status);
    olap_store* s = nullptr;
    OLAP_CALL(olap_store_create(size, type, kind, status, &s));
    return wrap_store(env, s);
}

// upload(store, Float32Array | Float64Array)
static napi_value Upload(napi_env env, napi_callback_info info) {
    GET_ARGS();
    napi_typedarray_type t;
    size_t len;
    void* data;
    NAPI_OK(napi_get_typedarray_info(env, a.v[1], &t, &len, &data, nullptr, nullptr));
    if (t == napi_float32_array) OLAP_CALL(olap_store_upload_f32(unwrap(env, a.v[0]), static_cast<float*>(data), (int64_t)len));
    else OLAP_CALL(olap_store_upload_f64(unwrap(env, a.v[0]), static_cast<double*>(data), (int64_t)len));
    return nullptr;
}

// download(store, Float32Array | Float64Array)
static napi_value Download(napi_env env, napi_callback_info info) {
    GET_ARGS();
    napi_typedarray_type t;
    size_t len;
    void* data;
    NAPI_OK(napi_get_typedarray_info(env, a.v[1], &t, &len, &data, nullptr, nullptr));
    if (t == napi_float32_array) OLAP_CALL(olap_store_download_f32(unwrap(env, a.v[0]), static_cast<float*>(data), (int64_t)len));
    else OLAP_CALL(olap_store_download_f64(unwrap(env, a.v[0]), static_cast<double*>(data), (int64_t)len));
    return nullptr;
}

// drillUp(stores[], methods Int32Array, oldLen[], newLen[], maps Int32Array[]) -> stores[]
static napi_value DrillUp(napi_env env, napi_callback_info info) {
    GET_ARGS();
    auto src = store_list(env, a.v[0]);
    napi_typedarray_type t;
    size_t len;
    void* mdata;
    NAPI_OK(napi_get_typedarray_info(env, a.v[1], &t, &len, &mdata, nullptr, nullptr));
    auto old_len = int64_list(env, a.v[2]), new_len = int64_list(env, a.v[3]);
    auto maps = map_list(env, a.v[4]);
    std::vector<olap_store*> out(src.size());
    OLAP_CALL(olap_drill_up(src.data(), (int)src.size(), static_cast<const int*>(mdata), (int)old_len.size(),
                            old_len.data(), new_len.data(), maps.data(), out.data()));
    return store_array_out(env, out);
}

// dice(stores[], oldLen[], newLen[], keep Int32Array[]) -> stores[]
static napi_value Dice(napi_env env, napi_callback_info info) {
    GET_ARGS();
    auto src = store_list(env, a.v[0]);
    auto old_len = int64_list(env, a.v[1]), new_len = int64_list(env, a.v[2]);
    auto keep = map_list(env, a.v[3]);
    std::vector<olap_store*> out(src.size());
    OLAP_CALL(olap_dice(src.data(), (int)src.size(), (int)old_len.size(), old_len.data(), new_len.data(), keep.data(), out.data()));
    return store_array_out(env, out);
}

// reorder(stores[], oldLen[], newToOld Int32Array) -> stores[]
static napi_value Reorder(napi_env env, napi_callback_info info) {
    GET_ARGS();
    auto src = store_list(env, a.v[0]);
    auto old_len = int64_list(env, a.v[1]);
    napi_typedarray_type t;
    size_t len;
    void* perm;
    NAPI_OK(napi_get_typedarray_info(env, a.v[2], &t, &len, &perm, nullptr, nullptr));
    std::vector<olap_store*> out(src.size());
    OLAP_CALL(olap_reorder(src.data(), (int)src.size(), (int)old_len.size(), old_len.data(), static_cast<const int32_t*>(perm), out.data()));
    return store_array_out(env, out);
}

// total(store) -> number
static napi_value Total(napi_env env, napi_callback_info info) {
    GET_ARGS();
    double v;
    OLAP_CALL(olap_store_total(unwrap(env, a.v[0]), &v));
    napi_value out;
    NAPI_OK(napi_create_double(env, v, &out));
    return out;
}

// drillDown, load, eval, presence, exportSparse, setValue(s), fill, clone follow the same
// pattern (unpack -> one olap_* call -> wrap) and are omitted from this excerpt only for
// length; INTEGRATION.md lists the full table of bindings.

static napi_value Init(napi_env env, napi_value exports) {
    const napi_property_descriptor props[] = {
        {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"upload", nullptr, Upload, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"download", nullptr, Download, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"drillUp", nullptr, DrillUp, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"dice", nullptr, Dice, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"reorder", nullptr, Reorder, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"total", nullptr, Total, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
    if (olap_init(0) != OLAP_OK) return throw_last(env);  // no CPU fallback: fail at require() time
    return exports;
}
NAPI_MODULE(olap_gpu, Init)
