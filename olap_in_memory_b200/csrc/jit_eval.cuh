// Computed measures (cube.js:331-363, parser.js:3-26): the formula arrives in postfix
// text, is lowered to ONE fused elementwise CUDA kernel evaluating in double on the
// float32 cells of its inputs, compiled for sm_100a with NVRTC and cached per
// (program, input count, total count).  libnvrtc and libcuda are resolved with dlopen
// so the library loads on machines without a driver (where every call then fails).
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <sstream>

#include "common.cuh"

namespace olap {

struct JitApi {
    bool tried = false, ok = false;
    std::string why;
    // nvrtc
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    // driver
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void**, void**);
    CUresult (*GetErrorString)(CUresult, const char**);
};

inline JitApi& jit_api() {
    static JitApi api;
    if (api.tried) return api;
    api.tried = true;
    void* rtc = nullptr;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
        rtc = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (rtc) break;
    }
    if (!rtc) { api.why = "libnvrtc.so.12 not found"; return api; }
    void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!drv) { api.why = "libcuda.so.1 not found (no NVIDIA driver)"; return api; }
#define OLAP_SYM(lib, field, sym)                                         \
    *(void**)(&api.field) = dlsym(lib, sym);                              \
    if (!api.field) { api.why = std::string("missing symbol ") + sym; return api; }
    OLAP_SYM(rtc, CreateProgram, "nvrtcCreateProgram")
    OLAP_SYM(rtc, CompileProgram, "nvrtcCompileProgram")
    OLAP_SYM(rtc, GetCUBINSize, "nvrtcGetCUBINSize")
    OLAP_SYM(rtc, GetCUBIN, "nvrtcGetCUBIN")
    OLAP_SYM(rtc, GetProgramLogSize, "nvrtcGetProgramLogSize")
    OLAP_SYM(rtc, GetProgramLog, "nvrtcGetProgramLog")
    OLAP_SYM(rtc, DestroyProgram, "nvrtcDestroyProgram")
    OLAP_SYM(drv, ModuleLoadData, "cuModuleLoadData")
    OLAP_SYM(drv, ModuleGetFunction, "cuModuleGetFunction")
    OLAP_SYM(drv, LaunchKernel, "cuLaunchKernel")
    OLAP_SYM(drv, GetErrorString, "cuGetErrorString")
#undef OLAP_SYM
    api.ok = true;
    return api;
}

// ---- postfix -> C expression ----------------------------------------------------
struct Lowered {
    std::string expr;
    int max_input = -1, max_total = -1;
};

inline bool lower_program(const char* program, Lowered& out, std::string& err) {
    std::vector<std::string> st;
    std::istringstream in(program ? program : "");
    std::string tok;
    auto pop = [&](std::string& s) {
        if (st.empty()) return false;
        s = st.back();
        st.pop_back();
        return true;
    };
    static const std::map<std::string, std::string> unary = {
        {"abs", "fabs"}, {"ceil", "ceil"}, {"floor", "floor"}, {"trunc", "trunc"}, {"sqrt", "sqrt"}, {"cbrt", "cbrt"},
        {"exp", "exp"}, {"expm1", "expm1"}, {"ln", "log"}, {"log", "log"}, {"log1p", "log1p"}, {"log2", "log2"},
        {"log10", "log10"}, {"lg", "log10"}, {"sin", "sin"}, {"cos", "cos"}, {"tan", "tan"}, {"asin", "asin"},
        {"acos", "acos"}, {"atan", "atan"}, {"sinh", "sinh"}, {"cosh", "cosh"}, {"tanh", "tanh"}, {"asinh", "asinh"},
        {"acosh", "acosh"}, {"atanh", "atanh"}, {"sign", "olap_sign"}, {"round", "olap_round"}, {"isNaN", "olap_isnan"}};
    while (in >> tok) {
        if (tok[0] == 'v' && tok.size() > 1 && isdigit((unsigned char)tok[1])) {
            int k = atoi(tok.c_str() + 1);
            if (k > out.max_input) out.max_input = k;
            st.push_back("x" + std::to_string(k));
        } else if (tok[0] == 't' && tok.size() > 1 && isdigit((unsigned char)tok[1])) {
            int k = atoi(tok.c_str() + 1);
            if (k > out.max_total) out.max_total = k;
            st.push_back("t" + std::to_string(k));
        } else if (tok[0] == '#') {
            const std::string num = tok.substr(1);
            if (num == "nan") st.push_back("olap_nan()");
            else if (num == "inf") st.push_back("olap_inf()");
            else if (num == "-inf") st.push_back("(-olap_inf())");
            else {
                char* end = nullptr;
                const double v = strtod(num.c_str(), &end);
                if (!end || *end) { err = "bad number token: " + tok; return false; }
                char buf[64];
                snprintf(buf, sizeof buf, "%.17g", v);
                std::string lit = buf;
                if (lit.find_first_of(".eEn") == std::string::npos) lit += ".0";
                st.push_back("(" + lit + ")");
            }
        } else if (tok == "neg") {
            std::string a;
            if (!pop(a)) { err = "stack underflow at neg"; return false; }
            st.push_back("(-" + a + ")");
        } else if (tok == "+" || tok == "-" || tok == "*" || tok == "/" || tok == "%" || tok == "^" || tok == "||") {
            std::string a, b;
            if (!pop(b) || !pop(a)) { err = "stack underflow at " + tok; return false; }
            if (tok == "%") st.push_back("fmod(" + a + ", " + b + ")");
            else if (tok == "^") st.push_back("olap_pow(" + a + ", " + b + ")");
            else if (tok == "||") st.push_back("olap_coalesce_add(" + a + ", " + b + ")");
            else st.push_back("(" + a + " " + tok + " " + b + ")");
        } else if (tok == "?:") {
            std::string c, a, b;
            if (!pop(b) || !pop(a) || !pop(c)) { err = "stack underflow at ?:"; return false; }
            st.push_back("(olap_truthy(" + c + ") ? " + a + " : " + b + ")");
        } else if (tok.rfind("call:", 0) == 0) {
            const size_t c2 = tok.rfind(':');
            if (c2 <= 5) { err = "bad call token: " + tok; return false; }
            const std::string name = tok.substr(5, c2 - 5);
            const int argc = atoi(tok.c_str() + c2 + 1);
            if (argc < 1 || (int)st.size() < argc) { err = "stack underflow at " + tok; return false; }
            std::vector<std::string> args(st.end() - argc, st.end());
            st.resize(st.size() - argc);
            auto u = unary.find(name);
            if (u != unary.end() && argc == 1) st.push_back(u->second + "(" + args[0] + ")");
            else if (name == "min" || name == "max") {
                // a single argument still goes through the helper so NaN stays NaN
                std::string e = argc == 1 ? "olap_" + name + "(" + args[0] + ", " + args[0] + ")" : args[0];
                for (int k = 1; k < argc; ++k) e = "olap_" + name + "(" + e + ", " + args[k] + ")";
                st.push_back(e);
            } else if (name == "hypot") {
                std::string e = "sqrt(";
                for (int k = 0; k < argc; ++k) e += (k ? " + " : "") + ("(" + args[k] + ")*(" + args[k] + ")");
                st.push_back(e + ")");
            } else if (name == "pow" && argc == 2) st.push_back("olap_pow(" + args[0] + ", " + args[1] + ")");
            else if (name == "atan2" && argc == 2) st.push_back("atan2(" + args[0] + ", " + args[1] + ")");
            else if (name == "roundTo" && argc == 2)
                st.push_back("(olap_round(" + args[0] + " * pow(10.0, " + args[1] + ")) / pow(10.0, " + args[1] + "))");
            else if (name == "if" && argc == 3)
                st.push_back("(olap_truthy(" + args[0] + ") ? " + args[1] + " : " + args[2] + ")");
            else { err = "unsupported function in formula: " + name; return false; }
        } else {
            err = "unknown token in formula program: " + tok;
            return false;
        }
    }
    if (st.size() != 1) { err = "formula program does not reduce to one value"; return false; }
    out.expr = st[0];
    return true;
}

inline std::string eval_source(const Lowered& lw, int n_inputs, int n_totals) {
    std::ostringstream s;
    s << "typedef long long i64;\n"
         "__device__ __forceinline__ double olap_nan() { return __longlong_as_double(0x7ff8000000000000LL); }\n"
         "__device__ __forceinline__ double olap_inf() { return __longlong_as_double(0x7ff0000000000000LL); }\n"
         "__device__ __forceinline__ bool olap_truthy(double c) { return c == c && c != 0.0; }\n"
         "__device__ __forceinline__ double olap_isnan(double a) { return a != a ? 1.0 : 0.0; }\n"
         "__device__ __forceinline__ double olap_round(double a) { return floor(a + 0.5); }\n"
         "__device__ __forceinline__ double olap_sign(double a) { return a != a ? a : (a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : a)); }\n"
         // parser.js:18-23
         "__device__ __forceinline__ double olap_coalesce_add(double a, double b) {\n"
         "  if (a != a && b == b) return b; if (a == a && b != b) return a; return a + b; }\n"
         // Math.max / Math.min / Math.pow semantics
         "__device__ __forceinline__ double olap_max(double a, double b) {\n"
         "  if (a != a || b != b) return olap_nan();\n"
         "  if (a == b) return __longlong_as_double(__double_as_longlong(a) & __double_as_longlong(b));\n"
         "  return a > b ? a : b; }\n"
         "__device__ __forceinline__ double olap_min(double a, double b) {\n"
         "  if (a != a || b != b) return olap_nan();\n"
         "  if (a == b) return __longlong_as_double(__double_as_longlong(a) | __double_as_longlong(b));\n"
         "  return a < b ? a : b; }\n"
         "__device__ __forceinline__ double olap_pow(double a, double b) {\n"
         "  if (b != b) return olap_nan(); if (b == 0.0) return 1.0;\n"
         "  if (fabs(a) == 1.0 && isinf(b)) return olap_nan();\n"
         // integer exponents by squaring: exact whenever the result is representable
         // (CUDA pow() is only 2-ulp accurate, which breaks e.g. (7 ^ 2) % 7)
         "  if (b == trunc(b) && fabs(b) <= 1024.0) { double r = 1.0, x = a; long long e = (long long)fabs(b);\n"
         "    while (e) { if (e & 1) r *= x; x *= x; e >>= 1; } return b < 0.0 ? 1.0 / r : r; }\n"
         "  return pow(a, b); }\n"
         "__device__ __forceinline__ float olap_canon(float v, int nan_default) {\n"
         "  if (v != v) return __int_as_float(0x7fc00000); if (!nan_default && v == 0.0f) return 0.0f; return v; }\n";
    s << "__device__ __forceinline__ double olap_formula(";
    bool first = true;
    for (int k = 0; k < n_inputs; ++k) { s << (first ? "" : ", ") << "double x" << k; first = false; }
    for (int k = 0; k < n_totals; ++k) { s << (first ? "" : ", ") << "double t" << k; first = false; }
    s << ") { return " << lw.expr << "; }\n";
    s << "extern \"C\" __global__ void __launch_bounds__(256) olap_eval_kernel(";
    for (int k = 0; k < n_inputs; ++k) s << "const float* __restrict__ v" << k << ", ";
    for (int k = 0; k < n_totals; ++k) s << "double t" << k << ", ";
    s << "float* __restrict__ out32, unsigned char* __restrict__ st_out, double* __restrict__ out64, i64 n, "
         "int nan_default) {\n"
         "  const i64 stride = (i64)gridDim.x * blockDim.x * 4;\n"
         "  for (i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {\n"
         "    if (i + 4 <= n) {\n";
    for (int k = 0; k < n_inputs; ++k)
        s << "      const float4 a" << k << " = __ldcs(reinterpret_cast<const float4*>(v" << k << " + i));\n";
    const char* comp[4] = {"x", "y", "z", "w"};
    for (int e = 0; e < 4; ++e) {
        s << "      const double r" << e << " = olap_formula(";
        first = true;
        for (int k = 0; k < n_inputs; ++k) { s << (first ? "" : ", ") << "(double)a" << k << "." << comp[e]; first = false; }
        for (int k = 0; k < n_totals; ++k) { s << (first ? "" : ", ") << "t" << k; first = false; }
        s << ");\n";
    }
    s << "      if (out64) { double2 lo = make_double2(r0, r1), hi = make_double2(r2, r3);\n"
         "        __stcs(reinterpret_cast<double2*>(out64 + i), lo); __stcs(reinterpret_cast<double2*>(out64 + i + 2), hi); }\n"
         "      if (out32) { float4 o = make_float4(olap_canon((float)r0, nan_default), olap_canon((float)r1, nan_default),\n"
         "                                          olap_canon((float)r2, nan_default), olap_canon((float)r3, nan_default));\n"
         "        __stcs(reinterpret_cast<float4*>(out32 + i), o);\n"
         "        if (st_out) { uchar4 q;\n"
         "          q.x = (nan_default ? o.x == o.x : o.x != 0.0f) ? 2 : 1; q.y = (nan_default ? o.y == o.y : o.y != 0.0f) ? 2 : 1;\n"
         "          q.z = (nan_default ? o.z == o.z : o.z != 0.0f) ? 2 : 1; q.w = (nan_default ? o.w == o.w : o.w != 0.0f) ? 2 : 1;\n"
         "          *reinterpret_cast<uchar4*>(st_out + i) = q; } }\n"
         "    } else {\n"
         "      for (i64 j = i; j < n; ++j) {\n"
         "        const double r = olap_formula(";
    first = true;
    for (int k = 0; k < n_inputs; ++k) { s << (first ? "" : ", ") << "(double)v" << k << "[j]"; first = false; }
    for (int k = 0; k < n_totals; ++k) { s << (first ? "" : ", ") << "t" << k; first = false; }
    s << ");\n"
         "        if (out64) out64[j] = r;\n"
         "        if (out32) { const float o = olap_canon((float)r, nan_default); out32[j] = o;\n"
         "          if (st_out) st_out[j] = (nan_default ? o == o : o != 0.0f) ? 2 : 1; }\n"
         "      }\n"
         "    }\n"
         "  }\n"
         "}\n";
    return s.str();
}

struct JitKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
};

inline int jit_get(const char* program, int n_inputs, int n_totals, JitKernel** out) {
    static std::map<std::string, JitKernel> cache;
    const std::string key = std::to_string(n_inputs) + "|" + std::to_string(n_totals) + "|" + (program ? program : "");
    auto it = cache.find(key);
    if (it != cache.end()) { *out = &it->second; return OLAP_OK; }

    Lowered lw;
    std::string err;
    if (!lower_program(program, lw, err)) return fail(OLAP_E_INVALID, "%s", err.c_str());
    if (lw.max_input >= n_inputs) return fail(OLAP_E_INVALID, "formula reads input %d but only %d given", lw.max_input, n_inputs);
    if (lw.max_total >= n_totals) return fail(OLAP_E_INVALID, "formula reads total %d but only %d given", lw.max_total, n_totals);

    JitApi& api = jit_api();
    if (!api.ok) return fail(OLAP_E_CUDA, "computed-measure JIT unavailable: %s", api.why.c_str());

    const std::string src = eval_source(lw, n_inputs, n_totals);
    nvrtcProgram prog;
    if (api.CreateProgram(&prog, src.c_str(), "olap_eval.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(OLAP_E_CUDA, "nvrtcCreateProgram failed");
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=false"};
    const nvrtcResult rc = api.CompileProgram(prog, 4, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        api.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) api.GetProgramLog(prog, &log[0]);
        api.DestroyProgram(&prog);
        return fail(OLAP_E_CUDA, "NVRTC failed to compile formula kernel: %s", log.c_str());
    }
    size_t sz = 0;
    api.GetCUBINSize(prog, &sz);
    std::vector<char> cubin(sz);
    api.GetCUBIN(prog, cubin.data());
    api.DestroyProgram(&prog);

    JitKernel k;
    CUresult cr = api.ModuleLoadData(&k.mod, cubin.data());
    if (cr == CUDA_SUCCESS) cr = api.ModuleGetFunction(&k.fn, k.mod, "olap_eval_kernel");
    if (cr != CUDA_SUCCESS) {
        const char* msg = "?";
        api.GetErrorString(cr, &msg);
        return fail(OLAP_E_CUDA, "loading the formula kernel failed: %s", msg);
    }
    auto ins = cache.emplace(key, k);
    *out = &ins.first->second;
    return OLAP_OK;
}

}  // namespace olap
