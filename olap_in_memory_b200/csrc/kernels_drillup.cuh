// drillUp kernels — the segmented reduce of in-memory.js:265-334 on a dense layout.
//
// A cube op changes one dimension (cube.js:999-1000), so the cube is viewed as
// [O, C, I] -> [O, P, I] (outer product, changed dimension, inner product).  The
// child->parent map is turned into a CSR (parent -> ascending children) on the host;
// when the map is monotone (every time dimension) the children of a parent are a
// contiguous range and the indirection disappears.
//
// Every kernel is OUTPUT-driven: one thread owns one output cell (or 4 adjacent
// ones) and walks the children of its parent in ascending child index.  That is the
// reference's iteration order for every store built by setData/dice/reorder/drillDown
// (SURVEY.md F6), so first/last are exact, and the double accumulation happens in the
// same order as the reference's, so sum/average round to the same float32.
//
// Presence: a child takes part only if it is set (value != default), exactly like the
// reference that iterates its Map of set cells (in-memory.js:298); the count of
// contributions for `average` counts set children only (in-memory.js:320).
#pragma once
#include "common.cuh"

namespace olap {

struct UpMeasure {
    const float* in;
    float* out;
    const uint8_t* st_in;  // nullable
    uint8_t* st_out;       // nullable
    int method;
    int nan_default;
};

// ---- per-lane accumulator -------------------------------------------------
// `has` mirrors "newStore._dataMap.has(newIdx)" (in-memory.js:311): after every
// aggregate the result goes through setValue, so an accumulator that lands on the
// default is deleted and the next set child restarts it.
template <int METHOD>
struct Acc {
    double acc = 0.0;
    uint32_t cnt = 0;
    bool has = false;

    __device__ __forceinline__ void step(float v, int nan_default) {
        if (!present_f(v, nan_default)) return;
        ++cnt;
        if (!has) {
            acc = (double)v;
            has = true;
            return;
        }
        if (METHOD == OLAP_FIRST) return;
        if (METHOD == OLAP_LAST) {
            acc = (double)v;
            return;
        }
        double r;
        if (METHOD == OLAP_SUM || METHOD == OLAP_AVERAGE) r = acc + (double)v;
        else if (METHOD == OLAP_PRODUCT) r = acc * (double)v;
        else if (METHOD == OLAP_HIGHEST) r = (double)js_max((float)acc, v);
        else r = (double)js_min((float)acc, v);
        acc = r;
        // only these can land on the default: inf + -inf under a NaN default, a
        // product underflowing to 0 / hitting NaN; a plain sum hitting 0 restarts
        // with 0 + v == v, so the check is skipped there.
        if (METHOD == OLAP_PRODUCT || ((METHOD == OLAP_SUM || METHOD == OLAP_AVERAGE) && nan_default))
            has = present_d(r, nan_default);
    }

    __device__ __forceinline__ float result(int nan_default) const {
        if (!has) {
            // average of an absent accumulator: default / count (in-memory.js:323-331)
            return default_of(nan_default);
        }
        double r = acc;
        if (METHOD == OLAP_AVERAGE && cnt) r = acc / (double)cnt;
        return canon_store((float)r, nan_default);
    }
};

// ---- kernel A: one changed dimension, any I ----------------------------------
// Thread (tx, ty) of a block owns output vector j = bx*blockDim.x + tx of row
// o = by*blockDim.y + ty, where a row is the P*IV output vectors of one outer index.
// Loads along I are 128-bit and fully coalesced when VEC == 4.
struct UpMidParams {
    const UpMeasure* meas;
    const int32_t* pstart;    // [P+1]
    const int32_t* children;  // [C] ascending per parent, or nullptr when ranges are contiguous
    int64_t O;
    int32_t C, P;
    int64_t I;            // elements of the inner run handled by this launch
    int64_t in_row;       // C * I_total  (elements between consecutive o in the input)
    int64_t out_row;      // P * I_total
    int64_t I_total;      // full inner length (stride between children)
    int64_t i_base;       // first inner element of this launch (chunking of huge I)
    uint32_t IV;          // I / VEC
    FastDiv div_iv;
    uint32_t row_vecs;    // P * IV
    uint32_t blocks_per_row;
    int n_measures;
};

template <int METHOD, int VEC>
__device__ __forceinline__ void up_mid_body(const UpMidParams& p, const UpMeasure& m, int64_t o, uint32_t pi,
                                            uint32_t iv) {
    const int nan_default = m.nan_default;
    const int32_t k0 = p.pstart[pi], k1 = p.pstart[pi + 1];
    const int64_t inner = p.i_base + (int64_t)iv * VEC;
    const float* src = m.in + o * p.in_row + inner;
    const uint8_t* st_src = m.st_in ? m.st_in + o * p.in_row + inner : nullptr;
    const int64_t stride = p.I_total;

    Acc<METHOD> a[VEC];
    uint32_t st = 0;

    constexpr int U = 4;  // children in flight per thread
    int32_t k = k0;
    for (; k + U <= k1; k += U) {
        float v[U][VEC];
        uint32_t s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t c = p.children ? p.children[k + u] : (k + u);
            if (VEC == 4) {
                float4 t = ld_stream4(src + c * stride);
                v[u][0] = t.x; v[u][1 % VEC] = t.y; v[u][2 % VEC] = t.z; v[u][3 % VEC] = t.w;
                s[u] = st_src ? ld_stream_u32(st_src + c * stride) : 0u;
            } else {
                v[u][0] = ld_stream1(src + c * stride);
                s[u] = st_src ? (uint32_t)st_src[c * stride] : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) a[e].step(v[u][e], nan_default);
            st |= s[u];
        }
    }
    for (; k < k1; ++k) {
        const int64_t c = p.children ? p.children[k] : k;
        if (VEC == 4) {
            float4 t = ld_stream4(src + c * stride);
            a[0].step(t.x, nan_default); a[1 % VEC].step(t.y, nan_default);
            a[2 % VEC].step(t.z, nan_default); a[3 % VEC].step(t.w, nan_default);
            if (st_src) st |= ld_stream_u32(st_src + c * stride);
        } else {
            a[0].step(ld_stream1(src + c * stride), nan_default);
            if (st_src) st |= (uint32_t)st_src[c * stride];
        }
    }

    float* dst = m.out + o * p.out_row + (int64_t)pi * p.I_total + inner;
    if (VEC == 4) {
        float4 r;
        r.x = a[0].result(nan_default); r.y = a[1 % VEC].result(nan_default);
        r.z = a[2 % VEC].result(nan_default); r.w = a[3 % VEC].result(nan_default);
        st_stream4(dst, r);
    } else {
        *dst = a[0].result(nan_default);
    }
    if (m.st_out) {
        uint8_t* st_dst = m.st_out + o * p.out_row + (int64_t)pi * p.I_total + inner;
        if (k0 == k1) st = VEC == 4 ? 0x01010101u : 0x1u;  // no child at all: not set
        if (VEC == 4) *reinterpret_cast<uint32_t*>(st_dst) = st;
        else *st_dst = (uint8_t)st;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) drillup_mid_kernel(const __grid_constant__ UpMidParams p) {
    const uint32_t brow = blockIdx.x / p.blocks_per_row;  // uniform per block
    const uint32_t bcol = blockIdx.x - brow * p.blocks_per_row;
    const int64_t o = (int64_t)brow * blockDim.y + threadIdx.y;
    const uint32_t j = bcol * blockDim.x + threadIdx.x;
    if (o >= p.O || j >= p.row_vecs) return;
    const uint32_t pi = p.div_iv.div(j);
    const uint32_t iv = j - pi * p.IV;
    const UpMeasure m = p.meas[blockIdx.y];
    switch (m.method) {
        case OLAP_SUM: up_mid_body<OLAP_SUM, VEC>(p, m, o, pi, iv); break;
        case OLAP_AVERAGE: up_mid_body<OLAP_AVERAGE, VEC>(p, m, o, pi, iv); break;
        case OLAP_HIGHEST: up_mid_body<OLAP_HIGHEST, VEC>(p, m, o, pi, iv); break;
        case OLAP_LOWEST: up_mid_body<OLAP_LOWEST, VEC>(p, m, o, pi, iv); break;
        case OLAP_FIRST: up_mid_body<OLAP_FIRST, VEC>(p, m, o, pi, iv); break;
        case OLAP_LAST: up_mid_body<OLAP_LAST, VEC>(p, m, o, pi, iv); break;
        default: up_mid_body<OLAP_PRODUCT, VEC>(p, m, o, pi, iv); break;
    }
}

// ---- kernel G: several dimensions change at once (store API generality,
// in-memory.js:270-274).  One thread per output cell walks the cartesian product of
// its parents' children in ascending old index (odometer, last dimension fastest).
struct UpGenParams {
    const UpMeasure* meas;
    int nd;
    int64_t n_out;
    int64_t new_len[OLAP_MAX_DIMS];
    int64_t old_stride[OLAP_MAX_DIMS];
    const int32_t* pstart[OLAP_MAX_DIMS];
    const int32_t* children[OLAP_MAX_DIMS];
};

template <int METHOD>
__device__ void up_gen_body(const UpGenParams& p, const UpMeasure& m, int64_t j) {
    int32_t lo[OLAP_MAX_DIMS], hi[OLAP_MAX_DIMS], cur[OLAP_MAX_DIMS];
    int64_t rest = j;
    bool empty = false;
    for (int d = p.nd - 1; d >= 0; --d) {
        const int64_t c = rest % p.new_len[d];
        rest /= p.new_len[d];
        lo[d] = p.pstart[d][c];
        hi[d] = p.pstart[d][c + 1];
        cur[d] = lo[d];
        empty |= lo[d] == hi[d];
    }
    Acc<METHOD> a;
    uint32_t st = 0;
    if (!empty) {
        while (true) {
            int64_t off = 0;
            for (int d = 0; d < p.nd; ++d) off += (int64_t)p.children[d][cur[d]] * p.old_stride[d];
            a.step(m.in[off], m.nan_default);
            if (m.st_in) st |= m.st_in[off];
            int d = p.nd - 1;
            for (; d >= 0; --d) {
                if (++cur[d] < hi[d]) break;
                cur[d] = lo[d];
            }
            if (d < 0) break;
        }
    } else {
        st = OLAP_STATUS_UNSET;
    }
    m.out[j] = a.result(m.nan_default);
    if (m.st_out) m.st_out[j] = (uint8_t)st;
}

__global__ void __launch_bounds__(256) drillup_generic_kernel(const __grid_constant__ UpGenParams p) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.n_out) return;
    const UpMeasure m = p.meas[blockIdx.y];
    switch (m.method) {
        case OLAP_SUM: up_gen_body<OLAP_SUM>(p, m, j); break;
        case OLAP_AVERAGE: up_gen_body<OLAP_AVERAGE>(p, m, j); break;
        case OLAP_HIGHEST: up_gen_body<OLAP_HIGHEST>(p, m, j); break;
        case OLAP_LOWEST: up_gen_body<OLAP_LOWEST>(p, m, j); break;
        case OLAP_FIRST: up_gen_body<OLAP_FIRST>(p, m, j); break;
        case OLAP_LAST: up_gen_body<OLAP_LAST>(p, m, j); break;
        default: up_gen_body<OLAP_PRODUCT>(p, m, j); break;
    }
}

}  // namespace olap
