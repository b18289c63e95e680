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
    // st_in == nullptr but st_out != nullptr: the source's status plane holds nothing its values do not
    // say (status == set ? SET : UNSET in every cell, olap_store::derived): the kernel derives the
    // children's status bytes from the values it loads anyway and never reads the plane
    int derive = 0;
};

// status handling of a kernel body: no plane, load the plane, derive it from the values
enum { ST_NONE = 0, ST_LOAD = 1, ST_DERIVE = 2 };

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
        if (METHOD == OLAP_FIRST || METHOD == OLAP_COUNT) return;
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
        if (METHOD == OLAP_COUNT) r = (double)cnt;
        return canon_store((float)r, nan_default);
    }
};

// ---- branch-free lanes for the mid kernel -------------------------------------------
// Unset cells hold the canonical default, which lets most methods fold the presence test
// into the arithmetic:
//  * zero default: an unset cell is +0.0, and x + 0.0 == x, so `sum` adds every child;
//    set values are never 0, so "accumulator == 0" means "nothing yet" for
//    first/last/highest/lowest.
//  * NaN default: an unset cell is NaN and no set value is NaN, so "accumulator is NaN"
//    means "nothing yet" for first/last/highest/lowest; `sum` masks NaN children to 0.
// The only stateful corner is the reference's restart after the accumulator lands on the
// default (in-memory.js:311-318): under a NaN default that is inf + -inf, handled by
// `dead` below; under a zero default the restart is numerically a no-op.
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

template <int METHOD, bool NANDEF>
struct Lane {  // generic (product): the faithful state machine
    Acc<METHOD> a;
    __device__ __forceinline__ void step(float v) { a.step(v, NANDEF); }
    __device__ __forceinline__ float result() const { return a.result(NANDEF); }
    // fold a LATER chunk of the same parent's children into this one (split reductions)
    __device__ __forceinline__ void merge(const Lane& o) {
        if (!o.a.has) { a.cnt += o.a.cnt; return; }
        if (!a.has) { const uint32_t c = a.cnt; a = o.a; a.cnt += c; return; }
        if (METHOD == OLAP_PRODUCT) { a.acc *= o.a.acc; a.has = present_d(a.acc, NANDEF); }
        a.cnt += o.a.cnt;
    }
};

template <>
struct Lane<OLAP_SUM, false> {
    double acc = 0.0;
    __device__ __forceinline__ void step(float v) { acc += (double)v; }
    __device__ __forceinline__ float result() const { return canon_store((float)acc, 0); }
    __device__ __forceinline__ void merge(const Lane& o) { acc += o.acc; }
};
template <>
struct Lane<OLAP_AVERAGE, false> {
    double acc = 0.0;
    uint32_t cnt = 0;
    __device__ __forceinline__ void step(float v) {
        acc += (double)v;
        cnt += (v != 0.0f) ? 1u : 0u;  // NaN != 0: a stored NaN is a set cell
    }
    __device__ __forceinline__ float result() const { return canon_store(cnt ? (float)(acc / (double)cnt) : 0.0f, 0); }
    __device__ __forceinline__ void merge(const Lane& o) { acc += o.acc; cnt += o.cnt; }
};
// NaN default: unset children are NaN and are masked to -0 (the identity of IEEE addition,
// so a lone set -0 stays -0).  The reference's restart after inf + -inf (the NaN sum equals
// the default, the key is deleted, the next set child starts over: in-memory.js:311-318) is
// NOT tracked in the hot loop: a NaN accumulator with set children (`poisoned`) sends the
// output through the faithful state machine once more (exact_redo below) — it needs both
// +inf and -inf among the children of one parent.
template <>
struct Lane<OLAP_SUM, true> {
    double acc = -0.0;
    bool has = false;
    __device__ __forceinline__ void step(float v) {
        const bool pres = v == v;
        acc += (double)(pres ? v : -0.0f);
        has |= pres;
    }
    __device__ __forceinline__ bool poisoned() const { return has && acc != acc; }
    __device__ __forceinline__ float result() const { return has ? canon_store((float)acc, 1) : canon_nan(); }
    __device__ __forceinline__ void merge(const Lane& o) {
        acc += o.acc;
        has |= o.has;
    }
};
template <>
struct Lane<OLAP_AVERAGE, true> {
    double acc = -0.0;
    uint32_t cnt = 0;
    __device__ __forceinline__ void step(float v) {
        const bool pres = v == v;
        acc += (double)(pres ? v : -0.0f);
        cnt += pres ? 1u : 0u;
    }
    __device__ __forceinline__ bool poisoned() const { return cnt && acc != acc; }
    __device__ __forceinline__ float result() const { return cnt ? canon_store((float)(acc / (double)cnt), 1) : canon_nan(); }
    __device__ __forceinline__ void merge(const Lane& o) {
        acc += o.acc;
        cnt += o.cnt;
    }
};
// product: same key-exists state machine as the reference (a product that lands on the
// default — underflow to 0, inf * 0 — is deleted and the next set child restarts it),
// written with selects instead of branches
template <bool NANDEF>
struct ProductLane {
    double acc = 1.0;
    bool has = false;
    __device__ __forceinline__ void step(float v) {
        const bool pres = present_f(v, NANDEF);
        const double next = has ? acc * (double)v : (double)v;
        acc = pres ? next : acc;
        has = pres ? present_d(next, NANDEF) : has;
    }
    __device__ __forceinline__ float result() const { return has ? canon_store((float)acc, NANDEF) : default_of(NANDEF); }
    __device__ __forceinline__ void merge(const ProductLane& o) {
        if (!o.has) return;
        const double next = has ? acc * o.acc : o.acc;
        acc = next;
        has = present_d(next, NANDEF);
    }
};
template <>
struct Lane<OLAP_PRODUCT, false> : ProductLane<false> {};
template <>
struct Lane<OLAP_PRODUCT, true> : ProductLane<true> {};
template <bool NANDEF>
struct CountLane {
    uint32_t cnt = 0;
    __device__ __forceinline__ void step(float v) { cnt += present_f(v, NANDEF) ? 1u : 0u; }
    __device__ __forceinline__ float result() const { return cnt ? (float)cnt : default_of(NANDEF); }
    __device__ __forceinline__ void merge(const CountLane& o) { cnt += o.cnt; }
};
template <>
struct Lane<OLAP_COUNT, false> : CountLane<false> {};
template <>
struct Lane<OLAP_COUNT, true> : CountLane<true> {};
template <>
struct Lane<OLAP_HIGHEST, false> {
    float acc = 0.0f;
    __device__ __forceinline__ void step(float v) {
        const float m = max_nan(acc, v);  // set values are never 0 here: no signed-zero case
        acc = (v != 0.0f) ? ((acc == 0.0f) ? v : m) : acc;
    }
    __device__ __forceinline__ float result() const { return canon_store(acc, 0); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};
template <>
struct Lane<OLAP_LOWEST, false> {
    float acc = 0.0f;
    __device__ __forceinline__ void step(float v) {
        const float m = min_nan(acc, v);
        acc = (v != 0.0f) ? ((acc == 0.0f) ? v : m) : acc;
    }
    __device__ __forceinline__ float result() const { return canon_store(acc, 0); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};
template <>
struct Lane<OLAP_HIGHEST, true> {
    float acc;
    __device__ __forceinline__ Lane() : acc(canon_nan()) {}
    __device__ __forceinline__ void step(float v) {
        const float m = js_max(acc, v);  // Math.max(-0, +0) = +0
        acc = (v == v) ? ((acc != acc) ? v : m) : acc;
    }
    __device__ __forceinline__ float result() const { return canon_store(acc, 1); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};
template <>
struct Lane<OLAP_LOWEST, true> {
    float acc;
    __device__ __forceinline__ Lane() : acc(canon_nan()) {}
    __device__ __forceinline__ void step(float v) {
        const float m = js_min(acc, v);
        acc = (v == v) ? ((acc != acc) ? v : m) : acc;
    }
    __device__ __forceinline__ float result() const { return canon_store(acc, 1); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};
template <>
struct Lane<OLAP_FIRST, false> {
    float acc = 0.0f;
    __device__ __forceinline__ void step(float v) { acc = (acc == 0.0f) ? v : acc; }
    __device__ __forceinline__ float result() const { return canon_store(acc, 0); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};
template <>
struct Lane<OLAP_FIRST, true> {
    float acc;
    __device__ __forceinline__ Lane() : acc(canon_nan()) {}
    __device__ __forceinline__ void step(float v) { acc = (acc != acc) ? v : acc; }
    __device__ __forceinline__ float result() const { return canon_store(acc, 1); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};
template <>
struct Lane<OLAP_LAST, false> {
    float acc = 0.0f;
    __device__ __forceinline__ void step(float v) { acc = (v != 0.0f) ? v : acc; }
    __device__ __forceinline__ float result() const { return canon_store(acc, 0); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};
template <>
struct Lane<OLAP_LAST, true> {
    float acc;
    __device__ __forceinline__ Lane() : acc(canon_nan()) {}
    __device__ __forceinline__ void step(float v) { acc = (v == v) ? v : acc; }
    __device__ __forceinline__ float result() const { return canon_store(acc, 1); }
    __device__ __forceinline__ void merge(const Lane& o) { step(o.acc); }
};

// Lanes that can be poisoned (see Lane<OLAP_SUM, true>) report it; the others never are.
template <int METHOD, bool NANDEF>
__device__ __forceinline__ bool lane_poisoned(const Lane<METHOD, NANDEF>& l) {
    if constexpr (NANDEF && (METHOD == OLAP_SUM || METHOD == OLAP_AVERAGE)) return l.poisoned();
    else return false;
}
// Faithful recomputation of ONE output cell: children k0..k1 of its parent, ascending.
template <int METHOD, bool RANGE>
__device__ __noinline__ float exact_redo(const float* cell0, int64_t stride, const int32_t* children, int32_t k0,
                                         int32_t k1) {
    Acc<METHOD> a;
    for (int32_t k = k0; k < k1; ++k) {
        const int64_t child = RANGE ? (int64_t)k : (int64_t)children[k];
        a.step(cell0[child * stride], 1);
    }
    return a.result(1);
}

// ---- kernel A: one changed dimension, any I ----------------------------------
// Thread (tx, ty) of a block owns output vector j = bx*blockDim.x + tx of row
// o = by*blockDim.y + ty, where a row is the P*IV output vectors of one outer index.
// Loads along I are 128-bit and fully coalesced when VEC == 4; U children are in flight
// per thread.  No shared memory: there is no reuse, every input byte is read once.
constexpr int kInlineMeasures = 16;  // measure descriptors that travel in the kernel parameters

struct UpMidParams {
    const UpMeasure* meas;                      // device table, or nullptr: use meas_inline
    UpMeasure meas_inline[kInlineMeasures];
    const int32_t* pstart;    // [P+1]
    const int32_t* children;  // [C] ascending per parent (unused when RANGE)
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
    // optional (O == 1): output row pi of measure k does not go to meas[k].out but to
    // row_out[k * P + pi] (and row_st[...]) — e.g. straight into the receive buffer of the
    // rank that owns the row, through a peer-mapped pointer (olap_drill_up_rows)
    float* const* row_out;
    uint8_t* const* row_st;
};

template <int VEC>
struct Cells {
    float v[VEC];
    uint32_t st;
};

// SET / UNSET bytes of VEC cells from their values (the status of a store whose plane is `derived`)
template <int VEC, bool NANDEF>
__device__ __forceinline__ uint32_t derived_status(const float (&v)[VEC]) {
    uint32_t st = 0;
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        st |= (present_f(v[e], NANDEF) ? (uint32_t)OLAP_STATUS_SET : (uint32_t)OLAP_STATUS_UNSET) << (8 * e);
    return st;
}

template <int VEC, int STATUS, bool NANDEF = false>
__device__ __forceinline__ Cells<VEC> load_cells(const float* src, const uint8_t* st_src, int64_t off) {
    Cells<VEC> c;
    if (VEC == 4) {
        const float4 t = ld_stream4(src + off);
        c.v[0] = t.x; c.v[1 % VEC] = t.y; c.v[2 % VEC] = t.z; c.v[3 % VEC] = t.w;
        c.st = STATUS == ST_LOAD ? ld_stream_u32(st_src + off) : 0u;
    } else if (VEC == 2) {
        const float2 t = ld_stream2(src + off);
        c.v[0] = t.x; c.v[1 % VEC] = t.y;
        c.st = STATUS == ST_LOAD ? ld_stream_u16(st_src + off) : 0u;
    } else {
        c.v[0] = ld_stream1(src + off);
        c.st = STATUS == ST_LOAD ? (uint32_t)st_src[off] : 0u;
    }
    return c;  // ST_DERIVE: the bytes are computed where the cells are folded (cells_status), never next to the load
}
// status bytes of loaded cells at the point of use: keeps every load of a batch ahead of the first dependent instruction
template <int VEC, int STATUS, bool NANDEF>
__device__ __forceinline__ uint32_t cells_status(const Cells<VEC>& c) {
    if (STATUS == ST_DERIVE) return derived_status<VEC, NANDEF>(c.v);
    return c.st;
}

// VEC results (+ VEC status bytes packed in `st`) to consecutive cells
template <int VEC>
__device__ __forceinline__ void store_cells(float* dst, const float (&r)[VEC]) {
    if (VEC == 4) st_stream4(dst, make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]));
    else if (VEC == 2) *reinterpret_cast<float2*>(dst) = make_float2(r[0], r[1 % VEC]);
    else *dst = r[0];
}
template <int VEC>
__device__ __forceinline__ void store_status(uint8_t* dst, uint32_t st) {
    if (VEC == 4) *reinterpret_cast<uint32_t*>(dst) = st;
    else if (VEC == 2) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)st;
    else *dst = (uint8_t)st;
}
template <int VEC>
__device__ __forceinline__ uint32_t unset_status() {  // OLAP_STATUS_UNSET in each of the VEC bytes
    return VEC == 4 ? 0x01010101u : (VEC == 2 ? 0x0101u : 0x01u);
}

template <int METHOD, bool NANDEF, int VEC, bool RANGE, int STATUS, int U>
__device__ __forceinline__ void up_mid_body(const UpMidParams& p, const UpMeasure& m, int64_t o, uint32_t pi,
                                            uint32_t iv) {
    const int32_t k0 = p.pstart[pi], k1 = p.pstart[pi + 1];
    const int64_t inner = p.i_base + (int64_t)iv * VEC;
    const float* src = m.in + o * p.in_row + inner;
    const uint8_t* st_src = STATUS == ST_LOAD ? m.st_in + o * p.in_row + inner : nullptr;
    const int64_t stride = p.I_total;

    Lane<METHOD, NANDEF> lane[VEC];
    uint32_t st = 0;

    // U children in flight per thread
    int32_t k = k0;
    for (; k + U <= k1; k += U) {
        Cells<VEC> c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t child = RANGE ? (int64_t)(k + u) : (int64_t)p.children[k + u];
            c[u] = load_cells<VEC, STATUS, NANDEF>(src, st_src, child * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) lane[e].step(c[u].v[e]);
            st |= cells_status<VEC, STATUS, NANDEF>(c[u]);
        }
    }
    for (; k < k1; ++k) {
        const int64_t child = RANGE ? (int64_t)k : (int64_t)p.children[k];
        const Cells<VEC> c = load_cells<VEC, STATUS, NANDEF>(src, st_src, child * stride);
#pragma unroll
        for (int e = 0; e < VEC; ++e) lane[e].step(c.v[e]);
        st |= cells_status<VEC, STATUS, NANDEF>(c);
    }

    const int64_t out_off = o * p.out_row + (int64_t)pi * p.I_total + inner;
    float r[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        r[e] = lane_poisoned(lane[e]) ? exact_redo<METHOD, RANGE>(src + e, stride, p.children, k0, k1) : lane[e].result();
    store_cells<VEC>(m.out + out_off, r);
    if (STATUS) store_status<VEC>(m.st_out + out_off, k0 == k1 ? unset_status<VEC>() : st);  // no child: not set
}

template <bool NANDEF, int VEC, bool RANGE, int STATUS, int U>
__device__ __forceinline__ void up_mid_dispatch(const UpMidParams& p, const UpMeasure& m, int64_t o, uint32_t pi,
                                                uint32_t iv) {
    switch (m.method) {
        case OLAP_SUM: up_mid_body<OLAP_SUM, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
        case OLAP_AVERAGE: up_mid_body<OLAP_AVERAGE, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
        case OLAP_HIGHEST: up_mid_body<OLAP_HIGHEST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
        case OLAP_LOWEST: up_mid_body<OLAP_LOWEST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
        case OLAP_FIRST: up_mid_body<OLAP_FIRST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
        case OLAP_LAST: up_mid_body<OLAP_LAST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
        case OLAP_COUNT: up_mid_body<OLAP_COUNT, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
        default: up_mid_body<OLAP_PRODUCT, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv); break;
    }
}

// The second launch bound matters: without it ptxas aims for <= 64 registers, cannot keep the U child
// vectors of a batch live, and interleaves the loads with the double adds of the sum / average lanes
// (SASS: LDG x4, DADD x3, LDG, DADD x6, ...): 4-5 loads in flight instead of 2U, and builds that differ
// in unrelated code land on different allocations (48 / 58 / 60 registers: 0.52 - 0.85 of peak on
// day -> year).  With "3 CTAs per SM" (<= 80 registers) every batch is LDG x 2U, then the adds.
#ifndef OLAP_MID_MINB
#define OLAP_MID_MINB 3
#endif
template <int VEC, bool RANGE, int U>
__global__ void __launch_bounds__(256, U == 8 ? (VEC == 4 ? OLAP_MID_MINB : 4) : 4) drillup_mid_kernel(const __grid_constant__ UpMidParams p) {
    const uint32_t brow = blockIdx.x / p.blocks_per_row;  // uniform per block
    const uint32_t bcol = blockIdx.x - brow * p.blocks_per_row;
    const int64_t o = (int64_t)brow * blockDim.y + threadIdx.y;
    const uint32_t j = bcol * blockDim.x + threadIdx.x;
    if (o >= p.O || j >= p.row_vecs) return;
    const uint32_t pi = p.div_iv.div(j);
    const uint32_t iv = j - pi * p.IV;
    UpMeasure m = p.meas ? p.meas[blockIdx.y] : p.meas_inline[blockIdx.y];
    if (p.row_out) {  // rebase so that  out + pi * I_total + inner  lands in the row's own buffer
        const size_t slot = (size_t)blockIdx.y * p.P + pi;
        m.out = p.row_out[slot] - (int64_t)pi * p.I_total;
        if (m.st_in || m.derive) m.st_out = p.row_st[slot] - (int64_t)pi * p.I_total;
    }
    const int status = m.st_in ? ST_LOAD : (m.derive ? ST_DERIVE : ST_NONE);
    if (m.nan_default) {
        if (status == ST_LOAD) up_mid_dispatch<true, VEC, RANGE, ST_LOAD, U>(p, m, o, pi, iv);
        else if (status == ST_DERIVE) up_mid_dispatch<true, VEC, RANGE, ST_DERIVE, U>(p, m, o, pi, iv);
        else up_mid_dispatch<true, VEC, RANGE, ST_NONE, U>(p, m, o, pi, iv);
    } else {
        if (status == ST_LOAD) up_mid_dispatch<false, VEC, RANGE, ST_LOAD, U>(p, m, o, pi, iv);
        else if (status == ST_DERIVE) up_mid_dispatch<false, VEC, RANGE, ST_DERIVE, U>(p, m, o, pi, iv);
        else up_mid_dispatch<false, VEC, RANGE, ST_NONE, U>(p, m, o, pi, iv);
    }
}

// ---- kernel A/split: few outputs, long child lists (drillUp to 'all', year, ...) ------
// When O*P*I/4 threads cannot fill the chip, G thread rows share one output vector: row g
// reduces the g-th contiguous chunk of the parent's children (same coalesced 128-bit loads
// as kernel A), lane states meet in shared memory and row 0 folds them IN CHUNK ORDER, so
// first/last stay exact and the double sums only change their association.
template <int METHOD, bool NANDEF, int VEC, bool RANGE, int STATUS, int U>
__device__ __forceinline__ void up_split_body(const UpMidParams& p, const UpMeasure& m, int64_t o, uint32_t pi,
                                              uint32_t iv, bool live, unsigned char* smem_raw) {
    typedef Lane<METHOD, NANDEF> L;
    const int G = blockDim.y, g = threadIdx.y, tx = threadIdx.x;
    L* s_lane = reinterpret_cast<L*>(smem_raw);                                   // [G][32][VEC]
    uint32_t* s_st = reinterpret_cast<uint32_t*>(smem_raw + (size_t)G * 32 * VEC * sizeof(L));  // [G][32]
    L lane[VEC];
    uint32_t st = 0;
    int32_t k0 = 0, k1 = 0;
    const int64_t inner = p.i_base + (int64_t)iv * VEC;
    if (live) {
        k0 = p.pstart[pi];
        k1 = p.pstart[pi + 1];
        const int32_t per = (k1 - k0 + G - 1) / G;
        const int32_t ks = min(k1, k0 + g * per), ke = min(k1, ks + per);
        const float* src = m.in + o * p.in_row + inner;
        const uint8_t* st_src = STATUS == ST_LOAD ? m.st_in + o * p.in_row + inner : nullptr;
        int32_t k = ks;
        for (; k + U <= ke; k += U) {
            Cells<VEC> c[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t child = RANGE ? (int64_t)(k + u) : (int64_t)p.children[k + u];
                c[u] = load_cells<VEC, STATUS, NANDEF>(src, st_src, child * p.I_total);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) lane[e].step(c[u].v[e]);
                st |= cells_status<VEC, STATUS, NANDEF>(c[u]);
            }
        }
        for (; k < ke; ++k) {
            const int64_t child = RANGE ? (int64_t)k : (int64_t)p.children[k];
            const Cells<VEC> c = load_cells<VEC, STATUS, NANDEF>(src, st_src, child * p.I_total);
#pragma unroll
            for (int e = 0; e < VEC; ++e) lane[e].step(c.v[e]);
            st |= cells_status<VEC, STATUS, NANDEF>(c);
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) s_lane[((size_t)g * 32 + tx) * VEC + e] = lane[e];
    s_st[g * 32 + tx] = st;
    __syncthreads();
    if (g != 0 || !live) return;
    for (int q = 1; q < G; ++q) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) lane[e].merge(s_lane[((size_t)q * 32 + tx) * VEC + e]);
        st |= s_st[q * 32 + tx];
    }
    const int64_t out_off = o * p.out_row + (int64_t)pi * p.I_total + inner;
    float r[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        r[e] = lane_poisoned(lane[e])
                   ? exact_redo<METHOD, RANGE>(m.in + o * p.in_row + inner + e, p.I_total, p.children, k0, k1)
                   : lane[e].result();
    store_cells<VEC>(m.out + out_off, r);
    if (STATUS) store_status<VEC>(m.st_out + out_off, k0 == k1 ? unset_status<VEC>() : st);
}

template <bool NANDEF, int VEC, bool RANGE, int STATUS, int U>
__device__ __forceinline__ void up_split_dispatch(const UpMidParams& p, const UpMeasure& m, int64_t o, uint32_t pi,
                                                  uint32_t iv, bool live, unsigned char* smem_raw) {
    switch (m.method) {
        case OLAP_SUM: up_split_body<OLAP_SUM, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
        case OLAP_AVERAGE: up_split_body<OLAP_AVERAGE, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
        case OLAP_HIGHEST: up_split_body<OLAP_HIGHEST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
        case OLAP_LOWEST: up_split_body<OLAP_LOWEST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
        case OLAP_FIRST: up_split_body<OLAP_FIRST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
        case OLAP_LAST: up_split_body<OLAP_LAST, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
        case OLAP_COUNT: up_split_body<OLAP_COUNT, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
        default: up_split_body<OLAP_PRODUCT, NANDEF, VEC, RANGE, STATUS, U>(p, m, o, pi, iv, live, smem_raw); break;
    }
}

// blockDim = (32, G); one outer row per block row: blockIdx.x = o * blocks_per_row + column block.
// WIDE (G <= 8, 256 threads): 8 children in flight per thread under the register budget of the mid
// kernel (3 CTAs per SM); otherwise G <= 32 thread rows with 4 children in flight each.
template <int VEC, bool RANGE, bool WIDE>
__global__ void __launch_bounds__(WIDE ? 256 : 1024, WIDE ? 3 : 1) drillup_split_kernel(const __grid_constant__ UpMidParams p) {
    constexpr int U = WIDE ? 8 : 4;
    extern __shared__ __align__(16) unsigned char smem_split[];
    const uint32_t brow = blockIdx.x / p.blocks_per_row;
    const uint32_t bcol = blockIdx.x - brow * p.blocks_per_row;
    const int64_t o = brow;
    const uint32_t j = bcol * 32 + threadIdx.x;
    const bool live = j < p.row_vecs;
    const uint32_t pi = live ? p.div_iv.div(j) : 0u;
    const uint32_t iv = live ? j - pi * p.IV : 0u;
    const UpMeasure m = p.meas ? p.meas[blockIdx.y] : p.meas_inline[blockIdx.y];
    const int status = m.st_in ? ST_LOAD : (m.derive ? ST_DERIVE : ST_NONE);
    if (m.nan_default) {
        if (status == ST_LOAD) up_split_dispatch<true, VEC, RANGE, ST_LOAD, U>(p, m, o, pi, iv, live, smem_split);
        else if (status == ST_DERIVE) up_split_dispatch<true, VEC, RANGE, ST_DERIVE, U>(p, m, o, pi, iv, live, smem_split);
        else up_split_dispatch<true, VEC, RANGE, ST_NONE, U>(p, m, o, pi, iv, live, smem_split);
    } else {
        if (status == ST_LOAD) up_split_dispatch<false, VEC, RANGE, ST_LOAD, U>(p, m, o, pi, iv, live, smem_split);
        else if (status == ST_DERIVE) up_split_dispatch<false, VEC, RANGE, ST_DERIVE, U>(p, m, o, pi, iv, live, smem_split);
        else up_split_dispatch<false, VEC, RANGE, ST_NONE, U>(p, m, o, pi, iv, live, smem_split);
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

static __global__ void __launch_bounds__(256) drillup_generic_kernel(const __grid_constant__ UpGenParams p) {
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
        case OLAP_COUNT: up_gen_body<OLAP_COUNT>(p, m, j); break;
        default: up_gen_body<OLAP_PRODUCT>(p, m, j); break;
    }
}

}  // namespace olap

// =====================================================================================
// drillDown with ONE changed dimension: [O, P, I] -> [O, C, I] (in-memory.js:336-430).
// Mirror image of the mid kernel: one thread owns one PARENT vector, reads it once, and
// writes it (scaled) to each of the parent's children, so the parent plane is read once
// from HBM and every store is a coalesced 128-bit write along I.  `pstart/children` is
// the CSR parent -> ascending new items; the rank of a child among its siblings
// (contributionsIds, in-memory.js:406-426) is its position in that list.
namespace olap {

struct DownMeasure {
    const float* in;
    float* out;
    const uint8_t* st_in;
    uint8_t* st_out;
    int nan_default;
    int kind;  // 0: float sum (v / n)   1: copy (method != sum)   2: integer sum with spreading
};

struct DownMidParams {
    const DownMeasure* meas;
    const int32_t* pstart;    // [P+1]
    const int32_t* children;  // [C] new items per parent, ascending
    int64_t O;
    int32_t C, P;
    int64_t I_total, in_row, out_row, i_base;
    uint32_t IV;
    FastDiv div_iv;
    uint32_t row_vecs;  // P * IV
    uint32_t blocks_per_row;
};

// KIND is the measure's kind (0: float sum, 1: copy, 2: integer spreading), a template parameter so
// that the float / copy loops are a bare address step + two stores per child and stay unrolled.
template <int VEC, bool RANGE, bool STATUS, int KIND>
__device__ __forceinline__ void down_mid_body(const DownMidParams& p, const DownMeasure& m, int64_t o, uint32_t pi,
                                              uint32_t iv) {
    const int32_t k0 = p.pstart[pi], k1 = p.pstart[pi + 1];
    if (k0 == k1) return;
    const int64_t inner = p.i_base + (int64_t)iv * VEC;
    const int64_t in_off = o * p.in_row + (int64_t)pi * p.I_total + inner;
    const Cells<VEC> c = load_cells<VEC, STATUS>(m.in, m.st_in, in_off);
    const int nan_default = m.nan_default;
    const double dn = (double)(uint32_t)(k1 - k0);
    float r[VEC];
    bool truthy[VEC];
    double base[VEC], step[VEC], prev[VEC];
    uint32_t st_ok = 0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const float x = c.v[e];
        truthy[e] = x != 0.0f && x == x;  // `if (!oldValue) continue` (in-memory.js:386-387)
        if (KIND == 0) r[e] = canon_store((float)((double)x / dn), nan_default);
        else if (KIND == 1) r[e] = x;
        else {
            const double q = (double)x / dn;
            base[e] = floor(q);
            step[e] = fma(-trunc(q), dn, (double)x) / dn;  // (v % n) / n
            prev[e] = floor(-step[e]);                     // floor((k - 1) * step) at k = 0
            r[e] = 0.0f;
        }
        if (!truthy[e]) r[e] = default_of(nan_default);
        const uint32_t sb = STATUS ? ((c.st >> (8 * e)) & 0xffu) : 0u;
        const bool ok = truthy[e] && (KIND == 2 || present_f(r[e], nan_default));
        st_ok |= (ok ? ((sb | OLAP_STATUS_INTERPOLATED) & 0xffu) : (uint32_t)OLAP_STATUS_UNSET) << (8 * e);
    }
    float* dst = m.out + o * p.out_row + inner;
    uint8_t* st_dst = STATUS ? m.st_out + o * p.out_row + inner : nullptr;
    if (KIND != 2) {
        // every child receives the same vector: nothing but stores
        int32_t k = k0;
#pragma unroll 4
        for (; k < k1; ++k) {
            const int64_t off = (RANGE ? (int64_t)k : (int64_t)p.children[k]) * p.I_total;
            store_cells<VEC>(dst + off, r);
            if (STATUS) store_status<VEC>(st_dst + off, st_ok);
        }
    } else
    for (int32_t k = k0; k < k1; ++k) {
        const int64_t off = (RANGE ? (int64_t)k : (int64_t)p.children[k]) * p.I_total;
        const double kk = (double)(k - k0);
        uint32_t st = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            // floor(k * step) !== floor((k - 1) * step) (in-memory.js:409-416); the previous floor is carried along
            const double cur = floor(kk * step[e]);
            const bool last_is_same = cur == prev[e];
            prev[e] = cur;
            const float val = canon_store((float)(last_is_same ? base[e] : base[e] + 1.0), nan_default);
            r[e] = truthy[e] ? val : default_of(nan_default);
            const uint32_t sb = STATUS ? ((c.st >> (8 * e)) & 0xffu) : 0u;
            const bool ok = truthy[e] && present_f(val, nan_default);
            st |= (ok ? ((sb | OLAP_STATUS_INTERPOLATED) & 0xffu) : (uint32_t)OLAP_STATUS_UNSET) << (8 * e);
        }
        store_cells<VEC>(dst + off, r);
        if (STATUS) store_status<VEC>(st_dst + off, st);
    }
}

// KIND >= 0: every measure of the call has that kind (the common case: its loop alone decides the
// register count — 8 resident CTAs for the float / copy kernels); KIND < 0: kinds differ, dispatch per measure.
template <int VEC, bool RANGE, int KIND>
__global__ void __launch_bounds__(256, KIND == 0 || KIND == 1 ? 6 : 4) drilldown_mid_kernel(const __grid_constant__ DownMidParams p) {
    const uint32_t brow = blockIdx.x / p.blocks_per_row;
    const uint32_t bcol = blockIdx.x - brow * p.blocks_per_row;
    const int64_t o = (int64_t)brow * blockDim.y + threadIdx.y;
    const uint32_t j = bcol * blockDim.x + threadIdx.x;
    if (o >= p.O || j >= p.row_vecs) return;
    const uint32_t pi = p.div_iv.div(j);
    const uint32_t iv = j - pi * p.IV;
    const DownMeasure m = p.meas[blockIdx.y];
    if (KIND >= 0) {
        constexpr int K = KIND >= 0 ? KIND : 0;
        if (m.st_in) down_mid_body<VEC, RANGE, true, K>(p, m, o, pi, iv);
        else down_mid_body<VEC, RANGE, false, K>(p, m, o, pi, iv);
    } else if (m.st_in) {
        if (m.kind == 0) down_mid_body<VEC, RANGE, true, 0>(p, m, o, pi, iv);
        else if (m.kind == 1) down_mid_body<VEC, RANGE, true, 1>(p, m, o, pi, iv);
        else down_mid_body<VEC, RANGE, true, 2>(p, m, o, pi, iv);
    } else {
        if (m.kind == 0) down_mid_body<VEC, RANGE, false, 0>(p, m, o, pi, iv);
        else if (m.kind == 1) down_mid_body<VEC, RANGE, false, 1>(p, m, o, pi, iv);
        else down_mid_body<VEC, RANGE, false, 2>(p, m, o, pi, iv);
    }
}

}  // namespace olap

// =====================================================================================
// drillDown of the INNERMOST axis: [O, P] -> [O, C] (time innermost, I == 1).
// A CTA owns RB consecutive rows: their RB*P parents and the per-child tables (parent,
// rank) sit in shared memory; every thread produces 4 consecutive children and writes them
// with one 128-bit store (the output span of a CTA is contiguous).
namespace olap {

struct DownInnerParams {
    const DownMeasure* meas;
    const int32_t* parent_of;  // [C] new item -> parent
    const int32_t* rank_of;    // [C] rank among siblings
    const int32_t* cnt_of;     // [P] siblings per parent
    int64_t O;
    int32_t P, C;
    uint32_t I;                // cells of the untouched inner run (1: the drilled axis is the innermost one)
    uint32_t RB;               // rows (outer indices) per CTA
    FastDiv div_i, div_pi, div_ci;  // by I, P * I, C * I
    int vec4;                  // I % 4 == 0 (or I == 1 and C % 4 == 0): 128-bit stores
};

// A CTA owns RB consecutive outer indices: their RB * P * I parent cells are staged once, and
// everything that depends only on the PARENT cell (value / n, integer base and step, status) is
// computed while staging; the CTA then emits its RB * C * I output cells in memory order, every
// thread 4 consecutive cells with one 128-bit store.  The output span of a CTA is contiguous and
// written front to back: full sectors, one sequential write stream per CTA — which is what a short
// inner run (I = 100: 400-byte runs per child with the parent-driven kernel) needs.
template <bool STATUS, bool I1>
__device__ __forceinline__ void down_inner_body(const DownInnerParams& p, const DownMeasure& m, unsigned char* smem) {
    const uint32_t PI = I1 ? (uint32_t)p.P : (uint32_t)p.P * p.I, CI = I1 ? (uint32_t)p.C : (uint32_t)p.C * p.I;
    const uint32_t n_par = p.RB * PI;
    double* s_step = reinterpret_cast<double*>(smem);                 // [RB * P * I]  integer spreading only
    float* s_res = reinterpret_cast<float*>(s_step + n_par);          // [RB * P * I]  child value (or floor(v/n))
    int32_t* s_parent = reinterpret_cast<int32_t*>(s_res + n_par);    // [C]
    int32_t* s_rank = s_parent + p.C;                                 // [C]
    uint8_t* s_st = reinterpret_cast<uint8_t*>(s_rank + p.C);         // [RB * P * I]  child status
    uint8_t* s_truthy = s_st + n_par;                                 // [RB * P * I]
    const int64_t row0 = (int64_t)blockIdx.x * p.RB;
    const uint32_t rows = (uint32_t)min((int64_t)p.RB, p.O - row0);
    const int nan_default = m.nan_default;
    for (uint32_t i = threadIdx.x; i < rows * PI; i += blockDim.x) {
        const float x = m.in[row0 * PI + i];
        const uint32_t sb = STATUS ? m.st_in[row0 * PI + i] : 0u;
        const uint32_t in_row = i - p.div_pi.div(i) * PI;
        const uint32_t par = I1 ? in_row : p.div_i.div(in_row);
        const double dn = (double)p.cnt_of[par];
        const bool truthy = x != 0.0f && x == x;  // `if (!oldValue) continue` (in-memory.js:386-387)
        float r;
        bool ok = truthy;
        if (m.kind == 0) { r = canon_store((float)((double)x / dn), nan_default); ok = truthy && present_f(r, nan_default); }
        else if (m.kind == 1) r = x;
        else {
            const double q = (double)x / dn;
            r = (float)floor(q);
            s_step[i] = fma(-trunc(q), dn, (double)x) / dn;  // (v % n) / n
        }
        s_res[i] = truthy ? r : default_of(nan_default);
        s_truthy[i] = truthy;
        s_st[i] = (uint8_t)(ok ? ((sb | OLAP_STATUS_INTERPOLATED) & 0xffu) : (uint32_t)OLAP_STATUS_UNSET);
    }
    for (uint32_t c = threadIdx.x; c < (uint32_t)p.C; c += blockDim.x) {
        s_parent[c] = p.parent_of[c];
        s_rank[c] = p.rank_of[c];
    }
    __syncthreads();
    const uint32_t cells = rows * CI;
    float* out = m.out + row0 * CI;
    uint8_t* st_out = STATUS ? m.st_out + row0 * CI : nullptr;
    // cell `src` of the staged parents as seen by child `c` (rank among its siblings: s_rank[c])
    auto cell = [&](uint32_t src, uint32_t c, uint32_t& so) {
        float v = s_res[src];
        so = s_st[src];
        if (m.kind == 2 && s_truthy[src]) {
            const double k = (double)s_rank[c], step = s_step[src];
            const bool last_is_same = floor(k * step) == floor((k - 1.0) * step);
            v = canon_store(last_is_same ? v : v + 1.0f, nan_default);
            if (STATUS && !present_f(v, nan_default)) so = OLAP_STATUS_UNSET;
        }
        return v;
    };
    if (p.vec4) {
        for (uint32_t t = threadIdx.x * 4; t < cells; t += blockDim.x * 4) {
            const uint32_t r = p.div_ci.div(t);
            const uint32_t w = t - r * CI;
            float v[4];
            uint32_t st = 0;
            if (I1) {  // C % 4 == 0: four consecutive children of one row
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint32_t so;
                    v[e] = cell(r * PI + (uint32_t)s_parent[w + e], w + e, so);
                    st |= so << (8 * e);
                }
            } else {         // I % 4 == 0: four consecutive inner cells of one child
                const uint32_t c = p.div_i.div(w), i = w - c * p.I;
                const uint32_t src = r * PI + (uint32_t)s_parent[c] * p.I + i;
                if (m.kind != 2) {
                    const float4 q = *reinterpret_cast<const float4*>(s_res + src);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                    st = *reinterpret_cast<const uint32_t*>(s_st + src);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint32_t so;
                        v[e] = cell(src + e, c, so);
                        st |= so << (8 * e);
                    }
                }
            }
            st_stream4(out + t, make_float4(v[0], v[1], v[2], v[3]));
            if (STATUS) *reinterpret_cast<uint32_t*>(st_out + t) = st;
        }
    } else {
        for (uint32_t t = threadIdx.x; t < cells; t += blockDim.x) {
            const uint32_t r = p.div_ci.div(t);
            const uint32_t w = t - r * CI;
            const uint32_t c = I1 ? w : p.div_i.div(w), i = I1 ? 0u : w - c * p.I;
            uint32_t so;
            out[t] = cell(r * PI + (I1 ? (uint32_t)s_parent[c] : (uint32_t)s_parent[c] * p.I + i), c, so);
            if (STATUS) st_out[t] = (uint8_t)so;
        }
    }
}

static __global__ void __launch_bounds__(256, 4) drilldown_inner_kernel(const __grid_constant__ DownInnerParams p) {
    extern __shared__ __align__(16) unsigned char smem_di[];
    const DownMeasure m = p.meas[blockIdx.y];
    if (p.I == 1) {
        if (m.st_in) down_inner_body<true, true>(p, m, smem_di);
        else down_inner_body<false, true>(p, m, smem_di);
    } else {
        if (m.st_in) down_inner_body<true, false>(p, m, smem_di);
        else down_inner_body<false, false>(p, m, smem_di);
    }
}

}  // namespace olap
