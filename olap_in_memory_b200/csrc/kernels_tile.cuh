// Shared-memory staged kernels.
//
// drillup_tile_kernel — drillUp along the innermost axis or with a short inner run
// (I < 32): the output-driven mid kernel would read 4..124-byte fragments.  Here a CTA
// owns R consecutive outer rows; their R*C*I input cells are ONE contiguous span, staged
// into shared memory with a bulk async copy (cp.async.bulk / TMA 1-D, completion on an
// mbarrier), then every thread reduces whole parents out of shared memory and the
// R*P*I results leave as one contiguous, coalesced span.  Several CTAs are resident per
// SM, so one CTA's copy overlaps its neighbours' reduction without an explicit pipeline.
#pragma once
#include "common.cuh"
#include "kernels_drillup.cuh"
#include "kernels_gather.cuh"

namespace olap {

// ---- mbarrier / bulk-copy primitives (PTX) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; src, dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct UpTileParams {
    const UpMeasure* meas;
    const int32_t* pstart;
    const int32_t* children;
    int64_t O;
    int32_t C, P, I;
    int32_t R;              // rows per tile
    int32_t row_in, row_out;  // C*I, P*I
    FastDiv div_row_out, div_i;
    int32_t bulk_values;    // tile starts are 16-byte aligned for the float plane
    int32_t bulk_status;    // ... and for the status plane
    uint32_t st_offset;     // byte offset of the status tile in dynamic shared memory
};

struct TileDecision {
    bool use = false;
    int R = 1;
    bool bulk_values = false, bulk_status = false;
    size_t smem = 0;
    uint32_t st_offset = 0;
};

// Use the tile kernel when the inner run is short and at least one row fits in shared memory.
inline TileDecision tile_plan(int64_t O, int64_t C, int64_t P, int64_t I, bool any_status) {
    TileDecision t;
    static const int force = [] { const char* e = getenv("OLAP_TILE"); return e ? atoi(e) : -1; }();  // tuning knob
    if (force == 0) return t;
    if (I >= 32 && (I % 4 == 0 || I >= 128)) return t;  // the vectorised mid kernel streams these
    const int64_t row_in = C * I;
    const int64_t per_cell = any_status ? 5 : 4;
    const int64_t budget = 48 * 1024, hard = 200 * 1024;
    if (row_in * per_cell > hard || row_in > 0x3fffffff || P * I > 0x3fffffff) return t;
    int64_t R = std::max<int64_t>(1, budget / (row_in * per_cell));
    R = std::min<int64_t>(R, std::max<int64_t>(O, 1));
    // alignment of every tile start: R*row_in % 4 == 0 (floats), % 16 == 0 (status bytes)
    auto round_to = [&](int64_t mult) {
        int64_t r = (R / mult) * mult;
        if (r == 0) r = mult;
        return r;
    };
    const int64_t mult_v = 4 / std::__gcd<int64_t>(row_in % 4 == 0 ? 4 : row_in % 4, 4);
    const int64_t mult_s = 16 / std::__gcd<int64_t>(row_in % 16 == 0 ? 16 : row_in % 16, 16);
    int64_t Rs = round_to(any_status ? mult_s : mult_v);
    if (Rs * row_in * per_cell <= hard) {
        R = Rs;
        t.bulk_values = true;
        t.bulk_status = any_status;
    } else {
        int64_t Rv = round_to(mult_v);
        if (Rv * row_in * per_cell <= hard) { R = Rv; t.bulk_values = true; }
    }
    t.R = (int)R;
    const size_t vbytes = ((size_t)R * row_in * 4 + 15) & ~(size_t)15;
    t.st_offset = (uint32_t)vbytes;
    t.smem = vbytes + (any_status ? (((size_t)R * row_in + 15) & ~(size_t)15) : 0) + 16;
    t.use = true;
    return t;
}

template <int METHOD, bool NANDEF, bool RANGE, bool STATUS>
__device__ __forceinline__ void up_tile_reduce(const UpTileParams& p, const UpMeasure& m, const float* s_val,
                                               const uint8_t* s_st, int64_t o0, int rows) {
    const int n_out = rows * p.row_out;
    float* out = m.out + o0 * p.row_out;
    uint8_t* st_out = STATUS ? m.st_out + o0 * p.row_out : nullptr;
    for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
        const uint32_t r = p.div_row_out.div((uint32_t)j);
        const uint32_t q = (uint32_t)j - r * (uint32_t)p.row_out;
        const uint32_t pi = p.div_i.div(q);
        const uint32_t i = q - pi * (uint32_t)p.I;
        const int32_t k0 = p.pstart[pi], k1 = p.pstart[pi + 1];
        const uint32_t base = r * (uint32_t)p.row_in + i;
        Lane<METHOD, NANDEF> lane;
        uint32_t st = 0;
#pragma unroll 4
        for (int32_t k = k0; k < k1; ++k) {
            const uint32_t c = RANGE ? (uint32_t)k : (uint32_t)p.children[k];
            const uint32_t idx = base + c * (uint32_t)p.I;
            lane.step(s_val[idx]);
            if (STATUS) st |= s_st[idx];
        }
        out[j] = lane.result();
        if (STATUS) st_out[j] = (uint8_t)(k0 == k1 ? OLAP_STATUS_UNSET : st);
    }
}

template <bool NANDEF, bool RANGE, bool STATUS>
__device__ __forceinline__ void up_tile_dispatch(const UpTileParams& p, const UpMeasure& m, const float* s_val,
                                                 const uint8_t* s_st, int64_t o0, int rows) {
    switch (m.method) {
        case OLAP_SUM: up_tile_reduce<OLAP_SUM, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows); break;
        case OLAP_AVERAGE: up_tile_reduce<OLAP_AVERAGE, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows); break;
        case OLAP_HIGHEST: up_tile_reduce<OLAP_HIGHEST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows); break;
        case OLAP_LOWEST: up_tile_reduce<OLAP_LOWEST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows); break;
        case OLAP_FIRST: up_tile_reduce<OLAP_FIRST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows); break;
        case OLAP_LAST: up_tile_reduce<OLAP_LAST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows); break;
        default: up_tile_reduce<OLAP_PRODUCT, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows); break;
    }
}

template <bool RANGE>
__global__ void __launch_bounds__(256) drillup_tile_kernel(const __grid_constant__ UpTileParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* s_val = reinterpret_cast<float*>(smem);
    uint8_t* s_st = smem + p.st_offset;
    const UpMeasure m = p.meas[blockIdx.y];
    const bool status = m.st_in != nullptr;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.st_offset + (status ? (((size_t)p.R * p.row_in + 15) & ~(size_t)15) : 0));

    const int64_t o0 = (int64_t)blockIdx.x * p.R;
    const int rows = (int)min((int64_t)p.R, p.O - o0);
    const int64_t n_cells = (int64_t)rows * p.row_in;
    const float* g_val = m.in + o0 * p.row_in;
    const uint8_t* g_st = status ? m.st_in + o0 * p.row_in : nullptr;

    // --- stage the tile: bulk async copy for the 16-byte multiple, plain loads for the rest
    uint32_t bulk_v = p.bulk_values ? (uint32_t)((n_cells * 4) & ~(int64_t)15) : 0u;
    uint32_t bulk_s = (status && p.bulk_status) ? (uint32_t)(n_cells & ~(int64_t)15) : 0u;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bulk_v + bulk_s);
        if (bulk_v) bulk_g2s(s_val, g_val, bulk_v, bar);
        if (bulk_s) bulk_g2s(s_st, g_st, bulk_s, bar);
    }
    for (int64_t i = (bulk_v >> 2) + threadIdx.x; i < n_cells; i += blockDim.x) s_val[i] = ld_stream1(g_val + i);
    if (status)
        for (int64_t i = bulk_s + threadIdx.x; i < n_cells; i += blockDim.x) s_st[i] = g_st[i];
    mbar_wait(bar, 0);
    __syncthreads();

    if (m.nan_default) {
        if (status) up_tile_dispatch<true, RANGE, true>(p, m, s_val, s_st, o0, rows);
        else up_tile_dispatch<true, RANGE, false>(p, m, s_val, s_st, o0, rows);
    } else {
        if (status) up_tile_dispatch<false, RANGE, true>(p, m, s_val, s_st, o0, rows);
        else up_tile_dispatch<false, RANGE, false>(p, m, s_val, s_st, o0, rows);
    }
}

inline int launch_up_tile(const UpMeasure* d_meas, int n, bool contiguous, const int32_t* d_pstart,
                          const int32_t* d_children, int64_t O, int64_t C, int64_t P, int64_t I,
                          const TileDecision& t) {
    UpTileParams p{};
    p.meas = d_meas;
    p.pstart = d_pstart;
    p.children = d_children;
    p.O = O; p.C = (int32_t)C; p.P = (int32_t)P; p.I = (int32_t)I;
    p.R = t.R;
    p.row_in = (int32_t)(C * I);
    p.row_out = (int32_t)(P * I);
    p.div_row_out = FastDiv((uint32_t)p.row_out);
    p.div_i = FastDiv((uint32_t)I);
    p.bulk_values = t.bulk_values;
    p.bulk_status = t.bulk_status;
    p.st_offset = t.st_offset;
    const int64_t tiles = ceil_div(O, t.R);
    if (tiles > 0x7fffffffLL) return fail(OLAP_E_UNSUPPORTED, "drillUp: grid too large (%lld tiles)", (long long)tiles);
    static bool attr_set[2] = {false, false};
    auto kern = contiguous ? drillup_tile_kernel<true> : drillup_tile_kernel<false>;
    if (!attr_set[contiguous]) {
        OLAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[contiguous] = true;
    }
    kern<<<dim3((unsigned)tiles, (unsigned)n), 256, t.smem, g.stream>>>(p);
    ++g_launches;
    return OLAP_OK;
}

struct TransposePlan {
    bool use = false;
};
inline TransposePlan transpose_plan(const std::vector<GDim>&) { return TransposePlan{}; }
inline int launch_transpose(const std::vector<GatherMeasure>&, int, const TransposePlan&) {
    return fail(OLAP_E_UNSUPPORTED, "transpose path not built");
}

}  // namespace olap
