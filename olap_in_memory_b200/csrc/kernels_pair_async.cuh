// transpose_async_kernel — the tiles of transpose_pair_kernel (kernels_pair.cuh: A cells along the input's
// contiguous run x B cells along the output's) moved the way gather_inner_flat_kernel moves its blocks: nothing
// passes through registers on the way in.
//   * persistent CTAs, two per SM; per tile every warp takes whole INPUT runs and copies them with 8-byte
//     cp.async into an input-ordered shared tile S[j][i] (all copies of a tile in flight at once, status bytes as
//     4-byte copies into a byte tile of their own); no register staging, no shared-memory transposition pass;
//   * the pitch of S is A + 2 words (an odd number of 8-byte units): the output phase reads S[j][i] with lanes
//     along j (consecutive cells of an OUTPUT run) at 2-way bank conflicts, and every warp writes whole output
//     runs with 128-byte coalesced stores (status: 32-byte sectors);
//   * lane offsets (j * pitch) are computed once per CTA, run offsets come from the plan's tables in shared memory.
// The register-staged kernel keeps 40 KB of loads in flight per SM and spends three CTA barriers per tile; this one
// keeps a whole tile (80 KB) per CTA in flight while the other CTA of the SM drains its own.  The recipe that took
// the block-local rearrangements (contiguous spans) from 0.66 to 0.93 of peak does NOT help here: bit-exact, but
// 2.96 against 2.35 ms on the 6-D reversal — with 400 / 800-byte runs on both sides neither bytes in flight nor
// instruction count is the limit.  Kept opt-in (OLAP_PAIR_ASYNC=1) with its test, like the tensor-map variant.
#pragma once

#include "kernels_pair.cuh"

namespace olap {

constexpr int kAsyncThreads = 512, kAsyncWarps = kAsyncThreads / 32, kAsyncMaxJ = 8;  // B <= 256 cells per output run of a tile

struct AsyncGeo {
    uint32_t PA, PS;             // pitch of the value tile (cells) and of the status tile (bytes)
    uint32_t st_off, tab_off;    // byte offsets inside the dynamic shared memory
    size_t smem;
};

inline AsyncGeo async_geo(const PairParams& p, bool loaded_status) {
    AsyncGeo g{};
    g.PA = p.A + 2;                                   // A % 4 == 0: (A + 2) / 2 is odd
    g.PS = ((p.A / 4) & 1) ? p.A : p.A + 4;           // an odd number of words
    g.st_off = (uint32_t)(((size_t)p.B * g.PA * 4 + 15) & ~(size_t)15);
    g.tab_off = g.st_off + (loaded_status ? (uint32_t)(((size_t)p.B * g.PS + 15) & ~(size_t)15) : 0u);
    g.smem = g.tab_off + ((size_t)p.A + p.B) * 4;
    return g;
}

__device__ __forceinline__ void async_cp8(void* smem_dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void async_cp4(void* smem_dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}

static __global__ void __launch_bounds__(kAsyncThreads, 2) transpose_async_kernel(const __grid_constant__ PairParams p,
                                                                                  const __grid_constant__ AsyncGeo geo, uint32_t n_boxes) {
    extern __shared__ __align__(16) unsigned char smem_a[];
    float* s_val = reinterpret_cast<float*>(smem_a);
    uint8_t* s_st = smem_a + geo.st_off;
    uint32_t* s_src_row = reinterpret_cast<uint32_t*>(smem_a + geo.tab_off);
    uint32_t* s_dst_row = s_src_row + p.B;
    const GatherMeasure m = p.meas[blockIdx.y];
    for (uint32_t i = threadIdx.x; i < p.B; i += kAsyncThreads) s_src_row[i] = __ldg(p.src_row + i);
    for (uint32_t i = threadIdx.x; i < p.A; i += kAsyncThreads) s_dst_row[i] = __ldg(p.dst_row + i);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool load_plane = m.st_in != nullptr, write_plane = m.st_out != nullptr, nan_default = m.nan_default != 0;
    const uint32_t nJ = (p.B + 31) >> 5;  // lane positions along an output run
    for (uint32_t tile = blockIdx.x; tile < n_boxes; tile += gridDim.x) {
        int64_t sb, db;
        uint32_t a_eff, b_eff;
        pair_decode(p, tile, sb, db, a_eff, b_eff);
        // ---- in: warp w copies input runs j = w, w + 16, ... (a_eff cells each)
        for (uint32_t j = warp; j < b_eff; j += kAsyncWarps) {
            const size_t row = (size_t)s_src_row[j] << 2;
            const float* g_in = m.in + sb + row;
            float* s_row = s_val + (size_t)j * geo.PA;
            for (uint32_t i2 = lane; 2 * i2 < a_eff; i2 += 32) async_cp8(s_row + 2 * i2, g_in + 2 * i2);
            if (load_plane) {
                const uint8_t* g_st = m.st_in + sb + row;
                uint8_t* t_row = s_st + (size_t)j * geo.PS;
                for (uint32_t i4 = lane; 4 * i4 < a_eff; i4 += 32) async_cp4(t_row + 4 * i4, g_st + 4 * i4);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- out: warp w writes output runs i = w, w + 16, ... (b_eff cells each), lanes along the run
        for (uint32_t i = warp; i < a_eff; i += kAsyncWarps) {
            const size_t run = (size_t)s_dst_row[i] << 2;
            float* g_out = m.out + db + run + lane;
            const float* s_col = s_val + i + (size_t)lane * geo.PA;
            if (!write_plane) {
#pragma unroll
                for (uint32_t jj = 0; jj < kAsyncMaxJ; ++jj)
                    if (jj < nJ && lane + 32 * jj < b_eff) g_out[32 * jj] = s_col[(size_t)32 * jj * geo.PA];
            } else {
                uint8_t* g_so = m.st_out + db + run + lane;
                const uint8_t* t_col = s_st + i + (size_t)lane * geo.PS;
#pragma unroll
                for (uint32_t jj = 0; jj < kAsyncMaxJ; ++jj)
                    if (jj < nJ && lane + 32 * jj < b_eff) {
                        const float v = s_col[(size_t)32 * jj * geo.PA];
                        g_out[32 * jj] = v;
                        g_so[32 * jj] = load_plane ? t_col[(size_t)32 * jj * geo.PS]
                                                   : (uint8_t)(present_f(v, nan_default) ? OLAP_STATUS_SET : OLAP_STATUS_UNSET);  // derived plane
                    }
            }
        }
        __syncthreads();  // the next tile overwrites the staged runs
    }
}

// Same plan, tables and tile numbering as launch_transpose_pair (non-cluster plans only).
inline bool transpose_async_fits(const PairPlan& plan, bool loaded_status) {
    // opt-in (read at every call so that a test can switch it on): measured SLOWER than the register-staged kernel on
    // the 6-D reversal of 100^3 x 10^3 (2.96 against 2.35 ms with a loaded plane, 2.54 against 2.16 derived)
    const char* e = getenv("OLAP_PAIR_ASYNC");
    const int knob = e ? atoi(e) : 0;
    if (!knob || !plan.use || plan.p.split != 1 || plan.p.B > 32u * kAsyncMaxJ) return false;
    return async_geo(plan.p, loaded_status).smem <= 110 * 1024;
}

inline int launch_transpose_async(const GatherMeasure* d_meas, const uint32_t* d_src_row, const uint32_t* d_dst_row, int n,
                                  PairPlan& plan, bool loaded_status) {
    plan.p.meas = d_meas;
    plan.p.src_row = d_src_row;
    plan.p.dst_row = d_dst_row;
    const AsyncGeo geo = async_geo(plan.p, loaded_status);
    static bool attr_set = false;
    if (!attr_set) {
        OLAP_CUDA(cudaFuncSetAttribute(transpose_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        attr_set = true;
    }
    const int64_t ctas = std::min<int64_t>(plan.n_boxes, std::max<int64_t>(1, ceil_div((int64_t)g.sm_count * 2, n)));
    mark_kernels_begin();
    transpose_async_kernel<<<dim3((unsigned)ctas, (unsigned)n), kAsyncThreads, geo.smem, g.stream>>>(plan.p, geo, (uint32_t)plan.n_boxes);
    ++g_launches;
    return OLAP_OK;
}

}  // namespace olap
