// Reorder (in-memory.js:178-211) when the innermost axis moves and the cube offers two
// DISJOINT axis groups of >= 32 cells: the input's trailing axes (a contiguous INPUT run of A
// cells) and the output's trailing axes (a contiguous OUTPUT run of B cells).  A CTA owns the
// A x B tile spanned by the two groups (every other axis has extent 1):
//
//   phase 1  every thread owns 4x4 micro-tiles: four 128-bit loads along four input runs (and
//            four 32-bit loads of their status bytes), a register transpose (free for floats,
//            8 PRMT for the bytes) and four 128-bit / 32-bit shared-memory stores into an
//            OUTPUT-ordered tile  s[i & 3][i >> 2][j]  (pitch 4*odd: conflict-free);
//   phase 2  128-bit shared-memory loads along j, 128-bit coalesced stores along output runs.
//
// ~3 thread-instructions per cell instead of ~70 for the generic box kernel, every global
// access is 16 bytes per lane, every run a contiguous span.  CTAs are numbered so that
// neighbours in the input's memory order (and then in the output's) run at the same time:
// runs that start or end inside a DRAM atom share it through L2 instead of fetching it twice.
// Needs run lengths and strides that are multiples of 4 cells; everything else stays on
// transpose_kernel (kernels_tile.cuh).
#pragma once

#include <algorithm>
#include <vector>

#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

#include "common.cuh"
#include "kernels_gather.cuh"

namespace olap {

struct PairParams {
    const GatherMeasure* meas;
    uint32_t A, B;            // cells of a full tile along the input run / output run
    uint32_t nIg, nJq;        // A / 4, B / 4
    FastDiv div_nIg, div_nJq;
    uint32_t PB;              // shared-memory pitch of one [k][ig] row, cells (4 * odd)
    const uint32_t* src_row;  // [B] tile-relative source offset / 4 of output-run position j
    const uint32_t* dst_row;  // [A] tile-relative destination offset / 4 of input-run position i
    int in_axis, out_axis;    // grid slot of the partially covered axis of each group, or -1
    uint32_t in_mult, out_mult;  // cells of the fully covered axes of the group
    // grid decomposition: slots in traversal order, the LAST slot varies fastest
    int n_axes;
    uint32_t boxes[OLAP_MAX_DIMS];
    FastDiv div_boxes[OLAP_MAX_DIMS];
    uint32_t len[OLAP_MAX_DIMS], bsize[OLAP_MAX_DIMS];
    int64_t src_stride[OLAP_MAX_DIMS], dst_stride[OLAP_MAX_DIMS];
    uint32_t st_offset, tab_offset;  // byte offsets of the status tile / the staged tables
    // split == 2: a 2-CTA cluster shares one tile through distributed shared memory.  CTA r
    // LOADS the input rows of micro-row groups [r*nJqLoc, (r+1)*nJqLoc) and OWNS (stores to
    // global) the output runs i in [r*A/2, (r+1)*A/2): both sides see full-length runs.
    uint32_t split, nIgLoc, nJqLoc;
    FastDiv div_nJqLoc;
};

// One (merged) axis of the permutation as the planner sees it — what the tensor-map variant
// (kernels_tma.cuh) needs to describe the tile to the TMA unit.
struct PairAxis {
    int64_t len, box;                // items, items of it inside one tile
    int64_t src_stride, dst_stride;  // cells
    int group;                       // 1: part of the input run, 2: part of the output run, 0: extent 1
    int slot;                        // grid slot (PairParams::len/bsize/... index) of this axis
};

struct PairPlan {
    bool use = false;
    PairParams p{};
    int64_t n_boxes = 0;
    size_t smem = 0;
    std::vector<uint32_t> src_row, dst_row;
    std::vector<PairAxis> axes;
    bool geometry = false;           // A, B, tables, grid and axes are valid (even when `use` is false for lack of shared memory)
};

// Memory accessors: the device flavour streams through L1; the host flavour lets
// tests/host/plan_check.cu run the very same phase code on the CPU.
struct PairDevMem {
    static __device__ __forceinline__ float4 ld4(const float* q) { return ld_stream4(q); }
    static __device__ __forceinline__ uint32_t ld_u32(const uint8_t* q) { return ld_stream_u32(q); }
    static __device__ __forceinline__ void st4(float* q, float4 v) { st_stream4(q, v); }
};
struct PairHostMem {
    static __host__ __device__ float4 ld4(const float* q) { return *reinterpret_cast<const float4*>(q); }
    static __host__ __device__ uint32_t ld_u32(const uint8_t* q) { return *reinterpret_cast<const uint32_t*>(q); }
    static __host__ __device__ void st4(float* q, float4 v) { *reinterpret_cast<float4*>(q) = v; }
};
// PRMT: byte n of the result is byte (sel >> 4n & 7) of {x, y}
__host__ __device__ __forceinline__ uint32_t pair_prmt(uint32_t x, uint32_t y, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, y, sel);
#else
    const uint64_t xy = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int n = 0; n < 4; ++n) r |= (uint32_t)((xy >> (8 * ((sel >> (4 * n)) & 7))) & 0xff) << (8 * n);
    return r;
#endif
}

// Which tile is block `block`: source / destination base offsets and the valid run lengths.
__host__ __device__ __forceinline__ void pair_decode(const PairParams& p, uint32_t block, int64_t& src, int64_t& dst,
                                                     uint32_t& a_eff, uint32_t& b_eff) {
    uint32_t rest = block;
    a_eff = p.A;
    b_eff = p.B;
    src = dst = 0;
    for (int a = p.n_axes - 1; a >= 0; --a) {
        const uint32_t q = p.div_boxes[a].div(rest);
        const uint32_t bi = rest - q * p.boxes[a];
        rest = q;
        const uint32_t start = bi * p.bsize[a];
        src += (int64_t)start * p.src_stride[a];
        dst += (int64_t)start * p.dst_stride[a];
        const uint32_t rem = p.len[a] - start, e = p.bsize[a] < rem ? p.bsize[a] : rem;
        if (a == p.in_axis) a_eff = e * p.in_mult;
        if (a == p.out_axis) b_eff = e * p.out_mult;
    }
}

// ---- phase 1: one 4x4 micro-tile per thread.  The loads (four input runs x 4 cells, and
// their status bytes) land in registers, so a persistent CTA can issue the NEXT tile's loads
// before it drains the current tile from shared memory: loads stay in flight all the time.
struct PairRegs {
    float4 v0, v1, v2, v3;
    uint32_t b0, b1, b2, b3;
    uint32_t sidx;  // shared-memory index of the micro-tile, or 0xffffffff: nothing loaded
};

template <bool STATUS, class Mem>
__host__ __device__ __forceinline__ void pair_load(const PairParams& p, uint32_t tid, const float* src,
                                                   const uint8_t* st_src, const uint32_t* s_src_row, uint32_t n_ig,
                                                   uint32_t n_jq, PairRegs& r) {
    const uint32_t jg = p.div_nIg.div(tid), ig = tid - jg * p.nIg;
    r.sidx = 0xffffffffu;
    if (ig < n_ig && jg < n_jq) {
        const uint4 ro = *reinterpret_cast<const uint4*>(s_src_row + 4 * jg);
        const size_t o0 = ((size_t)ro.x + ig) << 2, o1 = ((size_t)ro.y + ig) << 2;
        const size_t o2 = ((size_t)ro.z + ig) << 2, o3 = ((size_t)ro.w + ig) << 2;
        r.v0 = Mem::ld4(src + o0); r.v1 = Mem::ld4(src + o1);
        r.v2 = Mem::ld4(src + o2); r.v3 = Mem::ld4(src + o3);
        if (STATUS) {
            r.b0 = Mem::ld_u32(st_src + o0); r.b1 = Mem::ld_u32(st_src + o1);
            r.b2 = Mem::ld_u32(st_src + o2); r.b3 = Mem::ld_u32(st_src + o3);
        }
        r.sidx = ig * p.PB + 4 * jg;
    }
}

// register transpose (free for the floats, 8 PRMT for the bytes) into the OUTPUT-ordered tile
template <bool STATUS>
__host__ __device__ __forceinline__ void pair_stash(const PairParams& p, const PairRegs& r, float* s_val, uint8_t* s_st) {
    if (r.sidx == 0xffffffffu) return;
    const uint32_t plane = p.nIg * p.PB, sidx = r.sidx;
    *reinterpret_cast<float4*>(s_val + sidx) = make_float4(r.v0.x, r.v1.x, r.v2.x, r.v3.x);
    *reinterpret_cast<float4*>(s_val + sidx + plane) = make_float4(r.v0.y, r.v1.y, r.v2.y, r.v3.y);
    *reinterpret_cast<float4*>(s_val + sidx + 2 * plane) = make_float4(r.v0.z, r.v1.z, r.v2.z, r.v3.z);
    *reinterpret_cast<float4*>(s_val + sidx + 3 * plane) = make_float4(r.v0.w, r.v1.w, r.v2.w, r.v3.w);
    if (STATUS) {
        // 4x4 byte transpose: word k collects byte k of the four rows
        const uint32_t p01 = pair_prmt(r.b0, r.b1, 0x5140), q01 = pair_prmt(r.b0, r.b1, 0x7362);
        const uint32_t p23 = pair_prmt(r.b2, r.b3, 0x5140), q23 = pair_prmt(r.b2, r.b3, 0x7362);
        *reinterpret_cast<uint32_t*>(s_st + sidx) = pair_prmt(p01, p23, 0x5410);
        *reinterpret_cast<uint32_t*>(s_st + sidx + plane) = pair_prmt(p01, p23, 0x7632);
        *reinterpret_cast<uint32_t*>(s_st + sidx + 2 * plane) = pair_prmt(q01, q23, 0x5410);
        *reinterpret_cast<uint32_t*>(s_st + sidx + 3 * plane) = pair_prmt(q01, q23, 0x7632);
    }
}

// ---- phase 2: output runs, 4 cells per lane
// STATUS: the status bytes travel through the shared tile.  nan_default >= 0 (and !STATUS): the source's status
// plane is derived from its values — nothing is loaded or staged, the bytes are recomputed from the cells on
// their way out (-1: no status plane at all).
template <bool STATUS, class Mem>
__host__ __device__ __forceinline__ void pair_phase2(const PairParams& p, uint32_t tid, float* dst, uint8_t* st_dst,
                                                     const float* s_val, const uint8_t* s_st,
                                                     const uint32_t* s_dst_row, uint32_t a_eff, uint32_t n_jq,
                                                     uint32_t n_threads, int nan_default = -1) {
    const uint32_t n_it = p.A * p.nJq;
    for (uint32_t it = tid; it < n_it; it += n_threads) {
        const uint32_t i = p.div_nJq.div(it), jq = it - i * p.nJq;
        if (i < a_eff && jq < n_jq) {
            const uint32_t sidx = ((i & 3u) * p.nIg + (i >> 2)) * p.PB + 4 * jq;
            const float4 v = *reinterpret_cast<const float4*>(s_val + sidx);
            const size_t g = ((size_t)s_dst_row[i] + jq) << 2;
            Mem::st4(dst + g, v);
            if (STATUS) *reinterpret_cast<uint32_t*>(st_dst + g) = *reinterpret_cast<const uint32_t*>(s_st + sidx);
            else if (nan_default >= 0) {
                const bool nd = nan_default != 0;
                const uint32_t b0 = (nd ? v.x == v.x : v.x != 0.0f) ? 2u : 1u, b1 = (nd ? v.y == v.y : v.y != 0.0f) ? 2u : 1u;
                const uint32_t b2 = (nd ? v.z == v.z : v.z != 0.0f) ? 2u : 1u, b3 = (nd ? v.w == v.w : v.w != 0.0f) ? 2u : 1u;
                *reinterpret_cast<uint32_t*>(st_dst + g) = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
            }
        }
    }
}

// ---- cluster variant (split == 2) ---------------------------------------------------------
struct PairRegs2 {
    float4 v0, v1, v2, v3;
    uint32_t b0, b1, b2, b3;
    uint32_t sidx;  // index inside the OWNER's tile, or 0xffffffff
    uint32_t owner; // rank of the CTA that owns these output runs
};

// micro-tile `mt` of the rows this CTA (rank) loads
template <bool STATUS, class Mem>
__host__ __device__ __forceinline__ void pair2_load(const PairParams& p, uint32_t rank, uint32_t mt, const float* src,
                                                    const uint8_t* st_src, const uint32_t* s_src_row, uint32_t n_ig,
                                                    uint32_t n_jq, PairRegs2& r) {
    const uint32_t jl = p.div_nIg.div(mt), ig = mt - jl * p.nIg, jg = rank * p.nJqLoc + jl;
    r.sidx = 0xffffffffu;
    r.owner = 0;
    if (jl < p.nJqLoc && ig < n_ig && jg < n_jq) {
        const uint4 ro = *reinterpret_cast<const uint4*>(s_src_row + 4 * jg);
        const size_t o0 = ((size_t)ro.x + ig) << 2, o1 = ((size_t)ro.y + ig) << 2;
        const size_t o2 = ((size_t)ro.z + ig) << 2, o3 = ((size_t)ro.w + ig) << 2;
        r.v0 = Mem::ld4(src + o0); r.v1 = Mem::ld4(src + o1);
        r.v2 = Mem::ld4(src + o2); r.v3 = Mem::ld4(src + o3);
        if (STATUS) {
            r.b0 = Mem::ld_u32(st_src + o0); r.b1 = Mem::ld_u32(st_src + o1);
            r.b2 = Mem::ld_u32(st_src + o2); r.b3 = Mem::ld_u32(st_src + o3);
        }
        r.owner = ig >= p.nIgLoc ? 1u : 0u;
        r.sidx = (ig - r.owner * p.nIgLoc) * p.PB + 4 * jg;
    }
}

// s_val_of[r] / s_st_of[r]: the tile of CTA r (own shared memory or the peer's, mapped)
template <bool STATUS>
__host__ __device__ __forceinline__ void pair2_stash(const PairParams& p, const PairRegs2& r, float* s_val0,
                                                     float* s_val1, uint8_t* s_st0, uint8_t* s_st1) {
    if (r.sidx == 0xffffffffu) return;
    float* s_val = r.owner ? s_val1 : s_val0;
    uint8_t* s_st = r.owner ? s_st1 : s_st0;
    const uint32_t plane = p.nIgLoc * p.PB, sidx = r.sidx;
    *reinterpret_cast<float4*>(s_val + sidx) = make_float4(r.v0.x, r.v1.x, r.v2.x, r.v3.x);
    *reinterpret_cast<float4*>(s_val + sidx + plane) = make_float4(r.v0.y, r.v1.y, r.v2.y, r.v3.y);
    *reinterpret_cast<float4*>(s_val + sidx + 2 * plane) = make_float4(r.v0.z, r.v1.z, r.v2.z, r.v3.z);
    *reinterpret_cast<float4*>(s_val + sidx + 3 * plane) = make_float4(r.v0.w, r.v1.w, r.v2.w, r.v3.w);
    if (STATUS) {
        const uint32_t p01 = pair_prmt(r.b0, r.b1, 0x5140), q01 = pair_prmt(r.b0, r.b1, 0x7362);
        const uint32_t p23 = pair_prmt(r.b2, r.b3, 0x5140), q23 = pair_prmt(r.b2, r.b3, 0x7362);
        *reinterpret_cast<uint32_t*>(s_st + sidx) = pair_prmt(p01, p23, 0x5410);
        *reinterpret_cast<uint32_t*>(s_st + sidx + plane) = pair_prmt(p01, p23, 0x7632);
        *reinterpret_cast<uint32_t*>(s_st + sidx + 2 * plane) = pair_prmt(q01, q23, 0x5410);
        *reinterpret_cast<uint32_t*>(s_st + sidx + 3 * plane) = pair_prmt(q01, q23, 0x7632);
    }
}

// output runs this CTA owns: i = rank * A/2 + il
template <bool STATUS, class Mem>
__host__ __device__ __forceinline__ void pair2_phase2(const PairParams& p, uint32_t rank, uint32_t tid, float* dst,
                                                      uint8_t* st_dst, const float* s_val, const uint8_t* s_st,
                                                      const uint32_t* s_dst_row, uint32_t a_eff, uint32_t n_jq,
                                                      uint32_t n_threads) {
    const uint32_t a_loc = p.nIgLoc * 4, n_it = a_loc * p.nJq;
    for (uint32_t it = tid; it < n_it; it += n_threads) {
        const uint32_t il = p.div_nJq.div(it), jq = it - il * p.nJq, i = rank * a_loc + il;
        if (i < a_eff && jq < n_jq) {
            const uint32_t sidx = ((il & 3u) * p.nIgLoc + (il >> 2)) * p.PB + 4 * jq;
            const float4 v = *reinterpret_cast<const float4*>(s_val + sidx);
            const size_t g = ((size_t)s_dst_row[i] + jq) << 2;
            Mem::st4(dst + g, v);
            if (STATUS) *reinterpret_cast<uint32_t*>(st_dst + g) = *reinterpret_cast<const uint32_t*>(s_st + sidx);
        }
    }
}

constexpr int kPairThreads = 640;  // >= micro-tiles of the largest tile (25 x 25)

// Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...; NM micro-tiles per thread.
template <bool STATUS, int NM, bool DERIVE = false>
__device__ __forceinline__ void pair_body(const PairParams& p, const GatherMeasure& m, unsigned char* smem_p,
                                          uint32_t n_boxes) {
    float* s_val = reinterpret_cast<float*>(smem_p);
    uint8_t* s_st = smem_p + p.st_offset;
    const uint32_t* s_src_row = reinterpret_cast<const uint32_t*>(smem_p + p.tab_offset);
    const uint32_t* s_dst_row = s_src_row + p.B;
    uint32_t tile = blockIdx.x;
    int64_t sb, db;
    uint32_t a_eff, b_eff;
    PairRegs r[NM];
    pair_decode(p, tile, sb, db, a_eff, b_eff);
#pragma unroll
    for (int q = 0; q < NM; ++q)
        pair_load<STATUS, PairDevMem>(p, threadIdx.x + q * blockDim.x, m.in + sb, STATUS ? m.st_in + sb : nullptr,
                                      s_src_row, a_eff >> 2, b_eff >> 2, r[q]);
    while (true) {
#pragma unroll
        for (int q = 0; q < NM; ++q) pair_stash<STATUS>(p, r[q], s_val, s_st);
        __syncthreads();  // the tile is complete in shared memory
        const int64_t db_cur = db;
        const uint32_t a_cur = a_eff, nj_cur = b_eff >> 2;
        tile += gridDim.x;
        const bool more = tile < n_boxes;
        if (more) {  // next tile's loads fly while this one drains
            pair_decode(p, tile, sb, db, a_eff, b_eff);
#pragma unroll
            for (int q = 0; q < NM; ++q)
                pair_load<STATUS, PairDevMem>(p, threadIdx.x + q * blockDim.x, m.in + sb,
                                              STATUS ? m.st_in + sb : nullptr, s_src_row, a_eff >> 2, b_eff >> 2, r[q]);
        }
        pair_phase2<STATUS, PairDevMem>(p, threadIdx.x, m.out + db_cur, (STATUS || DERIVE) ? m.st_out + db_cur : nullptr, s_val,
                                        s_st, s_dst_row, a_cur, nj_cur, blockDim.x, DERIVE ? m.nan_default : -1);
        if (!more) break;
        __syncthreads();  // everyone has drained the tile
    }
}

template <int NM>
__global__ void __launch_bounds__(kPairThreads, NM == 1 ? 2 : 1) transpose_pair_kernel(const __grid_constant__ PairParams p, uint32_t n_boxes) {
    extern __shared__ __align__(16) unsigned char smem_p[];
    uint32_t* s_src_row = reinterpret_cast<uint32_t*>(smem_p + p.tab_offset);
    uint32_t* s_dst_row = s_src_row + p.B;
    const GatherMeasure m = p.meas[blockIdx.y];
    for (uint32_t i = threadIdx.x; i < p.B; i += blockDim.x) s_src_row[i] = __ldg(p.src_row + i);
    for (uint32_t i = threadIdx.x; i < p.A; i += blockDim.x) s_dst_row[i] = __ldg(p.dst_row + i);
    __syncthreads();
    if (m.st_in) pair_body<true, NM>(p, m, smem_p, n_boxes);
    else if (m.derive && m.st_out) pair_body<false, NM, true>(p, m, smem_p, n_boxes);
    else pair_body<false, NM>(p, m, smem_p, n_boxes);
}

#ifdef __CUDACC__
// ---- cluster plumbing (PTX): remote shared-memory addresses, asynchronous remote stores that
// complete a transaction count on the OWNER's mbarrier (no fence anywhere on the data path)
__device__ __forceinline__ uint32_t pair_smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ uint32_t pair_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void pair_st_async4(uint32_t addr, float4 v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void pair_st_async1(uint32_t addr, uint32_t v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void pair_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pair_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pair_mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pair_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pair_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PDONE_%=;\n"
        "bra PWAIT_%=;\n"
        "PDONE_%=:\n"
        "}\n" ::"r"(pair_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// stash of the cluster variant: own output runs -> plain shared-memory stores, the peer's ->
// st.async into the peer's tile, counted on the peer's mbarrier
template <bool STATUS>
__device__ __forceinline__ void pair2_stash_dev(const PairParams& p, const PairRegs2& r, uint32_t rank, float* s_val,
                                                uint8_t* s_st, uint32_t peer_val, uint32_t peer_st, uint32_t peer_bar) {
    if (r.sidx == 0xffffffffu) return;
    const uint32_t plane = p.nIgLoc * p.PB, sidx = r.sidx;
    const float4 c0 = make_float4(r.v0.x, r.v1.x, r.v2.x, r.v3.x), c1 = make_float4(r.v0.y, r.v1.y, r.v2.y, r.v3.y);
    const float4 c2 = make_float4(r.v0.z, r.v1.z, r.v2.z, r.v3.z), c3 = make_float4(r.v0.w, r.v1.w, r.v2.w, r.v3.w);
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (STATUS) {
        const uint32_t p01 = pair_prmt(r.b0, r.b1, 0x5140), q01 = pair_prmt(r.b0, r.b1, 0x7362);
        const uint32_t p23 = pair_prmt(r.b2, r.b3, 0x5140), q23 = pair_prmt(r.b2, r.b3, 0x7362);
        w0 = pair_prmt(p01, p23, 0x5410); w1 = pair_prmt(p01, p23, 0x7632);
        w2 = pair_prmt(q01, q23, 0x5410); w3 = pair_prmt(q01, q23, 0x7632);
    }
    if (r.owner == rank) {
        *reinterpret_cast<float4*>(s_val + sidx) = c0;
        *reinterpret_cast<float4*>(s_val + sidx + plane) = c1;
        *reinterpret_cast<float4*>(s_val + sidx + 2 * plane) = c2;
        *reinterpret_cast<float4*>(s_val + sidx + 3 * plane) = c3;
        if (STATUS) {
            *reinterpret_cast<uint32_t*>(s_st + sidx) = w0;
            *reinterpret_cast<uint32_t*>(s_st + sidx + plane) = w1;
            *reinterpret_cast<uint32_t*>(s_st + sidx + 2 * plane) = w2;
            *reinterpret_cast<uint32_t*>(s_st + sidx + 3 * plane) = w3;
        }
    } else {
        pair_st_async4(peer_val + 4u * sidx, c0, peer_bar);
        pair_st_async4(peer_val + 4u * (sidx + plane), c1, peer_bar);
        pair_st_async4(peer_val + 4u * (sidx + 2 * plane), c2, peer_bar);
        pair_st_async4(peer_val + 4u * (sidx + 3 * plane), c3, peer_bar);
        if (STATUS) {
            pair_st_async1(peer_st + sidx, w0, peer_bar);
            pair_st_async1(peer_st + sidx + plane, w1, peer_bar);
            pair_st_async1(peer_st + sidx + 2 * plane, w2, peer_bar);
            pair_st_async1(peer_st + sidx + 3 * plane, w3, peer_bar);
        }
    }
}

// bytes CTA `rank` receives from its peer for a tile with n_ig x n_jq valid micro columns / rows
__device__ __forceinline__ uint32_t pair2_expected(const PairParams& p, uint32_t rank, uint32_t n_ig, uint32_t n_jq,
                                                   bool status) {
    const uint32_t lo_i = rank * p.nIgLoc, lo_j = (rank ^ 1u) * p.nJqLoc;
    const uint32_t mine = n_ig > lo_i ? min(n_ig - lo_i, p.nIgLoc) : 0u;
    const uint32_t theirs = n_jq > lo_j ? min(n_jq - lo_j, p.nJqLoc) : 0u;
    return mine * theirs * 4u * (status ? 20u : 16u);
}

// 2-CTA cluster, persistent: cluster c walks tiles c, c + n_clusters, ...
template <bool STATUS, int NM>
__device__ __forceinline__ void pair2_body(const PairParams& p, const GatherMeasure& m, unsigned char* smem_p,
                                           uint32_t n_boxes, uint64_t* bar) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    float* s_val = reinterpret_cast<float*>(smem_p);
    uint8_t* s_st = smem_p + p.st_offset;
    const uint32_t* s_src_row = reinterpret_cast<const uint32_t*>(smem_p + p.tab_offset);
    const uint32_t* s_dst_row = s_src_row + p.B;
    const uint32_t peer_val = pair_mapa(pair_smem_u32(s_val), rank ^ 1u);
    const uint32_t peer_st = pair_mapa(pair_smem_u32(s_st), rank ^ 1u);
    const uint32_t peer_bar = pair_mapa(pair_smem_u32(bar), rank ^ 1u);
    const uint32_t n_clusters = gridDim.x >> 1;
    uint32_t tile = blockIdx.x >> 1;
    int64_t sb, db;
    uint32_t a_eff, b_eff;
    PairRegs2 r[NM];
    pair_decode(p, tile, sb, db, a_eff, b_eff);
    if (threadIdx.x == 0) {
        pair_mbar_init(bar, 1);
        pair_mbar_expect(bar, pair2_expected(p, rank, a_eff >> 2, b_eff >> 2, STATUS));  // armed for the first tile
    }
    cluster.sync();  // the peer's barrier exists and is armed before anything is sent to it
#pragma unroll
    for (int q = 0; q < NM; ++q)
        pair2_load<STATUS, PairDevMem>(p, rank, threadIdx.x + q * blockDim.x, m.in + sb, STATUS ? m.st_in + sb : nullptr,
                                       s_src_row, a_eff >> 2, b_eff >> 2, r[q]);
    uint32_t parity = 0;
    while (true) {
#pragma unroll
        for (int q = 0; q < NM; ++q) pair2_stash_dev<STATUS>(p, r[q], rank, s_val, s_st, peer_val, peer_st, peer_bar);
        __syncthreads();              // my own half of the stash is visible to my threads
        pair_mbar_wait(bar, parity);  // ... and every byte the peer owes me has landed
        parity ^= 1u;
        const int64_t db_cur = db;
        const uint32_t a_cur = a_eff, nj_cur = b_eff >> 2;
        tile += n_clusters;
        const bool more = tile < n_boxes;
        if (more) {
            pair_decode(p, tile, sb, db, a_eff, b_eff);
#pragma unroll
            for (int q = 0; q < NM; ++q)
                pair2_load<STATUS, PairDevMem>(p, rank, threadIdx.x + q * blockDim.x, m.in + sb,
                                               STATUS ? m.st_in + sb : nullptr, s_src_row, a_eff >> 2, b_eff >> 2, r[q]);
        }
        pair2_phase2<STATUS, PairDevMem>(p, rank, threadIdx.x, m.out + db_cur, STATUS ? m.st_out + db_cur : nullptr,
                                         s_val, s_st, s_dst_row, a_cur, nj_cur, blockDim.x);
        // arm my barrier for the next tile BEFORE the peer may send (it sends after the cluster
        // barrier below).  That barrier only orders execution (both CTAs have drained their
        // tiles): a RELAXED arrive — a release would make every thread wait for its own global
        // stores of phase 2 (ncu of the cluster.sync() version: `membar` was the top stall).
        if (more && threadIdx.x == 0) pair_mbar_expect(bar, pair2_expected(p, rank, a_eff >> 2, b_eff >> 2, STATUS));
        asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if (!more) break;
    }
}

template <int NM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
transpose_pair2_kernel(const __grid_constant__ PairParams p, uint32_t n_boxes) {
    extern __shared__ __align__(16) unsigned char smem_p[];
    uint32_t* s_src_row = reinterpret_cast<uint32_t*>(smem_p + p.tab_offset);
    uint32_t* s_dst_row = s_src_row + p.B;
    const GatherMeasure m = p.meas[blockIdx.y];
    for (uint32_t i = threadIdx.x; i < p.B; i += blockDim.x) s_src_row[i] = __ldg(p.src_row + i);
    for (uint32_t i = threadIdx.x; i < p.A; i += blockDim.x) s_dst_row[i] = __ldg(p.dst_row + i);
    __shared__ __align__(8) uint64_t s_bar;  // static: the same shared-memory offset in both CTAs of the cluster
    __syncthreads();
    if (m.st_in) pair2_body<true, NM>(p, m, smem_p, n_boxes, &s_bar);
    else pair2_body<false, NM>(p, m, smem_p, n_boxes, &s_bar);
}
#endif

// `dims_in`: the output axes (outermost first) as linear GDims carrying their SOURCE strides.
inline PairPlan transpose_pair_plan_for(const std::vector<GDim>& dims_in, int64_t cap_in, int64_t cap_out, int split = 1) {
    PairPlan plan;
    static const int force = [] { const char* e = getenv("OLAP_TRANSPOSE_PAIR"); return e ? atoi(e) : -1; }();
    if (force == 0) return plan;
    static const int64_t cap = [] { const char* e = getenv("OLAP_PAIR_CAP"); return e ? (int64_t)atoi(e) : (int64_t)100; }();
    std::vector<GDim> dims;
    for (const GDim& d : dims_in) {
        if (!d.linear) return plan;
        if (d.len == 1) continue;
        if (!dims.empty() && dims.back().stride == d.len * d.stride) {
            dims.back().len *= d.len;
            dims.back().stride = d.stride;
        } else dims.push_back(d);
    }
    const int k = (int)dims.size();
    if (k < 2 || k > OLAP_MAX_DIMS) return plan;
    if (dims.back().stride == 1) return plan;  // innermost axis stays: the vectorised gather streams it
    for (const GDim& d : dims)
        if (d.len > 0x7fffffffLL) return plan;
    std::vector<int64_t> dst_stride(k);
    int64_t acc = 1;
    for (int i = k - 1; i >= 0; --i) { dst_stride[i] = acc; acc *= dims[i].len; }
    std::vector<int> by_src(k), by_dst(k);
    for (int i = 0; i < k; ++i) { by_src[i] = i; by_dst[i] = k - 1 - i; }
    std::sort(by_src.begin(), by_src.end(), [&](int a, int b) { return dims[a].stride < dims[b].stride; });
    if (dims[by_src[0]].stride != 1) return plan;

    // a group: trailing axes taken whole while the run stays <= cap, then one partial axis
    struct Group { std::vector<int> axes; int64_t mult = 1, ext = 1, run = 1; bool partial = false; };
    auto grow = [&](const std::vector<int>& order, Group& gr, int64_t cap_g, int64_t want) {
        int64_t pr = 1;
        for (int ax : order) {
            const int64_t L = dims[ax].len;
            if (pr * L <= cap_g) {
                gr.axes.push_back(ax);
                pr *= L;
                gr.mult = pr;  // every axis whole so far
                gr.ext = 1;
                gr.partial = false;
                if (pr >= want) break;
                continue;
            }
            // partial axis: an extent e with (pr * e) % 4 == 0 for full and ragged tiles,
            // preferring an even split of the axis
            const int64_t e_max = std::min<int64_t>(L, cap_g / pr);
            if (e_max < 2 && pr >= 32) break;  // nothing to gain from a sliver of the next axis
            int64_t pick = 0;
            for (int64_t e = e_max; e >= std::max<int64_t>(1, e_max / 2); --e) {
                if ((pr * e) % 4 || ((L % e) * pr) % 4) continue;
                if (L % e == 0) { pick = e; break; }
                if (!pick) pick = e;
            }
            if (!pick) return false;
            if ((pr * L) % 4) return false;  // strides of the axes outside the group
            gr.axes.push_back(ax);
            gr.mult = pr;
            gr.ext = pick;
            gr.partial = true;
            pr *= pick;
            break;
        }
        gr.run = pr;
        return pr >= 32 && pr % 4 == 0;
    };
    Group gi, go;
    // either run may be asked to be longer than the other (two micro-tiles per thread, one CTA
    // per SM): fewer, longer DRAM bursts on that side
    if (!grow(by_src, gi, std::max(cap, cap_in), cap_in ? cap_in : 64) ||
        !grow(by_dst, go, std::max(cap, cap_out), cap_out ? cap_out : 64)) return plan;
    for (int a : gi.axes)
        for (int b : go.axes)
            if (a == b) return plan;
    if ((int)(gi.axes.size() + go.axes.size()) < k) {
        // some axis lies outside both groups: its strides are multiples of the whole groups
        int64_t full_i = 1, full_o = 1;
        for (int a : gi.axes) full_i *= dims[a].len;
        for (int a : go.axes) full_o *= dims[a].len;
        if (full_i % 4 || full_o % 4) return plan;
    }
    // strides seen from the other group must keep 16-byte alignment too
    for (int a : go.axes) if (dims[a].stride % 4) return plan;   // source rows
    for (int a : gi.axes) if (dst_stride[a] % 4) return plan;    // destination rows

    PairParams& p = plan.p;
    p.A = (uint32_t)gi.run;
    p.B = (uint32_t)go.run;
    p.nIg = p.A / 4;
    p.nJq = p.B / 4;
    p.div_nIg = FastDiv(p.nIg);
    p.div_nJq = FastDiv(p.nJq);
    uint32_t pq = p.nJq | 1u;  // odd number of 16-byte chunks per row
    p.PB = pq * 4;
    // extent of every axis inside the tile
    std::vector<int64_t> b(k, 1);
    for (size_t q = 0; q < gi.axes.size(); ++q) b[gi.axes[q]] = (gi.partial && q + 1 == gi.axes.size()) ? gi.ext : dims[gi.axes[q]].len;
    for (size_t q = 0; q < go.axes.size(); ++q) b[go.axes[q]] = (go.partial && q + 1 == go.axes.size()) ? go.ext : dims[go.axes[q]].len;
    // tables: position along a group -> offset on the OTHER side, in units of 4 cells
    auto table = [&](const Group& gr, bool src_side, std::vector<uint32_t>& tab) {
        tab.resize((size_t)gr.run);
        for (int64_t pos = 0; pos < gr.run; ++pos) {
            int64_t t = pos, off = 0;
            for (int ax : gr.axes) {
                const int64_t c = t % b[ax];
                t /= b[ax];
                off += c * (src_side ? dims[ax].stride : dst_stride[ax]);
            }
            if (off % 4 || off / 4 > 0xffffffffLL) return false;
            tab[(size_t)pos] = (uint32_t)(off / 4);
        }
        return true;
    };
    if (!table(go, true, plan.src_row) || !table(gi, false, plan.dst_row)) return plan;
    // traversal order of the grid: fastest = the axis along which tiles are neighbours in the
    // input, then the one along which they are neighbours in the output, then the rest
    std::vector<int> order;  // fastest first
    auto boxes_of = [&](int ax) { return ceil_div(dims[ax].len, b[ax]); };
    static const int order_knob = [] { const char* e = getenv("OLAP_PAIR_ORDER"); return e ? atoi(e) : 0; }();
    if (order_knob == 0)
    for (int ax : by_src) if (boxes_of(ax) > 1) { order.push_back(ax); break; }
    if (order_knob == 0 || order_knob == 2)
    for (int ax : by_dst) if (boxes_of(ax) > 1 && std::find(order.begin(), order.end(), ax) == order.end()) { order.push_back(ax); break; }
    if (order_knob == 2)
        for (int ax : by_src) if (boxes_of(ax) > 1 && std::find(order.begin(), order.end(), ax) == order.end()) { order.push_back(ax); break; }
    for (int ax : by_dst) if (std::find(order.begin(), order.end(), ax) == order.end()) order.push_back(ax);
    p.n_axes = k;
    p.in_axis = p.out_axis = -1;
    int64_t n_boxes = 1;
    for (int q = 0; q < k; ++q) {
        const int ax = order[q], slot = k - 1 - q;
        p.len[slot] = (uint32_t)dims[ax].len;
        p.bsize[slot] = (uint32_t)b[ax];
        p.boxes[slot] = (uint32_t)boxes_of(ax);
        p.div_boxes[slot] = FastDiv(p.boxes[slot]);
        p.src_stride[slot] = dims[ax].stride;
        p.dst_stride[slot] = dst_stride[ax];
        n_boxes *= p.boxes[slot];
        if (gi.partial && ax == gi.axes.back()) p.in_axis = slot;
        if (go.partial && ax == go.axes.back()) p.out_axis = slot;
    }
    plan.axes.resize(k);
    for (int q = 0; q < k; ++q) {
        const int ax = order[q];
        int group = 0;
        if (std::find(gi.axes.begin(), gi.axes.end(), ax) != gi.axes.end()) group = 1;
        if (std::find(go.axes.begin(), go.axes.end(), ax) != go.axes.end()) group = 2;
        plan.axes[ax] = PairAxis{dims[ax].len, b[ax], dims[ax].stride, dst_stride[ax], group, k - 1 - q};
    }
    p.in_mult = (uint32_t)gi.mult;
    p.out_mult = (uint32_t)go.mult;
    if (!gi.partial) p.in_mult = p.A;
    if (!go.partial) p.out_mult = p.B;
    if (n_boxes > 0x7fffffffLL) return plan;
    plan.n_boxes = n_boxes;
    p.split = (uint32_t)split;
    if (split == 2 && p.nIg % 2) return plan;
    p.nIgLoc = p.nIg / split;
    p.nJqLoc = (p.nJq + split - 1) / split;
    p.div_nJqLoc = FastDiv(p.nJqLoc);
    const size_t cells = (size_t)(p.A / split) * p.PB;  // per CTA
    p.st_offset = (uint32_t)(cells * 4);
    p.tab_offset = (uint32_t)(p.st_offset + ((cells + 15) & ~(size_t)15));
    plan.smem = p.tab_offset + ((size_t)p.A + p.B) * sizeof(uint32_t);
    const uint32_t mt_per_cta = split == 2 ? p.nIg * p.nJqLoc : p.nIg * p.nJq;
    plan.geometry = split == 1;
    if (plan.smem > 200 * 1024 || mt_per_cta > 3u * kPairThreads) return plan;
    if (split == 1 && plan.smem > 100 * 1024 && mt_per_cta <= (uint32_t)kPairThreads) return plan;
    plan.use = true;
    return plan;
}

// Measured on the 100^3 x 10^3 reversal (profiles/README.md): longer runs mean fewer, longer
// DRAM bursts.  A 200-cell OUTPUT run (800-byte value bursts, 200-byte status bursts) beats
// the square 100 x 100 tile by 17 %, a 200-cell input run by 9 %; the square tile is the
// fallback when a longer run would collide with the other group.
inline PairPlan transpose_pair_plan(const std::vector<GDim>& dims_in) {
    static const int64_t want_in = [] { const char* e = getenv("OLAP_PAIR_IN"); return e ? (int64_t)atoi(e) : (int64_t)-1; }();
    static const int64_t want_out = [] { const char* e = getenv("OLAP_PAIR_OUT"); return e ? (int64_t)atoi(e) : (int64_t)-1; }();
    if (want_in >= 0 || want_out >= 0) {  // tuning knobs
        static const int split_knob = [] { const char* e = getenv("OLAP_PAIR_SPLIT"); return e ? atoi(e) : 1; }();
        PairPlan plan = transpose_pair_plan_for(dims_in, std::max<int64_t>(0, want_in), std::max<int64_t>(0, want_out), split_knob);
        if (plan.use) return plan;
        return transpose_pair_plan_for(dims_in, 0, 0);
    }
    // OLAP_PAIR_CLUSTER=1: a 2-CTA cluster shares a 200 x 200 tile through distributed shared
    // memory (800-byte runs on BOTH sides; st.async into the peer's tile, counted on its
    // mbarrier).  Correct (tests/test_gpu_parity.py::test_cluster_transpose) but measured slower
    // than the 100 x 200 single-CTA tile on the 6-D reversal (3.06 ms vs 2.36 ms): the two
    // lock-step cluster rendezvous per tile cost more than the longer input bursts save.  Off by
    // default; read at every call so that a test can switch it on.
    const char* cluster_env = getenv("OLAP_PAIR_CLUSTER");
    const int cluster_knob = cluster_env ? atoi(cluster_env) : 0;
    PairPlan plan;
    if (cluster_knob) plan = transpose_pair_plan_for(dims_in, 200, 200, 2);
    if (!plan.use) plan = transpose_pair_plan_for(dims_in, 0, 200);
    if (!plan.use) plan = transpose_pair_plan_for(dims_in, 200, 0);
    if (!plan.use) plan = transpose_pair_plan_for(dims_in, 0, 0);
    return plan;
}

inline int launch_transpose_pair(const GatherMeasure* d_meas, const uint32_t* d_src_row, const uint32_t* d_dst_row,
                                 int n, PairPlan& plan) {
    plan.p.meas = d_meas;
    plan.p.src_row = d_src_row;
    plan.p.dst_row = d_dst_row;
    static bool attr_set = false;
    if (!attr_set) {
        OLAP_CUDA(cudaFuncSetAttribute(transpose_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        OLAP_CUDA(cudaFuncSetAttribute(transpose_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        OLAP_CUDA(cudaFuncSetAttribute(transpose_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (plan.p.split == 2) {
        static bool attr2_set = false;
        if (!attr2_set) {
            OLAP_CUDA(cudaFuncSetAttribute(transpose_pair2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            OLAP_CUDA(cudaFuncSetAttribute(transpose_pair2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            OLAP_CUDA(cudaFuncSetAttribute(transpose_pair2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr2_set = true;
        }
        const uint32_t mt_cta = plan.p.nIg * plan.p.nJqLoc;
        const int nm2 = (int)ceil_div(mt_cta, kPairThreads);
        static const int cl_per_sm = [] { const char* e = getenv("OLAP_PAIR_CTAS"); return e ? atoi(e) : 1; }();
        // clusters of 2 CTAs: one CTA per SM -> sm_count / 2 clusters per measure plane set
        const int64_t clusters = std::min<int64_t>(plan.n_boxes, std::max<int64_t>(1, ceil_div((int64_t)g.sm_count / 2 * cl_per_sm, n)));
        const unsigned threads2 = (unsigned)((ceil_div(mt_cta, nm2) + 31) / 32 * 32);
        const dim3 grid2((unsigned)(clusters * 2), (unsigned)n);
        mark_kernels_begin();
        if (nm2 == 1) transpose_pair2_kernel<1><<<grid2, threads2, plan.smem, g.stream>>>(plan.p, (uint32_t)plan.n_boxes);
        else if (nm2 == 2) transpose_pair2_kernel<2><<<grid2, threads2, plan.smem, g.stream>>>(plan.p, (uint32_t)plan.n_boxes);
        else transpose_pair2_kernel<3><<<grid2, threads2, plan.smem, g.stream>>>(plan.p, (uint32_t)plan.n_boxes);
        OLAP_CUDA(cudaGetLastError());
        ++g_launches;
        return OLAP_OK;
    }
    // persistent CTAs, each walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t n_mt = plan.p.nIg * plan.p.nJq;
    const int nm = (int)ceil_div(n_mt, kPairThreads);
    static const int per_sm_knob = [] { const char* e = getenv("OLAP_PAIR_CTAS"); return e ? atoi(e) : 1; }();
    const int per_sm = nm >= 2 ? 1 : per_sm_knob;
    const int64_t ctas = std::min<int64_t>(plan.n_boxes, std::max<int64_t>(1, ceil_div((int64_t)g.sm_count * per_sm, n)));
    const unsigned threads = (unsigned)((ceil_div(n_mt, nm) + 31) / 32 * 32);
    const dim3 grid((unsigned)ctas, (unsigned)n);
    mark_kernels_begin();
    if (nm == 1) transpose_pair_kernel<1><<<grid, threads, plan.smem, g.stream>>>(plan.p, (uint32_t)plan.n_boxes);
    else if (nm == 3) transpose_pair_kernel<3><<<grid, threads, plan.smem, g.stream>>>(plan.p, (uint32_t)plan.n_boxes);
    else transpose_pair_kernel<2><<<grid, threads, plan.smem, g.stream>>>(plan.p, (uint32_t)plan.n_boxes);
    ++g_launches;
    return OLAP_OK;
}

}  // namespace olap
