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
#include <algorithm>
#include <numeric>

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
    const UpMeasure* meas;                      // device table, or nullptr: use meas_inline
    UpMeasure meas_inline[kInlineMeasures];
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
    uint32_t merge_offset;  // byte offset of the split-merge scratch (G > 1)
    int32_t G, logG;        // threads sharing one output (power of two)
};

struct TileDecision {
    bool use = false;
    int R = 1;
    bool bulk_values = false, bulk_status = false;
    size_t smem = 0;
    uint32_t st_offset = 0, merge_offset = 0;
    int G = 1;
};

// Use the tile kernel when the inner run is short and at least one row fits in shared memory.
inline TileDecision tile_plan(int64_t O, int64_t C, int64_t P, int64_t I, bool any_status) {
    TileDecision t;
    static const int force = [] { const char* e = getenv("OLAP_TILE"); return e ? atoi(e) : -1; }();  // tuning knob
    if (force == 0) return t;
    if (I >= 32 && (I % 4 == 0 || (I % 2 == 0 && I >= 64) || I >= 128)) return t;  // the vectorised mid kernel streams these
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
    const int64_t mult_v = 4 / std::gcd<int64_t, int64_t>(row_in % 4 == 0 ? 4 : row_in % 4, 4);
    const int64_t mult_s = 16 / std::gcd<int64_t, int64_t>(row_in % 16 == 0 ? 16 : row_in % 16, 16);
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
    // few outputs per tile and long child lists: let G threads share an output
    const int64_t n_out = R * P * I, avg_children = std::max<int64_t>(1, C / std::max<int64_t>(P, 1));
    int G = 1;
    while (G < 256 && n_out * G * 2 <= 256 && avg_children >= 8 * G) G *= 2;
    t.G = G;
    if (G > 1) {
        t.merge_offset = (uint32_t)t.smem;
        t.smem += 256 * 16 + 256;
    }
    t.use = true;
    return t;
}

template <int METHOD, bool NANDEF, bool RANGE, int STATUS>
__device__ __forceinline__ void up_tile_reduce(const UpTileParams& p, const UpMeasure& m, const float* s_val,
                                               const uint8_t* s_st, int64_t o0, int rows, unsigned char* s_merge) {
    typedef Lane<METHOD, NANDEF> L;
    const int n_out = rows * p.row_out;
    float* out = m.out + o0 * p.row_out;
    uint8_t* st_out = STATUS ? m.st_out + o0 * p.row_out : nullptr;
    if (p.G > 1) {
        // few outputs per tile: G consecutive threads share one output, each reduces one
        // contiguous chunk of the children, thread 0 of the group folds the chunks in order
        L* s_lane = reinterpret_cast<L*>(s_merge);
        uint8_t* s_stm = s_merge + 256 * sizeof(L);
        const int j = threadIdx.x >> p.logG, gq = threadIdx.x & (p.G - 1);
        L lane;
        uint32_t st = 0, redo_base = 0;
        int32_t k0 = 0, k1 = 0;
        if (j < n_out) {
            const uint32_t r = p.div_row_out.div((uint32_t)j);
            const uint32_t q = (uint32_t)j - r * (uint32_t)p.row_out;
            const uint32_t pi = p.div_i.div(q);
            const uint32_t i = q - pi * (uint32_t)p.I;
            k0 = p.pstart[pi];
            k1 = p.pstart[pi + 1];
            const int32_t per = (k1 - k0 + p.G - 1) >> p.logG;
            const int32_t ks = min(k1, k0 + gq * per), ke = min(k1, ks + per);
            const uint32_t base = r * (uint32_t)p.row_in + i;
            redo_base = base;
#pragma unroll 4
            for (int32_t k = ks; k < ke; ++k) {
                const uint32_t c = RANGE ? (uint32_t)k : (uint32_t)p.children[k];
                const uint32_t idx = base + c * (uint32_t)p.I;
                const float v = s_val[idx];
                lane.step(v);
                if (STATUS == ST_LOAD) st |= s_st[idx];
                if (STATUS == ST_DERIVE) st |= present_f(v, NANDEF) ? OLAP_STATUS_SET : OLAP_STATUS_UNSET;
            }
        }
        s_lane[threadIdx.x] = lane;
        s_stm[threadIdx.x] = (uint8_t)st;
        __syncthreads();
        if (j < n_out && gq == 0) {
            for (int q2 = 1; q2 < p.G; ++q2) {
                lane.merge(s_lane[threadIdx.x + q2]);
                st |= s_stm[threadIdx.x + q2];
            }
            out[j] = lane_poisoned(lane) ? exact_redo<METHOD, RANGE>(s_val + redo_base, p.I, p.children, k0, k1) : lane.result();
            if (STATUS) st_out[j] = (uint8_t)(k0 == k1 ? OLAP_STATUS_UNSET : st);
        }
        return;
    }
    for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
        const uint32_t r = p.div_row_out.div((uint32_t)j);
        const uint32_t q = (uint32_t)j - r * (uint32_t)p.row_out;
        const uint32_t pi = p.div_i.div(q);
        const uint32_t i = q - pi * (uint32_t)p.I;
        const int32_t k0 = p.pstart[pi], k1 = p.pstart[pi + 1];
        const uint32_t base = r * (uint32_t)p.row_in + i;
        L lane;
        uint32_t st = 0;
#pragma unroll 4
        for (int32_t k = k0; k < k1; ++k) {
            const uint32_t c = RANGE ? (uint32_t)k : (uint32_t)p.children[k];
            const uint32_t idx = base + c * (uint32_t)p.I;
            const float v = s_val[idx];
            lane.step(v);
            if (STATUS == ST_LOAD) st |= s_st[idx];
            if (STATUS == ST_DERIVE) st |= present_f(v, NANDEF) ? OLAP_STATUS_SET : OLAP_STATUS_UNSET;
        }
        out[j] = lane_poisoned(lane) ? exact_redo<METHOD, RANGE>(s_val + base, p.I, p.children, k0, k1) : lane.result();
        if (STATUS) st_out[j] = (uint8_t)(k0 == k1 ? OLAP_STATUS_UNSET : st);
    }
}

template <bool NANDEF, bool RANGE, int STATUS>
__device__ __forceinline__ void up_tile_dispatch(const UpTileParams& p, const UpMeasure& m, const float* s_val,
                                                 const uint8_t* s_st, int64_t o0, int rows, unsigned char* s_merge) {
    switch (m.method) {
        case OLAP_SUM: up_tile_reduce<OLAP_SUM, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
        case OLAP_AVERAGE: up_tile_reduce<OLAP_AVERAGE, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
        case OLAP_HIGHEST: up_tile_reduce<OLAP_HIGHEST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
        case OLAP_LOWEST: up_tile_reduce<OLAP_LOWEST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
        case OLAP_FIRST: up_tile_reduce<OLAP_FIRST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
        case OLAP_LAST: up_tile_reduce<OLAP_LAST, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
        case OLAP_COUNT: up_tile_reduce<OLAP_COUNT, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
        default: up_tile_reduce<OLAP_PRODUCT, NANDEF, RANGE, STATUS>(p, m, s_val, s_st, o0, rows, s_merge); break;
    }
}

template <bool RANGE>
__global__ void __launch_bounds__(256, 8) drillup_tile_kernel(const __grid_constant__ UpTileParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* s_val = reinterpret_cast<float*>(smem);
    uint8_t* s_st = smem + p.st_offset;
    const UpMeasure m = p.meas ? p.meas[blockIdx.y] : p.meas_inline[blockIdx.y];
    // ST_DERIVE: the source's status plane follows from its values (olap_store::derived): it is neither copied
    // nor read, the bytes are recomputed from the staged cells
    const int st_mode = m.st_in ? ST_LOAD : (m.derive ? ST_DERIVE : ST_NONE);
    const bool status = st_mode == ST_LOAD;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.st_offset + (st_mode != ST_NONE ? (((size_t)p.R * p.row_in + 15) & ~(size_t)15) : 0));
    unsigned char* s_merge = smem + p.merge_offset;

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
        if (st_mode == ST_LOAD) up_tile_dispatch<true, RANGE, ST_LOAD>(p, m, s_val, s_st, o0, rows, s_merge);
        else if (st_mode == ST_DERIVE) up_tile_dispatch<true, RANGE, ST_DERIVE>(p, m, s_val, s_st, o0, rows, s_merge);
        else up_tile_dispatch<true, RANGE, ST_NONE>(p, m, s_val, s_st, o0, rows, s_merge);
    } else {
        if (st_mode == ST_LOAD) up_tile_dispatch<false, RANGE, ST_LOAD>(p, m, s_val, s_st, o0, rows, s_merge);
        else if (st_mode == ST_DERIVE) up_tile_dispatch<false, RANGE, ST_DERIVE>(p, m, s_val, s_st, o0, rows, s_merge);
        else up_tile_dispatch<false, RANGE, ST_NONE>(p, m, s_val, s_st, o0, rows, s_merge);
    }
}

inline int launch_up_tile(const UpMeasure* d_meas, const UpMeasure* h_meas, int n, bool contiguous,
                          const int32_t* d_pstart, const int32_t* d_children, int64_t O, int64_t C, int64_t P,
                          int64_t I, const TileDecision& t) {
    UpTileParams p{};
    p.meas = d_meas;
    if (!d_meas) for (int k = 0; k < n; ++k) p.meas_inline[k] = h_meas[k];
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
    p.merge_offset = t.merge_offset;
    p.G = t.G;
    p.logG = 0;
    while ((1 << p.logG) < t.G) ++p.logG;
    const int64_t tiles = ceil_div(O, t.R);
    if (tiles > 0x7fffffffLL) return fail(OLAP_E_UNSUPPORTED, "drillUp: grid too large (%lld tiles)", (long long)tiles);
    static bool attr_set[2] = {false, false};
    auto kern = contiguous ? drillup_tile_kernel<true> : drillup_tile_kernel<false>;
    if (!attr_set[contiguous]) {
        OLAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        attr_set[contiguous] = true;
    }
    mark_kernels_begin();
    kern<<<dim3((unsigned)tiles, (unsigned)n), 256, t.smem, g.stream>>>(p);
    ++g_launches;
    return OLAP_OK;
}

// ---- reorder whose innermost axis moves: box-tiled permutation -------------------------
// reorder (in-memory.js:178-211) is out[new] = in[old] for a permutation of the axes.
// When the output's innermost axis is not the input's, a thread-per-output gather reads
// with a large stride.  Here a CTA owns a BOX of the cube: a few axes with extent b_d > 1
// chosen so that the box covers >= 32 contiguous cells of the INPUT's trailing axes and
// >= 32 contiguous cells of the OUTPUT's trailing axes.  Phase 1 walks the box in input
// order (coalesced loads) and drops each cell at its output-order position in shared
// memory (row pitches padded to odd numbers against bank conflicts); phase 2 walks the box
// in output order: linear shared-memory reads, coalesced stores.  The status plane rides
// along as bytes.
constexpr int kMaxBoxDims = 4;
#ifndef OLAP_TRANSPOSE_MIN_BLOCKS
#define OLAP_TRANSPOSE_MIN_BLOCKS 6  // <= 42 registers: 6 CTAs per SM keep more loads in flight (8 spills too much: measured slower)
#endif

struct BoxDim {
    uint32_t b;         // box extent along this axis (1 = padding entry)
    FastDiv div;        // by b
    uint32_t g_stride;  // box-relative global stride (source in phase 1, destination in phase 2)
    uint32_t s_stride;  // shared-memory pitch, elements
    uint32_t axis;      // output axis whose (possibly ragged) extent bounds this entry
    uint32_t ext_mult;  // cells of fully covered inner axes merged into this entry
    // entry 0 only: the run may span TWO globally contiguous axes that are not contiguous in
    // shared memory: position -> (pos / by) * s_outer + (pos % by) * s_stride
    uint32_t by;        // cells of the inner axis of the run (== b when the run is one axis)
    FastDiv div_by;
    uint32_t s_outer;   // shared-memory pitch of the outer axis of the run
};

struct TransposeParams {
    const GatherMeasure* meas;
    int nb;                        // entries used in rd[] / wr[] (both padded to nb)
    BoxDim rd[kMaxBoxDims], wr[kMaxBoxDims];  // input order / output order, innermost first
    uint32_t box_cells;            // prod b
    uint32_t runs_in, runs_out;    // prod of b over entries 1.. of rd / wr
    // box-relative {global, shared} offsets of every run, precomputed on the host: identical
    // for every full box, so the kernel's per-cell work is a table read plus the position
    const uint2* rd_tab;
    const uint2* wr_tab;
    // grid decomposition: every output axis, outermost first
    int n_axes;
    uint32_t boxes[OLAP_MAX_DIMS];   // number of boxes along the axis
    FastDiv div_boxes[OLAP_MAX_DIMS];
    uint32_t len[OLAP_MAX_DIMS];     // axis length
    uint32_t bsize[OLAP_MAX_DIMS];   // box extent (1 for axes outside the box)
    int64_t src_stride[OLAP_MAX_DIMS], dst_stride[OLAP_MAX_DIMS];
    uint32_t st_offset;              // byte offset of the status tile in shared memory
    uint32_t tab_offset;             // byte offset of the staged run tables in shared memory
    int rd_vec4;                     // phase 1 may use 128-bit loads along the input run
    int wr_vec4;                     // phase 2 may use 128-bit stores along the output run
};

struct TransposePlan {
    bool use = false;
    TransposeParams p{};
    int64_t n_boxes = 0;
    size_t smem = 0;
    std::vector<uint2> rd_tab, wr_tab;  // host copies, uploaded by the caller
};

// Decode the index of a RUN (everything but entry 0) into global / shared offsets.
template <int NB, bool CHECK>
__device__ __forceinline__ bool run_decode(const BoxDim (&dims)[kMaxBoxDims], const uint32_t (&ext)[kMaxBoxDims],
                                           uint32_t t, uint32_t& g_off, uint32_t& s_off) {
    bool ok = true;
    g_off = 0;
    s_off = 0;
#pragma unroll
    for (int d = 1; d < NB; ++d) {
        uint32_t c;
        if (d == NB - 1) c = t;
        else {
            const uint32_t q = dims[d].div.div(t);
            c = t - q * dims[d].b;
            t = q;
        }
        if (CHECK) ok = ok && c < ext[d];
        g_off += c * dims[d].g_stride;
        s_off += c * dims[d].s_stride;
    }
    return ok;
}

// Walk the box one RUN at a time (entry 0 of the order is contiguous in global memory).
// Long runs are cut into 128-cell chunks, one warp each; runs shorter than a warp are
// packed several to a warp.  Each lane resolves up to 4 cells per pass: `offs` receives
// their global / shared-memory offsets, the return value is the bit mask of valid slots.
struct CellOffs {
    uint32_t g[4], s[4];
};

template <int NB, bool CHECK>
struct RunWalker {
    const BoxDim (&dims)[kMaxBoxDims];
    const uint32_t (&ext)[kMaxBoxDims];
    uint32_t n_runs, len, warp, lane;
    bool long_runs;
    uint32_t chunks, tasks;            // long runs
    uint32_t w_log, per_pass, passes;  // short runs
    __device__ __forceinline__ RunWalker(const BoxDim (&d)[kMaxBoxDims], const uint32_t (&e)[kMaxBoxDims], uint32_t runs)
        : dims(d), ext(e), n_runs(runs) {
        warp = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        len = CHECK ? ext[0] : dims[0].b;
        long_runs = dims[0].b >= 32;
        chunks = (dims[0].b + 127) >> 7;
        tasks = n_runs * chunks;
        w_log = 0;
        while ((1u << w_log) < dims[0].b) ++w_log;
        per_pass = 32u >> w_log;
        passes = (n_runs + per_pass - 1) / per_pass;
    }
    __device__ __forceinline__ uint32_t s_pos(uint32_t pos) const {
        const uint32_t hi = dims[0].div_by.div(pos);
        return hi * dims[0].s_outer + (pos - hi * dims[0].by) * dims[0].s_stride;
    }
    // number of outer iterations this warp performs
    __device__ __forceinline__ uint32_t iterations() const {
        const uint32_t total = long_runs ? tasks : (passes + 3) / 4;
        return total > warp ? (total - warp + 7) / 8 : 0;
    }
    __device__ __forceinline__ uint32_t resolve(uint32_t it, CellOffs& o) const {
        uint32_t mask = 0;
        const uint32_t idx = warp + it * 8;
        if (long_runs) {
            const uint32_t run = idx / chunks, chunk = idx - run * chunks;
            uint32_t g_off, s_off;
            if (!run_decode<NB, CHECK>(dims, ext, run, g_off, s_off)) return 0;
            const uint32_t p0 = (chunk << 7) + lane;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t pos = p0 + u * 32;
                if (pos < len) {
                    o.g[u] = g_off + pos * dims[0].g_stride;
                    o.s[u] = s_off + s_pos(pos);
                    mask |= 1u << u;
                }
            }
        } else {
            const uint32_t sub = lane >> w_log, pos = lane & ((1u << w_log) - 1u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t run = (idx * 4 + u) * per_pass + sub;
                uint32_t g_off, s_off;
                if (run < n_runs && pos < len && run_decode<NB, CHECK>(dims, ext, run, g_off, s_off)) {
                    o.g[u] = g_off + pos * dims[0].g_stride;
                    o.s[u] = s_off + s_pos(pos);
                    mask |= 1u << u;
                }
            }
        }
        return mask;
    }
};

// Full boxes (the common case): run offsets come from the host-built tables.  128 lane
// slots per warp pass; runs shorter than 128 cells are packed 128/W to a pass (W = run
// length rounded up to a power of two), longer ones are cut into 128-cell chunks.
struct RunGeom {
    uint32_t len, w_log, per_pass, passes, chunks;
    bool packed;
};
__device__ __forceinline__ RunGeom run_geom(const BoxDim& e0, uint32_t n_runs, uint32_t unit = 1) {
    RunGeom r;
    r.len = e0.b / unit;  // positions are counted in units of `unit` cells
    r.packed = r.len < 128;
    r.w_log = 0;
    while ((1u << r.w_log) < r.len && r.w_log < 7) ++r.w_log;
    r.per_pass = 128u >> r.w_log;
    r.chunks = (r.len + 127) >> 7;
    r.passes = r.packed ? (n_runs + r.per_pass - 1) / r.per_pass : n_runs * r.chunks;
    return r;
}
// slot in [0,128) of pass `pass` -> (run, pos); false when the slot is padding
__device__ __forceinline__ bool slot_cell(const RunGeom& r, uint32_t n_runs, uint32_t pass, uint32_t slot,
                                          uint32_t& run, uint32_t& pos) {
    if (r.packed) {
        run = pass * r.per_pass + (slot >> r.w_log);
        pos = slot & ((1u << r.w_log) - 1u);
        return run < n_runs && pos < r.len;
    }
    run = pass / r.chunks;
    pos = (pass - run * r.chunks) * 128u + slot;
    return pos < r.len;
}
__device__ __forceinline__ uint32_t run_spos(const BoxDim& e0, uint32_t pos) {
    if (e0.by == e0.b) return pos * e0.s_stride;
    const uint32_t hi = e0.div_by.div(pos);
    return hi * e0.s_outer + (pos - hi * e0.by) * e0.s_stride;
}

// STATUS: ST_NONE, ST_LOAD (the bytes ride through the tile) or ST_DERIVE (the source plane follows from its values,
// olap_store::derived: nothing is loaded or staged, the bytes are recomputed from the cells on their way out)
__device__ __forceinline__ uint32_t box_status_of(float v, int nan_default) {
    return present_f(v, nan_default) ? (uint32_t)OLAP_STATUS_SET : (uint32_t)OLAP_STATUS_UNSET;
}

template <int STATUS>
__device__ __forceinline__ void transpose_full_box(const TransposeParams& p, const float* __restrict__ src,
                                                   const uint8_t* __restrict__ st_src, float* __restrict__ dst,
                                                   uint8_t* __restrict__ st_dst, float* s_val, uint8_t* s_st,
                                                   const uint2* s_rd, const uint2* s_wr, int nan_default) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.rd_vec4) {
        // 128-bit loads along the input run (4 cells per lane slot, 4 slots in flight)
        const RunGeom r = run_geom(p.rd[0], p.runs_in, 4);
        for (uint32_t pass = warp; pass < r.passes; pass += 8) {
            float4 v[4];
            uint32_t sb[4], so[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t run, pos;
                ok[u] = slot_cell(r, p.runs_in, pass, u * 32 + lane, run, pos);
                if (ok[u]) {
                    const uint2 t = s_rd[run];
                    const uint32_t g = t.x + pos * 4;
                    so[u] = t.y + pos * 4 * p.rd[0].s_stride;
                    v[u] = ld_stream4(src + g);
                    if (STATUS == ST_LOAD) sb[u] = ld_stream_u32(st_src + g);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) {
                    const uint32_t ss = p.rd[0].s_stride;
                    s_val[so[u]] = v[u].x; s_val[so[u] + ss] = v[u].y;
                    s_val[so[u] + 2 * ss] = v[u].z; s_val[so[u] + 3 * ss] = v[u].w;
                    if (STATUS == ST_LOAD) {
                        s_st[so[u]] = (uint8_t)sb[u]; s_st[so[u] + ss] = (uint8_t)(sb[u] >> 8);
                        s_st[so[u] + 2 * ss] = (uint8_t)(sb[u] >> 16); s_st[so[u] + 3 * ss] = (uint8_t)(sb[u] >> 24);
                    }
                }
        }
    } else {
        const RunGeom r = run_geom(p.rd[0], p.runs_in);
        const uint32_t gs = p.rd[0].g_stride;
        for (uint32_t pass = warp; pass < r.passes; pass += 8) {
            float v[4];
            uint8_t sb[4];
            uint32_t so[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t run, pos;
                ok[u] = slot_cell(r, p.runs_in, pass, u * 32 + lane, run, pos);
                if (ok[u]) {
                    const uint2 t = s_rd[run];
                    const uint32_t g = t.x + pos * gs;
                    so[u] = t.y + run_spos(p.rd[0], pos);
                    v[u] = ld_stream1(src + g);
                    if (STATUS == ST_LOAD) sb[u] = st_src[g];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) {
                    s_val[so[u]] = v[u];
                    if (STATUS == ST_LOAD) s_st[so[u]] = sb[u];
                }
        }
    }
    __syncthreads();
    if (p.wr_vec4) {
        // 128-bit stores along the output run: a lane gathers 4 consecutive output cells (and
        // their status bytes) from the tile and writes them with one store each
        const RunGeom r = run_geom(p.wr[0], p.runs_out, 4);
        const bool simple = p.wr[0].by == p.wr[0].b;
        const uint32_t ss = p.wr[0].s_stride;
        for (uint32_t pass = warp; pass < r.passes; pass += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t run, pos;
                if (slot_cell(r, p.runs_out, pass, u * 32 + lane, run, pos)) {
                    const uint2 t = s_wr[run];
                    const uint32_t g = t.x + pos * 4;
                    uint32_t so[4];
                    if (simple) {
                        so[0] = t.y + pos * 4 * ss;
                        so[1] = so[0] + ss; so[2] = so[1] + ss; so[3] = so[2] + ss;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) so[k] = t.y + run_spos(p.wr[0], pos * 4 + k);
                    }
                    const float4 q = make_float4(s_val[so[0]], s_val[so[1]], s_val[so[2]], s_val[so[3]]);
                    st_stream4(dst + g, q);
                    if (STATUS == ST_LOAD)
                        *reinterpret_cast<uint32_t*>(st_dst + g) = (uint32_t)s_st[so[0]] | ((uint32_t)s_st[so[1]] << 8) |
                                                                   ((uint32_t)s_st[so[2]] << 16) | ((uint32_t)s_st[so[3]] << 24);
                    if (STATUS == ST_DERIVE)
                        *reinterpret_cast<uint32_t*>(st_dst + g) = box_status_of(q.x, nan_default) | (box_status_of(q.y, nan_default) << 8) |
                                                                   (box_status_of(q.z, nan_default) << 16) | (box_status_of(q.w, nan_default) << 24);
                }
            }
        }
    } else {
        const RunGeom r = run_geom(p.wr[0], p.runs_out);
        const uint32_t gs = p.wr[0].g_stride;
        for (uint32_t pass = warp; pass < r.passes; pass += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t run, pos;
                if (slot_cell(r, p.runs_out, pass, u * 32 + lane, run, pos)) {
                    const uint2 t = s_wr[run];
                    const uint32_t g = t.x + pos * gs, so = t.y + run_spos(p.wr[0], pos);
                    const float q = s_val[so];
                    dst[g] = q;
                    if (STATUS == ST_LOAD) st_dst[g] = s_st[so];
                    if (STATUS == ST_DERIVE) st_dst[g] = (uint8_t)box_status_of(q, nan_default);
                }
            }
        }
    }
}

template <int NB, int STATUS, bool CHECK>
__device__ __forceinline__ void transpose_phases(const TransposeParams& p, const float* __restrict__ src,
                                                 const uint8_t* __restrict__ st_src, float* __restrict__ dst,
                                                 uint8_t* __restrict__ st_dst, float* s_val, uint8_t* s_st,
                                                 const uint32_t (&ext_rd)[kMaxBoxDims],
                                                 const uint32_t (&ext_wr)[kMaxBoxDims], int nan_default) {
    // ---- phase 1: input runs -> shared memory (4 loads in flight per lane)
    {
        const RunWalker<NB, CHECK> w(p.rd, ext_rd, p.runs_in);
        const uint32_t n_it = w.iterations();
        for (uint32_t it = 0; it < n_it; ++it) {
            CellOffs o;
            const uint32_t mask = w.resolve(it, o);
            float v[4];
            uint8_t sb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (mask & (1u << u)) {
                    v[u] = __ldg(src + o.g[u]);
                    if (STATUS == ST_LOAD) sb[u] = __ldg(st_src + o.g[u]);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (mask & (1u << u)) {
                    s_val[o.s[u]] = v[u];
                    if (STATUS == ST_LOAD) s_st[o.s[u]] = sb[u];
                }
        }
    }
    __syncthreads();
    // ---- phase 2: shared memory -> output runs
    {
        const RunWalker<NB, CHECK> w(p.wr, ext_wr, p.runs_out);
        const uint32_t n_it = w.iterations();
        for (uint32_t it = 0; it < n_it; ++it) {
            CellOffs o;
            const uint32_t mask = w.resolve(it, o);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (mask & (1u << u)) {
                    const float q = s_val[o.s[u]];
                    dst[o.g[u]] = q;
                    if (STATUS == ST_LOAD) st_dst[o.g[u]] = s_st[o.s[u]];
                    if (STATUS == ST_DERIVE) st_dst[o.g[u]] = (uint8_t)box_status_of(q, nan_default);
                }
        }
    }
}

// MINB: resident CTAs per SM the register budget is cut for.  Measured on the config-3 cube (r02): moving a loaded
// status plane wants 5 (48 registers: swap inner 2.28 ms against 2.57 ms), deriving it wants 6 (2.24 against 2.35 ms).
template <int NB, int MINB>
__global__ void __launch_bounds__(256, MINB) transpose_kernel(const __grid_constant__ TransposeParams p) {
    extern __shared__ __align__(16) unsigned char smem_t[];
    __shared__ uint32_t s_ext[OLAP_MAX_DIMS];
    __shared__ int64_t s_base[2];
    __shared__ int s_full;
    float* s_val = reinterpret_cast<float*>(smem_t);
    uint8_t* s_st = smem_t + p.st_offset;
    const GatherMeasure m = p.meas[blockIdx.y];
    // ---- which box am I, and how much of it is inside the cube
    if (threadIdx.x == 0) {
        uint32_t rest = blockIdx.x;
        int64_t src = 0, dst = 0;
        int full = 1;
        for (int a = p.n_axes - 1; a >= 0; --a) {
            const uint32_t q = p.div_boxes[a].div(rest);
            const uint32_t bi = rest - q * p.boxes[a];
            rest = q;
            const uint32_t start = bi * p.bsize[a];
            src += (int64_t)start * p.src_stride[a];
            dst += (int64_t)start * p.dst_stride[a];
            const uint32_t e = min(p.bsize[a], p.len[a] - start);
            s_ext[a] = e;
            full &= e == p.bsize[a];
        }
        s_base[0] = src;
        s_base[1] = dst;
        s_full = full;
    }
    __syncthreads();
    uint32_t ext_rd[kMaxBoxDims], ext_wr[kMaxBoxDims];
#pragma unroll
    for (int d = 0; d < NB; ++d) {
        ext_rd[d] = p.rd[d].b > 1 ? s_ext[p.rd[d].axis] * p.rd[d].ext_mult : 1u;
        ext_wr[d] = p.wr[d].b > 1 ? s_ext[p.wr[d].axis] * p.wr[d].ext_mult : 1u;
    }
    const float* src = m.in + s_base[0];
    float* dst = m.out + s_base[1];
    const int st_mode = m.st_in ? ST_LOAD : (m.derive && m.st_out ? ST_DERIVE : ST_NONE);
    const uint8_t* st_src = m.st_in ? m.st_in + s_base[0] : nullptr;
    uint8_t* st_dst = st_mode != ST_NONE ? m.st_out + s_base[1] : nullptr;
    if (s_full) {
        // stage the (box independent) run tables next to the tile
        uint2* s_rd = reinterpret_cast<uint2*>(smem_t + p.tab_offset);
        uint2* s_wr = s_rd + p.runs_in;
        for (uint32_t i = threadIdx.x; i < p.runs_in; i += 256) s_rd[i] = __ldg(p.rd_tab + i);
        for (uint32_t i = threadIdx.x; i < p.runs_out; i += 256) s_wr[i] = __ldg(p.wr_tab + i);
        __syncthreads();
        if (st_mode == ST_LOAD) transpose_full_box<ST_LOAD>(p, src, st_src, dst, st_dst, s_val, s_st, s_rd, s_wr, m.nan_default);
        else if (st_mode == ST_DERIVE) transpose_full_box<ST_DERIVE>(p, src, st_src, dst, st_dst, s_val, s_st, s_rd, s_wr, m.nan_default);
        else transpose_full_box<ST_NONE>(p, src, st_src, dst, st_dst, s_val, s_st, s_rd, s_wr, m.nan_default);
    } else {
        if (st_mode == ST_LOAD) transpose_phases<NB, ST_LOAD, true>(p, src, st_src, dst, st_dst, s_val, s_st, ext_rd, ext_wr, m.nan_default);
        else if (st_mode == ST_DERIVE) transpose_phases<NB, ST_DERIVE, true>(p, src, st_src, dst, st_dst, s_val, s_st, ext_rd, ext_wr, m.nan_default);
        else transpose_phases<NB, ST_NONE, true>(p, src, st_src, dst, st_dst, s_val, s_st, ext_rd, ext_wr, m.nan_default);
    }
}

// `dims`: the output axes (outermost first) as linear GDims with their SOURCE strides.
inline TransposePlan transpose_plan(const std::vector<GDim>& dims_in) {
    TransposePlan plan;
    static const int force = [] { const char* e = getenv("OLAP_TRANSPOSE"); return e ? atoi(e) : -1; }();
    if (force == 0) return plan;
    // drop unit axes and merge neighbours that stay adjacent (same rule as the gather path)
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
    if (dims.back().stride == 1) return plan;  // innermost axis unchanged: the vectorised gather streams it
    for (const GDim& d : dims)
        if (d.len > 0x7fffffffLL) return plan;
    std::vector<int64_t> dst_stride(k);
    int64_t acc = 1;
    for (int i = k - 1; i >= 0; --i) { dst_stride[i] = acc; acc *= dims[i].len; }
    // axes by ascending source stride = the input's trailing axes first
    std::vector<int> by_src(k);
    for (int i = 0; i < k; ++i) by_src[i] = i;
    std::sort(by_src.begin(), by_src.end(), [&](int a, int b) { return dims[a].stride < dims[b].stride; });
    std::vector<int64_t> b(k, 1);
    const int64_t run_target = 64, cells_target = 4096, cells_max = 6144;
    auto cells = [&] { int64_t c = 1; for (int i = 0; i < k; ++i) c *= b[i]; return c; };
    // an extent near `want` that splits the axis into equal parts (no ragged edge boxes)
    auto even_extent = [&](int64_t len, int64_t want) {
        want = std::max<int64_t>(1, std::min(want, len));
        const int64_t parts = std::max<int64_t>(1, len / want);
        return ceil_div(len, parts);
    };
    // (1) cover the input's trailing axes until a contiguous input run of >= run_target cells
    int64_t run = 1;
    for (int idx : by_src) {
        if (run >= run_target) break;
        b[idx] = std::max(b[idx], even_extent(dims[idx].len, ceil_div(run_target, run)));
        run *= b[idx];
        if (b[idx] < dims[idx].len) break;  // a partially covered axis ends the contiguous run
    }
    // (2) same for the output's trailing axes
    run = 1;
    for (int i = k - 1; i >= 0; --i) {
        if (run >= run_target) break;
        b[i] = std::max(b[i], even_extent(dims[i].len, ceil_div(run_target, run)));
        run *= b[i];
        if (b[i] < dims[i].len) break;
    }
    // largest even-split extent strictly below `cur`
    auto shrink = [&](int64_t len, int64_t cur) {
        int64_t parts = ceil_div(len, cur) + 1, extent = ceil_div(len, parts);
        while (extent >= cur && parts < len) extent = ceil_div(len, ++parts);
        return std::max<int64_t>(1, std::min(extent, cur - 1));
    };
    while (cells() > cells_max) {  // shrink the largest extent until the box fits
        int big = 0;
        for (int i = 1; i < k; ++i) if (b[i] > b[big]) big = i;
        if (b[big] <= 1) return plan;
        b[big] = shrink(dims[big].len, b[big]);
    }
    // (3) grow towards the target: widen axes already in the box, innermost output axes first
    for (int pass = 0; pass < 2 && cells() < cells_target; ++pass)
        for (int i = k - 1; i >= 0 && cells() < cells_target; --i) {
            if (pass == 0 && b[i] == 1) continue;
            if (b[i] >= dims[i].len) continue;
            const int64_t room = cells_max / (cells() / b[i]);
            // even split with an extent of AT MOST `room`
            const int64_t want = std::max<int64_t>(1, std::min<int64_t>(dims[i].len, room));
            const int64_t cand = ceil_div(dims[i].len, ceil_div(dims[i].len, want));
            if (cand > b[i] && cells() / b[i] * cand <= cells_max) b[i] = cand;
        }
    std::vector<int> box_axes;
    for (int i = 0; i < k; ++i) if (b[i] > 1) box_axes.push_back(i);
    if (box_axes.empty()) return plan;

    TransposeParams& p = plan.p;
    // shared-memory layout: INPUT order inside the box (ascending source stride), so that a
    // whole contiguous input run is also contiguous in shared memory; the pitch is padded to
    // an odd number wherever the next axis is not a continuation of the run, which makes the
    // strided shared-memory reads of phase 2 bank-conflict free.
    std::vector<uint32_t> pitch(k, 0);
    uint32_t sacc = 1;
    for (size_t q = 0; q < by_src.size(); ++q) {
        const int i = by_src[q];
        if (b[i] <= 1) continue;
        pitch[i] = sacc;
        sacc *= (uint32_t)b[i];
        // continuation = next box axis in input order is source-contiguous with this fully covered one
        bool continues = false;
        for (size_t q2 = q + 1; q2 < by_src.size(); ++q2) {
            const int j = by_src[q2];
            if (b[j] <= 1) continue;
            continues = b[i] == dims[i].len && dims[j].stride == dims[i].stride * dims[i].len;
            break;
        }
        if (!continues && sacc % 2 == 0) sacc += 1;
    }
    const size_t s_cells = sacc;
    // phase-1 order: ascending source stride; phase-2 order: ascending destination stride.
    // Neighbouring entries that are fully covered and contiguous (globally and in shared
    // memory) are merged into one, which is what keeps the decode at 2-3 divisions.
    struct Ent { int64_t b, g, s; int axis; bool whole; int64_t mult; };
    auto build = [&](bool read) {
        std::vector<int> order = box_axes;
        std::sort(order.begin(), order.end(), [&](int x, int y) {
            return read ? dims[x].stride < dims[y].stride : dst_stride[x] < dst_stride[y];
        });
        std::vector<Ent> ents;
        for (int i : order) {
            Ent e{b[i], read ? dims[i].stride : dst_stride[i], pitch[i], i, b[i] == dims[i].len, 1};
            if (!ents.empty()) {
                Ent& in = ents.back();
                // `in` fully covered (no ragged edge), and e continues it in both address spaces
                if (in.whole && e.g == in.g * in.b && e.s == in.s * in.b) {
                    in.mult = in.b;  // valid cells of the merged entry: mult * extent(outer axis)
                    in.b *= e.b;
                    in.whole = e.whole;
                    in.axis = e.axis;
                    continue;
                }
            }
            ents.push_back(e);
        }
        return ents;
    };
    std::vector<Ent> rd = build(true), wr = build(false);
    // compound run: let entry 0 also swallow entry 1 when the two are contiguous in GLOBAL
    // memory (they are not in shared memory, or build() would have merged them already)
    struct Run { int64_t by, s_outer; };
    auto compound = [&](std::vector<Ent>& ents) {
        Run r{ents[0].b, 0};
        if (ents.size() >= 2 && ents[0].whole && ents[0].mult == 1 && ents[1].mult == 1 &&
            ents[1].g == ents[0].g * ents[0].b && ents[0].b * ents[1].b <= 0x7fffffffLL) {
            r.by = ents[0].b;
            r.s_outer = ents[1].s;
            ents[0].mult = ents[0].b;  // extent of the run = by * extent(outer axis)
            ents[0].b *= ents[1].b;
            ents[0].axis = ents[1].axis;
            ents[0].whole = ents[1].whole;
            ents.erase(ents.begin() + 1);
        }
        return r;
    };
    const Run run_rd = compound(rd), run_wr = compound(wr);
    const int nb = (int)std::max(rd.size(), wr.size());
    if (nb > kMaxBoxDims) return plan;
    auto fill = [&](const std::vector<Ent>& ents, const Run& run, BoxDim* out) {
        for (int q = 0; q < kMaxBoxDims; ++q) {
            if (q < (int)ents.size()) {
                if (ents[q].g * (ents[q].b - 1) > 0x7fffffffLL) return false;
                out[q] = BoxDim{(uint32_t)ents[q].b, FastDiv((uint32_t)ents[q].b), (uint32_t)ents[q].g,
                                (uint32_t)ents[q].s, (uint32_t)ents[q].axis, (uint32_t)ents[q].mult,
                                (uint32_t)ents[q].b, FastDiv((uint32_t)ents[q].b), 0u};
            } else out[q] = BoxDim{1u, FastDiv(1u), 0u, 0u, 0u, 1u, 1u, FastDiv(1u), 0u};
        }
        out[0].by = (uint32_t)run.by;
        out[0].div_by = FastDiv((uint32_t)run.by);
        out[0].s_outer = (uint32_t)run.s_outer;
        return true;
    };
    if (!fill(rd, run_rd, p.rd) || !fill(wr, run_wr, p.wr)) return plan;
    p.nb = std::max(nb, 2);
    p.box_cells = (uint32_t)cells();
    p.runs_in = p.runs_out = 1;
    for (int q = 1; q < kMaxBoxDims; ++q) { p.runs_in *= p.rd[q].b; p.runs_out *= p.wr[q].b; }
    auto table = [&](const BoxDim* e, uint32_t runs, std::vector<uint2>& tab) {
        tab.resize(runs);
        for (uint32_t run = 0; run < runs; ++run) {
            uint32_t t = run, g_off = 0, s_off = 0;
            for (int d = 1; d < kMaxBoxDims; ++d) {
                const uint32_t c = t % e[d].b;
                t /= e[d].b;
                g_off += c * e[d].g_stride;
                s_off += c * e[d].s_stride;
            }
            tab[run] = make_uint2(g_off, s_off);
        }
    };
    table(p.rd, p.runs_in, plan.rd_tab);
    table(p.wr, p.runs_out, plan.wr_tab);
    // 128-bit loads along the input run: the run is a single axis group contiguous in global
    // memory, a multiple of 4 cells long, and every other source stride keeps 16-byte alignment
    p.rd_vec4 = p.rd[0].g_stride == 1 && p.rd[0].b % 4 == 0 && p.rd[0].by == p.rd[0].b;
    for (int i = 0; i < k; ++i)
        if (dims[i].stride != 1 && (dims[i].stride % 4 != 0)) {
            // an axis outside the run with an unaligned stride breaks alignment unless it is
            // part of the run itself (then its stride is covered by the run's contiguity)
            bool in_run = b[i] > 1 && dims[i].stride < p.rd[0].b;
            if (!in_run) p.rd_vec4 = 0;
        }
    // 128-bit stores along the output run: same conditions on the destination side (the run may
    // be a compound one, its shared-memory positions are resolved cell by cell)
    static const bool wr_vec_knob = [] { const char* e = getenv("OLAP_BOX_WR_VEC4"); return !e || atoi(e) != 0; }();
    p.wr_vec4 = wr_vec_knob && p.wr[0].g_stride == 1 && p.wr[0].b % 4 == 0;
    for (int i = 0; i < k; ++i)
        if (dst_stride[i] != 1 && (dst_stride[i] % 4 != 0)) {
            bool in_run = b[i] > 1 && dst_stride[i] < p.wr[0].b;
            if (!in_run) p.wr_vec4 = 0;
        }
    p.n_axes = k;
    // grid slots in traversal order (the LAST slot varies fastest): boxes that are neighbours in
    // the input's memory order run back to back, then neighbours in the output's order, so runs
    // that begin or end inside a DRAM atom share it through L2 (OLAP_BOX_ORDER=0: output order)
    static const bool neighbour_order = [] { const char* e = getenv("OLAP_BOX_ORDER"); return !e || atoi(e) != 0; }();
    std::vector<int> order;  // fastest first
    if (neighbour_order) {
        for (int ax : by_src) if (b[ax] < dims[ax].len) { order.push_back(ax); break; }
        for (int ax = k - 1; ax >= 0; --ax)
            if (b[ax] < dims[ax].len && std::find(order.begin(), order.end(), ax) == order.end()) { order.push_back(ax); break; }
    }
    for (int ax = k - 1; ax >= 0; --ax) if (std::find(order.begin(), order.end(), ax) == order.end()) order.push_back(ax);
    std::vector<uint32_t> slot_of(k);
    int64_t n_boxes = 1;
    for (int q = 0; q < k; ++q) {
        const int i = order[q], slot = k - 1 - q;
        slot_of[i] = (uint32_t)slot;
        p.len[slot] = (uint32_t)dims[i].len;
        p.bsize[slot] = (uint32_t)b[i];
        p.boxes[slot] = (uint32_t)ceil_div(dims[i].len, b[i]);
        p.div_boxes[slot] = FastDiv(p.boxes[slot]);
        p.src_stride[slot] = dims[i].stride;
        p.dst_stride[slot] = dst_stride[i];
        n_boxes *= p.boxes[slot];
    }
    for (size_t q = 0; q < rd.size(); ++q) p.rd[q].axis = slot_of[p.rd[q].axis];
    for (size_t q = 0; q < wr.size(); ++q) p.wr[q].axis = slot_of[p.wr[q].axis];
    if (n_boxes > 0x7fffffffLL) return plan;
    plan.n_boxes = n_boxes;
    p.st_offset = (uint32_t)((s_cells * 4 + 15) & ~(size_t)15);
    p.tab_offset = (uint32_t)(p.st_offset + ((s_cells + 15) & ~(size_t)15));
    plan.smem = p.tab_offset + ((size_t)p.runs_in + p.runs_out) * sizeof(uint2);
    if (plan.smem > 200 * 1024) return plan;
    plan.use = true;
    return plan;
}

inline int launch_transpose(bool derive_all, const GatherMeasure* d_meas, const uint2* d_rd_tab, const uint2* d_wr_tab, int n,
                            TransposePlan& plan) {
    plan.p.meas = d_meas;
    plan.p.rd_tab = d_rd_tab;
    plan.p.wr_tab = d_wr_tab;
    static bool attr_set = false;
    if (!attr_set) {
#define OLAP_T_ATTR(NB, MB) OLAP_CUDA(cudaFuncSetAttribute(transpose_kernel<NB, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
        OLAP_T_ATTR(2, 5); OLAP_T_ATTR(3, 5); OLAP_T_ATTR(4, 5); OLAP_T_ATTR(2, 6); OLAP_T_ATTR(3, 6); OLAP_T_ATTR(4, 6);
#undef OLAP_T_ATTR
        attr_set = true;
    }
    const dim3 grid((unsigned)plan.n_boxes, (unsigned)n);
    mark_kernels_begin();
    switch (plan.p.nb) {
        case 2: if (derive_all) transpose_kernel<2, 6><<<grid, 256, plan.smem, g.stream>>>(plan.p); else transpose_kernel<2, 5><<<grid, 256, plan.smem, g.stream>>>(plan.p); break;
        case 3: if (derive_all) transpose_kernel<3, 6><<<grid, 256, plan.smem, g.stream>>>(plan.p); else transpose_kernel<3, 5><<<grid, 256, plan.smem, g.stream>>>(plan.p); break;
        default: if (derive_all) transpose_kernel<4, 6><<<grid, 256, plan.smem, g.stream>>>(plan.p); else transpose_kernel<4, 5><<<grid, 256, plan.smem, g.stream>>>(plan.p); break;
    }
    ++g_launches;
    return OLAP_OK;
}

}  // namespace olap
