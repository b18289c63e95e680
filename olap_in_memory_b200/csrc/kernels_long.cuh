// drillup_long_kernel — drillUp (in-memory.js:265-334) of a LONG axis with a short inner run
// and few parents: [O, C, I] -> [O, P, I] with I < 32, C*I too large for one shared-memory
// tile and P*I <= 1024 outputs per row (rollups to 'all', customers -> segment, collapse of a
// whole cube, ...).  The output-driven kernels have nothing to parallelise over here (a
// thread per output would leave the chip idle), so the ROW is cut instead:
//
//   * a CTA owns a super-segment = T consecutive segments of Cs children of one row; each
//     segment (one contiguous span of Cs*I cells, values + status) is staged into shared
//     memory with cp.async.bulk on an mbarrier, like drillup_tile_kernel;
//   * per segment, the G threads that share an output reduce consecutive chunks of the
//     parent's children that fall into the segment (host-built table seg_ptr[s][p] = first
//     CSR position of parent p at or after child s*Cs) and the chunk states are folded IN
//     ORDER into the CTA's private accumulators in shared memory (Lane<METHOD>::merge, so
//     first/last and the restart rule stay exact);
//   * with more than one super-segment per row the private accumulators go to a scratch
//     array and drillup_long_merge_kernel folds them in row order.
// Deterministic, no atomics; every input byte is read once.
#pragma once

#include "kernels_tile.cuh"

namespace olap {

struct UpLongParams {
    const UpMeasure* meas;
    UpMeasure meas_inline[kInlineMeasures];
    const int32_t* pstart;    // [P+1]
    const int32_t* children;  // CSR (unused when RANGE)
    const int32_t* seg_ptr;   // RANGE: [(S_row+1) * P] CSR positions; else [S_row * (P+1)] positions in the segment's own list
    const uint16_t* perm16;   // !RANGE: [S_row * Cs] cells of every segment (offset from its first child), grouped by parent, ascending
    int64_t O;
    int32_t C, P, I;
    int32_t Cs, S_row, T, SS;  // children per segment, segments per row, segments per CTA, CTAs per row
    int32_t row_out;           // P * I
    FastDiv div_i, div_ss;
    uint32_t buf_stride, st_offset, perm_offset, bar_offset, state_offset, merge_offset;  // dynamic shared memory layout (bytes)
    int32_t G, logG;
    unsigned char* scratch;    // SS > 1: per measure [O*SS*row_out] 16-byte lane states, then as many status bytes
    int64_t scratch_stride;    // bytes per measure
    const uint8_t* map8;       // drillup_lanes_kernel (kernels_lanes.cuh): per-tile child lists
    int32_t lanes_cs;          // drillup_lanes_kernel: children per CTA segment
    uint32_t lanes_stage;      // drillup_lanes_kernel: bytes of one staging buffer of one warp
};

struct LongDecision {
    bool use = false;
    int32_t Cs = 0, S_row = 0, T = 0, SS = 0, G = 1;
    size_t smem = 0;
    uint32_t buf_stride = 0, st_offset = 0, perm_offset = 0, bar_offset = 0, state_offset = 0, merge_offset = 0;
    int64_t scratch_stride = 0;
};

constexpr int kLongStateBytes = 16;

// Called after tile_plan declined: rows too long for one tile, inner run short.
inline LongDecision long_plan(int64_t O, int64_t C, int64_t P, int64_t I, bool any_status, int n_meas, int sm_count,
                              bool contiguous) {
    LongDecision d;
    static const int force = [] { const char* e = getenv("OLAP_LONG"); return e ? atoi(e) : -1; }();
    if (force == 0) return d;
    if (I >= 32 && (I % 4 == 0 || (I % 2 == 0 && I >= 64) || I >= 128)) return d;  // mid kernel territory
    const int64_t row_out = P * I;
    if (row_out > 1024 || C > 0x7fffffffLL || O > 0x3fffffffLL) return d;
    if (C / std::max<int64_t>(P, 1) < 64) return d;  // short child lists: a thread per output is fine
    const int64_t per_cell = any_status ? 5 : 4;
    static const int64_t budget = [] { const char* e = getenv("OLAP_LONG_KB"); return (int64_t)(e ? atoi(e) : 32) * 1024; }();
    // children per segment: Cs*I cells fill the budget (one of TWO staging buffers) and are a multiple of 16 cells
    const int64_t per_child = per_cell * I + (contiguous ? 0 : 2);
    int64_t Cs = std::max<int64_t>(1, budget / per_child);
    int64_t unit = 16 / std::gcd<int64_t, int64_t>(I, 16);  // Cs % unit == 0  =>  (Cs*I) % 16 == 0
    if (!contiguous) unit = std::lcm<int64_t, int64_t>(unit, 8);    // ... and the 2-byte cell list of a segment is a 16-byte multiple
    Cs = std::max(unit, Cs / unit * unit);
    if (Cs * per_child > 64 * 1024 || (!contiguous && Cs * I > 65535)) return d;
    Cs = std::min<int64_t>(Cs, ceil_div(C, unit) * unit);
    const int64_t S_row = ceil_div(C, Cs);
    if ((S_row + 1) * (P + 1) > (int64_t)(8 << 20)) return d;  // seg_ptr table
    // CTAs per row: fill the chip a few times over
    const int64_t target = (int64_t)sm_count * 8;
    int64_t SS = std::min<int64_t>(S_row, std::max<int64_t>(1, ceil_div(target, std::max<int64_t>(1, O * n_meas))));
    const int64_t T = ceil_div(S_row, SS);
    SS = ceil_div(S_row, T);
    if (O * SS > 0x7fffffffLL) return d;
    d.Cs = (int32_t)Cs; d.S_row = (int32_t)S_row; d.T = (int32_t)T; d.SS = (int32_t)SS;
    const size_t cells = (size_t)Cs * I;
    d.st_offset = (uint32_t)(cells * 4);          // inside a buffer
    d.perm_offset = (uint32_t)(cells * per_cell); // cells % 16 == 0: 16-byte aligned
    d.buf_stride = d.perm_offset + (uint32_t)(contiguous ? 0 : Cs * 2);
    d.bar_offset = 2 * d.buf_stride;
    d.state_offset = d.bar_offset + 16;
    d.merge_offset = d.state_offset + (uint32_t)(((size_t)row_out * (kLongStateBytes + 1) + 15) & ~(size_t)15);
    int G = 1;
    while (G < 256 && row_out * G * 2 <= 256) G *= 2;
    d.G = G;
    d.smem = d.merge_offset + (G > 1 ? 256 * (kLongStateBytes + 1) : 0);
    if (d.smem > 160 * 1024) return d;
    if (SS > 1) {
        d.scratch_stride = (O * SS * row_out * (kLongStateBytes + 1) + 255) & ~(int64_t)255;
        if (d.scratch_stride * n_meas > ((int64_t)1 << 30)) return d;
    }
    d.use = true;
    return d;
}

template <int METHOD, bool NANDEF, bool RANGE, bool STATUS>
__device__ __forceinline__ void up_long_segment(const UpLongParams& p, const float* s_val, const uint8_t* s_st,
                                                const uint16_t* s_perm, int32_t s, int32_t c0,
                                                Lane<METHOD, NANDEF>* s_state, uint8_t* s_stacc, unsigned char* s_merge,
                                                int32_t pre_k0, int32_t pre_k1) {
    typedef Lane<METHOD, NANDEF> L;
    // RANGE: CSR positions == child indices; else positions in the segment's staged cell list
    const int32_t* lo = RANGE ? p.seg_ptr + (size_t)s * p.P : p.seg_ptr + (size_t)s * (p.P + 1);
    const int32_t* hi = RANGE ? lo + p.P : lo + 1;
    if (p.G > 1) {
        L* s_lane = reinterpret_cast<L*>(s_merge);
        uint8_t* s_stm = s_merge + 256 * kLongStateBytes;
        const int j = threadIdx.x >> p.logG, gq = threadIdx.x & (p.G - 1);
        L lane;
        uint32_t st = 0;
        if (j < p.row_out) {
            const uint32_t pi = p.div_i.div((uint32_t)j), i = (uint32_t)j - pi * (uint32_t)p.I;
            const int32_t k0 = pre_k0, k1 = pre_k1;
            // odd chunk length (in cells, or in 16-byte quads on the vector path): the G threads
            // of a group start in different shared-memory banks
            const bool vec = RANGE && p.I == 1;
            int32_t per = (k1 - k0 + p.G - 1) >> p.logG;
            per = vec ? ((((per + 3) >> 2) | 1) << 2) : (per | 1);
            const int32_t ks = min(k1, k0 + gq * per);
            int32_t ke = min(k1, ks + per);
            if (vec) {
                // children of a contiguous map are consecutive cells of the tile: 128-bit
                // shared-memory loads (and 32-bit loads of four status bytes) between the
                // scalar head / tail that reach 16-byte alignment
                int32_t k = ks;
                for (; k < ke && ((k - c0) & 3); ++k) {
                    lane.step(s_val[k - c0]);
                    if (STATUS) st |= s_st[k - c0];
                }
#pragma unroll 2
                for (; k + 4 <= ke; k += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(s_val + (k - c0));
                    lane.step(v.x); lane.step(v.y); lane.step(v.z); lane.step(v.w);
                    if (STATUS) {
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(s_st + (k - c0));
                        st |= w | (w >> 8) | (w >> 16) | (w >> 24);
                    }
                }
                for (; k < ke; ++k) {
                    lane.step(s_val[k - c0]);
                    if (STATUS) st |= s_st[k - c0];
                }
                ke = ks;  // nothing left for the scalar loop
            }
#pragma unroll 4
            for (int32_t k = ks; k < ke; ++k) {
                const uint32_t c = RANGE ? (uint32_t)k - (uint32_t)c0 : (uint32_t)s_perm[k];
                const uint32_t idx = c * (uint32_t)p.I + i;
                lane.step(s_val[idx]);
                if (STATUS) st |= s_st[idx];
            }
        }
        s_lane[threadIdx.x] = lane;
        s_stm[threadIdx.x] = (uint8_t)(st & 0xffu);
        __syncthreads();
        // ordered tree fold of the G chunk states (left operand = earlier children)
        for (int h = 1; h < p.G; h <<= 1) {
            if ((gq & (2 * h - 1)) == 0 && j < p.row_out) {
                L a = s_lane[threadIdx.x];
                a.merge(s_lane[threadIdx.x + h]);
                s_lane[threadIdx.x] = a;
                s_stm[threadIdx.x] |= s_stm[threadIdx.x + h];
            }
            __syncthreads();
        }
        if (j < p.row_out && gq == 0) {
            L acc = s_state[j];
            acc.merge(s_lane[threadIdx.x]);
            s_state[j] = acc;
            s_stacc[j] |= s_stm[threadIdx.x];
        }
        return;
    }
    for (int j = threadIdx.x; j < p.row_out; j += blockDim.x) {
        const uint32_t pi = p.div_i.div((uint32_t)j), i = (uint32_t)j - pi * (uint32_t)p.I;
        const int32_t k0 = lo[pi], k1 = hi[pi];
        L lane;
        uint32_t st = 0;
#pragma unroll 4
        for (int32_t k = k0; k < k1; ++k) {
            const uint32_t c = RANGE ? (uint32_t)k - (uint32_t)c0 : (uint32_t)s_perm[k];
            const uint32_t idx = c * (uint32_t)p.I + i;
            lane.step(s_val[idx]);
            if (STATUS) st |= s_st[idx];
        }
        L acc = s_state[j];
        acc.merge(lane);
        s_state[j] = acc;
        if (STATUS) s_stacc[j] |= (uint8_t)st;
    }
}

// One output cell from its folded state (shared by the single-CTA-per-row path and the merge kernel).
template <int METHOD, bool NANDEF, bool RANGE>
__device__ __forceinline__ void up_long_finish(const UpLongParams& p, const UpMeasure& m, int64_t o, int j,
                                               const Lane<METHOD, NANDEF>& lane, uint32_t st) {
    const uint32_t pi = p.div_i.div((uint32_t)j), i = (uint32_t)j - pi * (uint32_t)p.I;
    const int32_t k0 = p.pstart[pi], k1 = p.pstart[pi + 1];
    const int64_t out_off = o * p.row_out + j;
    m.out[out_off] = lane_poisoned(lane)
                         ? exact_redo<METHOD, RANGE>(m.in + o * (int64_t)p.C * p.I + i, p.I, p.children, k0, k1)
                         : lane.result();
    if (m.st_out) m.st_out[out_off] = (uint8_t)(k0 == k1 ? OLAP_STATUS_UNSET : st);
}

template <int METHOD, bool NANDEF, bool RANGE, bool STATUS>
__device__ __forceinline__ void up_long_body(const UpLongParams& p, const UpMeasure& m, unsigned char* smem,
                                             int64_t o, int32_t ss) {
    typedef Lane<METHOD, NANDEF> L;
    static_assert(sizeof(L) <= kLongStateBytes, "lane state larger than its slot");
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.bar_offset);  // two barriers, one per staging buffer
    L* s_state = reinterpret_cast<L*>(smem + p.state_offset);
    uint8_t* s_stacc = smem + p.state_offset + (size_t)p.row_out * kLongStateBytes;
    unsigned char* s_merge = smem + p.merge_offset;
    for (int j = threadIdx.x; j < p.row_out; j += blockDim.x) {
        s_state[j] = L();
        s_stacc[j] = 0;
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    __syncthreads();
    const int32_t s_begin = ss * p.T, s_end = min(p.S_row, s_begin + p.T);
    // stage segment s into buffer b: bulk copy for the 16-byte multiples when the span starts
    // on a 16-byte boundary, plain loads for the rest
    auto issue = [&](int32_t s, uint32_t b) {
        const int32_t c0 = s * p.Cs, c1 = min(p.C, c0 + p.Cs);
        const int64_t start = (o * (int64_t)p.C + c0) * p.I;
        const uint32_t n_cells = (uint32_t)(c1 - c0) * (uint32_t)p.I;
        float* s_val = reinterpret_cast<float*>(smem + b * p.buf_stride);
        uint8_t* s_st = smem + b * p.buf_stride + p.st_offset;
        const float* g_val = m.in + start;
        const uint8_t* g_st = STATUS ? m.st_in + start : nullptr;
        const uint32_t bulk_v = (start & 3) == 0 ? ((n_cells * 4u) & ~15u) : 0u;
        const uint32_t bulk_s = (STATUS && (start & 15) == 0) ? (n_cells & ~15u) : 0u;
        const uint32_t bulk_p = RANGE ? 0u : (uint32_t)p.Cs * 2u;  // the list is padded to whole segments
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar + b, bulk_v + bulk_s + bulk_p);
            if (bulk_v) bulk_g2s(s_val, g_val, bulk_v, bar + b);
            if (bulk_s) bulk_g2s(s_st, g_st, bulk_s, bar + b);
            if (bulk_p) bulk_g2s(smem + b * p.buf_stride + p.perm_offset, p.perm16 + (size_t)s * p.Cs, bulk_p, bar + b);
        }
        for (uint32_t q = (bulk_v >> 2) + threadIdx.x; q < n_cells; q += blockDim.x) s_val[q] = ld_stream1(g_val + q);
        if (STATUS)
            for (uint32_t q = bulk_s + threadIdx.x; q < n_cells; q += blockDim.x) s_st[q] = g_st[q];
    };
    uint32_t parity[2] = {0u, 0u};
    if (s_begin < s_end) issue(s_begin, 0u);
    for (int32_t s = s_begin; s < s_end; ++s) {
        const uint32_t b = (uint32_t)(s - s_begin) & 1u;
        if (s + 1 < s_end) issue(s + 1, b ^ 1u);  // the other buffer was released by the barrier below
        // this thread's child range inside the segment: fetched while the copy is in flight
        int32_t pre_k0 = 0, pre_k1 = 0;
        if (p.G > 1) {
            const int j = threadIdx.x >> p.logG;
            if (j < p.row_out) {
                const uint32_t pi = p.div_i.div((uint32_t)j);
                pre_k0 = __ldg(p.seg_ptr + (RANGE ? (size_t)s * p.P + pi : (size_t)s * (p.P + 1) + pi));
                pre_k1 = __ldg(p.seg_ptr + (RANGE ? (size_t)(s + 1) * p.P + pi : (size_t)s * (p.P + 1) + pi + 1));
            }
        }
        mbar_wait(bar + b, parity[b]);
        parity[b] ^= 1u;
        __syncthreads();
        up_long_segment<METHOD, NANDEF, RANGE, STATUS>(p, reinterpret_cast<const float*>(smem + b * p.buf_stride),
                                                       smem + b * p.buf_stride + p.st_offset,
                                                       reinterpret_cast<const uint16_t*>(smem + b * p.buf_stride + p.perm_offset),
                                                       s, s * p.Cs, s_state, s_stacc, s_merge, pre_k0, pre_k1);
        __syncthreads();  // buffer b and the chunk states are free again
    }
    if (p.SS == 1) {
        for (int j = threadIdx.x; j < p.row_out; j += blockDim.x)
            up_long_finish<METHOD, NANDEF, RANGE>(p, m, o, j, s_state[j], s_stacc[j]);
    } else {
        unsigned char* base = p.scratch + (size_t)blockIdx.y * p.scratch_stride;
        const int64_t n_states = p.O * p.SS * p.row_out;
        const int64_t slot0 = (o * p.SS + ss) * p.row_out;
        for (int j = threadIdx.x; j < p.row_out; j += blockDim.x) {
            *reinterpret_cast<L*>(base + (slot0 + j) * kLongStateBytes) = s_state[j];
            base[n_states * kLongStateBytes + slot0 + j] = s_stacc[j];
        }
    }
}

template <bool NANDEF, bool RANGE, bool STATUS>
__device__ __forceinline__ void up_long_dispatch(const UpLongParams& p, const UpMeasure& m, unsigned char* smem,
                                                 int64_t o, int32_t ss) {
    switch (m.method) {
        case OLAP_SUM: up_long_body<OLAP_SUM, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
        case OLAP_AVERAGE: up_long_body<OLAP_AVERAGE, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
        case OLAP_HIGHEST: up_long_body<OLAP_HIGHEST, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
        case OLAP_LOWEST: up_long_body<OLAP_LOWEST, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
        case OLAP_FIRST: up_long_body<OLAP_FIRST, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
        case OLAP_LAST: up_long_body<OLAP_LAST, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
        case OLAP_COUNT: up_long_body<OLAP_COUNT, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
        default: up_long_body<OLAP_PRODUCT, NANDEF, RANGE, STATUS>(p, m, smem, o, ss); break;
    }
}

template <bool RANGE>
__global__ void __launch_bounds__(256) drillup_long_kernel(const __grid_constant__ UpLongParams p) {
    extern __shared__ __align__(128) unsigned char smem_l[];
    const UpMeasure m = p.meas ? p.meas[blockIdx.y] : p.meas_inline[blockIdx.y];
    const uint32_t o = p.div_ss.div(blockIdx.x), ss = blockIdx.x - o * (uint32_t)p.SS;
    const bool status = m.st_in != nullptr;
    if (m.nan_default) {
        if (status) up_long_dispatch<true, RANGE, true>(p, m, smem_l, o, (int32_t)ss);
        else up_long_dispatch<true, RANGE, false>(p, m, smem_l, o, (int32_t)ss);
    } else {
        if (status) up_long_dispatch<false, RANGE, true>(p, m, smem_l, o, (int32_t)ss);
        else up_long_dispatch<false, RANGE, false>(p, m, smem_l, o, (int32_t)ss);
    }
}

constexpr int kLongMergeThreads = 64;

// One CTA per output cell: thread t folds a contiguous run of the SS partial states, the runs
// are folded in order through shared memory.
template <int METHOD, bool NANDEF, bool RANGE>
__device__ __forceinline__ void up_long_merge_body(const UpLongParams& p, const UpMeasure& m, int64_t o, int j,
                                                   unsigned char* s_raw) {
    typedef Lane<METHOD, NANDEF> L;
    L* s_lane = reinterpret_cast<L*>(s_raw);
    uint8_t* s_stm = s_raw + kLongMergeThreads * kLongStateBytes;
    const unsigned char* base = p.scratch + (size_t)blockIdx.y * p.scratch_stride;
    const int64_t n_states = p.O * p.SS * p.row_out;
    const int32_t per = (p.SS + kLongMergeThreads - 1) / kLongMergeThreads;
    const int32_t s0 = min(p.SS, (int32_t)threadIdx.x * per), s1 = min(p.SS, s0 + per);
    L acc;
    uint32_t st = 0;
    for (int32_t ss = s0; ss < s1; ++ss) {
        const int64_t slot = (o * p.SS + ss) * p.row_out + j;
        acc.merge(*reinterpret_cast<const L*>(base + slot * kLongStateBytes));
        st |= base[n_states * kLongStateBytes + slot];
    }
    s_lane[threadIdx.x] = acc;
    s_stm[threadIdx.x] = (uint8_t)st;
    __syncthreads();
    for (int h = 1; h < kLongMergeThreads; h <<= 1) {
        if ((threadIdx.x & (2 * h - 1)) == 0) {
            L a = s_lane[threadIdx.x];
            a.merge(s_lane[threadIdx.x + h]);
            s_lane[threadIdx.x] = a;
            s_stm[threadIdx.x] |= s_stm[threadIdx.x + h];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) up_long_finish<METHOD, NANDEF, RANGE>(p, m, o, j, s_lane[0], s_stm[0]);
}

template <bool RANGE>
__global__ void __launch_bounds__(kLongMergeThreads) drillup_long_merge_kernel(const __grid_constant__ UpLongParams p) {
    __shared__ __align__(16) unsigned char s_raw[kLongMergeThreads * (kLongStateBytes + 1)];
    const UpMeasure m = p.meas ? p.meas[blockIdx.y] : p.meas_inline[blockIdx.y];
    const int64_t t = blockIdx.x;
    const int64_t o = t / p.row_out;
    const int j = (int)(t - o * p.row_out);
#define OLAP_LONG_MERGE(M)                                                          \
    if (m.nan_default) up_long_merge_body<M, true, RANGE>(p, m, o, j, s_raw);       \
    else up_long_merge_body<M, false, RANGE>(p, m, o, j, s_raw);                    \
    break;
    switch (m.method) {
        case OLAP_SUM: OLAP_LONG_MERGE(OLAP_SUM)
        case OLAP_AVERAGE: OLAP_LONG_MERGE(OLAP_AVERAGE)
        case OLAP_HIGHEST: OLAP_LONG_MERGE(OLAP_HIGHEST)
        case OLAP_LOWEST: OLAP_LONG_MERGE(OLAP_LOWEST)
        case OLAP_FIRST: OLAP_LONG_MERGE(OLAP_FIRST)
        case OLAP_LAST: OLAP_LONG_MERGE(OLAP_LAST)
        case OLAP_COUNT: OLAP_LONG_MERGE(OLAP_COUNT)
        default: OLAP_LONG_MERGE(OLAP_PRODUCT)
    }
#undef OLAP_LONG_MERGE
}

// Contiguous map: seg_ptr[s * P + p] = first CSR position k of parent p with child >= s * Cs.
// Any other map: seg_ptr[s * (P+1) + p] = where parent p starts in segment s's own cell list
// (entry P = its length), and perm16[s * Cs + q] = the q-th cell of that list as an offset from
// the segment's first child: parents in order, children ascending within a parent.
struct LongTables {
    std::vector<int32_t> seg_ptr;
    std::vector<uint16_t> perm16;
};
inline LongTables long_seg_table(const std::vector<int32_t>& pstart, const std::vector<int32_t>& children,
                                 bool contiguous, int64_t C, int64_t P, const LongDecision& d) {
    LongTables t;
    if (contiguous) {
        t.seg_ptr.resize((size_t)(d.S_row + 1) * P);
        for (int64_t s = 0; s <= d.S_row; ++s) {
            const int64_t bound = std::min<int64_t>(C, s * (int64_t)d.Cs);
            for (int64_t q = 0; q < P; ++q)
                t.seg_ptr[(size_t)s * P + q] = (int32_t)std::min<int64_t>(pstart[q + 1], std::max<int64_t>(pstart[q], bound));
        }
        return t;
    }
    t.seg_ptr.assign((size_t)d.S_row * (P + 1), 0);
    t.perm16.assign((size_t)d.S_row * d.Cs, 0);
    std::vector<int32_t> cursor(pstart.begin(), pstart.end() - 1);  // next unread CSR position per parent
    for (int64_t s = 0; s < d.S_row; ++s) {
        const int64_t c0 = s * (int64_t)d.Cs, c1 = std::min<int64_t>(C, c0 + d.Cs);
        int32_t q = 0;
        for (int64_t par = 0; par < P; ++par) {
            t.seg_ptr[(size_t)s * (P + 1) + par] = q;
            int32_t k = cursor[par];
            while (k < pstart[par + 1] && children[k] < c1) t.perm16[(size_t)c0 + q++] = (uint16_t)(children[k++] - c0);
            cursor[par] = k;
        }
        t.seg_ptr[(size_t)s * (P + 1) + P] = q;
    }
    return t;
}

inline int launch_up_long(const UpMeasure* d_meas, const UpMeasure* h_meas, int n, bool contiguous,
                          const int32_t* d_pstart, const int32_t* d_children, const int32_t* d_seg_ptr,
                          const uint16_t* d_perm16, int64_t O,
                          int64_t C, int64_t P, int64_t I, const LongDecision& d, unsigned char* d_scratch) {
    UpLongParams p{};
    p.meas = d_meas;
    if (!d_meas) for (int k = 0; k < n; ++k) p.meas_inline[k] = h_meas[k];
    p.pstart = d_pstart;
    p.children = d_children;
    p.seg_ptr = d_seg_ptr;
    p.perm16 = d_perm16;
    p.O = O; p.C = (int32_t)C; p.P = (int32_t)P; p.I = (int32_t)I;
    p.Cs = d.Cs; p.S_row = d.S_row; p.T = d.T; p.SS = d.SS;
    p.row_out = (int32_t)(P * I);
    p.div_i = FastDiv((uint32_t)I);
    p.div_ss = FastDiv((uint32_t)d.SS);
    p.buf_stride = d.buf_stride; p.st_offset = d.st_offset; p.perm_offset = d.perm_offset; p.bar_offset = d.bar_offset; p.state_offset = d.state_offset; p.merge_offset = d.merge_offset;
    p.G = d.G;
    p.logG = 0;
    while ((1 << p.logG) < d.G) ++p.logG;
    p.scratch = d_scratch;
    p.scratch_stride = d.scratch_stride;
    static bool attr_set[2] = {false, false};
    auto kern = contiguous ? drillup_long_kernel<true> : drillup_long_kernel<false>;
    if (!attr_set[contiguous]) {
        OLAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set[contiguous] = true;
    }
    mark_kernels_begin();
    kern<<<dim3((unsigned)(O * d.SS), (unsigned)n), 256, d.smem, g.stream>>>(p);
    ++g_launches;
    if (d.SS > 1) {
        const int64_t blocks = O * p.row_out;  // one CTA per output cell
        if (blocks > 0x7fffffffLL) return fail(OLAP_E_UNSUPPORTED, "drillUp: grid too large");
        if (contiguous) drillup_long_merge_kernel<true><<<dim3((unsigned)blocks, (unsigned)n), kLongMergeThreads, 0, g.stream>>>(p);
        else drillup_long_merge_kernel<false><<<dim3((unsigned)blocks, (unsigned)n), kLongMergeThreads, 0, g.stream>>>(p);
        ++g_launches;
    }
    return OLAP_OK;
}

}  // namespace olap
