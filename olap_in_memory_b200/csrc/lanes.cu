// drillup_lanes_kernel (kernels_lanes.cuh) in a translation unit of its own: eight methods x two defaults x three
// status modes x unrolled parents make it the slowest kernel of the library to compile.
#include "kernels_long.cuh"
#include "kernels_lanes.cuh"

namespace olap {

static_assert(kLanesStateBytes == kLongStateBytes, "lanes states use the slots of the long kernel's merge pass");

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// status byte of a list of `n` cells of which `cnt` are set, when the plane follows from the values
__device__ __forceinline__ uint32_t derived_status(uint32_t cnt, uint32_t n) {
    return (cnt ? OLAP_STATUS_SET : 0u) | (cnt < n ? OLAP_STATUS_UNSET : 0u);
}

template <typename G, int METHOD, bool NANDEF, bool RANGE, int STATUS>
__device__ __forceinline__ void up_lanes_body(const UpLongParams& p, const UpMeasure& m, unsigned char* smem, uint32_t rg,
                                              int32_t ss) {
    typedef Lane<METHOD, NANDEF> L;
    static_assert(sizeof(L) <= kLongStateBytes, "lane state larger than its slot");
    constexpr int TILE = G::kTile, STAGES = G::kStages, PER_LANE = TILE / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)rg * kLanesRows;
    const int rows = (int)min((int64_t)kLanesRows, p.O - row0);
    // my warp's chunk of the CTA's segment (both start on tile boundaries)
    const int32_t chunk = p.lanes_cs / kLanesWarps;
    const int32_t c_begin = min(p.C, ss * p.lanes_cs + warp * chunk), c_end = min(p.C, c_begin + chunk);
    const int32_t n_tiles = (c_end - c_begin + TILE - 1) / TILE;
    unsigned char* my = smem + (size_t)warp * STAGES * p.lanes_stage;
    const char* g_row = reinterpret_cast<const char*>(m.in + row0 * (int64_t)p.C);
    const uint8_t* g_st = STATUS == ST_LOAD ? m.st_in + row0 * (int64_t)p.C : nullptr;
    const int64_t row_bytes = (int64_t)p.C * 4;

    // ranks of my children of a tile (children lane, lane + 32, ...): PER_LANE bytes of the table
    auto ranks_of = [&](int32_t t) -> uint32_t {
        const uint8_t* tab = p.map8 + (size_t)((c_begin + t * TILE) / TILE) * G::kTableBytes + G::kListBytes;
        if (PER_LANE == 4) return __ldg(reinterpret_cast<const uint32_t*>(tab) + lane);
        if (PER_LANE == 2) return __ldg(reinterpret_cast<const uint16_t*>(tab) + lane);
        return __ldg(tab + lane);
    };
    auto issue = [&](int32_t t, uint32_t ranks) {
        const int32_t c0 = c_begin + t * TILE;
        const int32_t n = min(TILE, c_end - c0);
        const uint32_t sb = smem_u32(my + (size_t)(t % STAGES) * p.lanes_stage);
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const int cc = lane + 32 * j;
            if (cc < n) {
                const char* src = g_row + (int64_t)(c0 + cc) * 4;
                const uint32_t dst = sb + 4u * ((ranks >> (8 * j)) & 0xffu);
                if (rows == kLanesRows) {
#pragma unroll
                    for (int r = 0; r < kLanesRows; ++r) cp_async4(dst + r * G::kPitchV, src + r * row_bytes);
                } else {
                    for (int r = 0; r < rows; ++r) cp_async4(dst + r * G::kPitchV, src + r * row_bytes);
                }
            }
        }
        if (STATUS == ST_LOAD) {  // C % 4 == 0: whole words of 4 status bytes, in child order
            constexpr int WORDS = TILE / 4, ROWS_PER_PASS = 32 / WORDS;  // 16 words: two rows per pass; 32 words: one
            const int w = lane % WORDS;
            if (4 * w < n) {
                const uint8_t* s_src = g_st + c0 + 4 * w;
                const uint32_t s_dst = sb + G::kValBytes + G::kListBytes + 4u * w;
                for (int r = lane / WORDS; r < rows; r += ROWS_PER_PASS) cp_async4(s_dst + r * G::kPitchS, s_src + (int64_t)r * p.C);
            }
        }
        // the tile's lists: positions (children grouped by parent) + 16 bounds
        if (lane < (int)(G::kListBytes / 16)) cp_async16(sb + G::kValBytes + 16u * lane, p.map8 + (size_t)(c0 / TILE) * G::kTableBytes + 16 * lane);
    };

    L part[kLanesMaxP];
    uint32_t pst[kLanesMaxP];
#pragma unroll
    for (int q = 0; q < kLanesMaxP; ++q) pst[q] = 0;

    // one loop, one copy site and one fold site: iterations -STAGES .. -1 only fill the pipeline
    uint32_t next_ranks = n_tiles > 0 ? ranks_of(0) : 0u;
    for (int32_t t = -STAGES; t < n_tiles; ++t) {
        const int32_t tn = t + STAGES;
        const uint32_t ranks = next_ranks;
        if (tn + 1 < n_tiles) next_ranks = ranks_of(tn + 1);  // in flight while this tile is folded
        if (t >= 0) {
            cp_async_wait<STAGES - 1>();
            __syncwarp();
            const unsigned char* sb = my + (size_t)(t % STAGES) * p.lanes_stage;
            const float* sv = reinterpret_cast<const float*>(sb + (size_t)lane * G::kPitchV);
            const uint8_t* s_perm = sb + G::kValBytes;
            const uint8_t* sst = sb + G::kValBytes + G::kListBytes + (size_t)lane * G::kPitchS;
            const uint4 ow = *reinterpret_cast<const uint4*>(s_perm + TILE);  // list bounds 0 .. 15
            const uint32_t owords[4] = {ow.x, ow.y, ow.z, ow.w};
#pragma unroll
            for (int q = 0; q < kLanesMaxP; ++q) {
                const int k0 = (int)((owords[q >> 2] >> (8 * (q & 3))) & 0xffu);
                const int k1 = (int)((owords[(q + 1) >> 2] >> (8 * ((q + 1) & 3))) & 0xffu);
                if (q < p.P && k0 < k1) {  // warp-uniform
                    const int h = (k1 - k0 + 1) >> 1, n2 = k1 - k0 - h;  // chain a0: h columns from k0, chain a1: n2 <= h columns from k0 + h
                    const float* pa = sv + k0;
                    const float* pb = sv + k0 + h;
                    L a0, a1;
                    uint32_t st = 0, cnt = 0;
                    int i = 0;
                    for (; i + 4 <= n2; i += 4) {
                        float va[4], vb[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { va[e] = pa[i + e]; vb[e] = pb[i + e]; }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (STATUS == ST_DERIVE) cnt += (present_f(va[e], NANDEF) ? 1u : 0u) + (present_f(vb[e], NANDEF) ? 1u : 0u);
                            a0.step(va[e]);
                            a1.step(vb[e]);
                        }
                    }
#pragma unroll 1
                    for (; i < n2; ++i) {
                        const float va = pa[i], vb = pb[i];
                        if (STATUS == ST_DERIVE) cnt += (present_f(va, NANDEF) ? 1u : 0u) + (present_f(vb, NANDEF) ? 1u : 0u);
                        a0.step(va);
                        a1.step(vb);
                    }
                    if (h > n2) {  // odd list: the first chain is one longer
                        const float va = pa[h - 1];
                        if (STATUS == ST_DERIVE) cnt += present_f(va, NANDEF) ? 1u : 0u;
                        a0.step(va);
                    }
                    if (STATUS == ST_DERIVE) st = derived_status(cnt, (uint32_t)(k1 - k0));
                    if (STATUS == ST_LOAD) {  // the plane sits in child order: through the list of the parent's children
#pragma unroll 4
                        for (int k = k0; k < k1; ++k) st |= sst[s_perm[k]];
                    }
                    a0.merge(a1);
                    part[q].merge(a0);
                    pst[q] |= st;
                }
            }
            __syncwarp();  // every lane is done with this stage: refill it
        }
        if (tn < n_tiles) issue(tn, ranks);
        cp_async_commit();
    }
    cp_async_wait<0>();
    __syncthreads();  // all warps are out of their pipelines: the staging memory now holds the warps' lanes
    uint8_t* s_pst = smem + (size_t)kLanesWarps * kLanesMaxP * 32 * kLongStateBytes;  // after the [warp][parent][row] states
#pragma unroll
    for (int q = 0; q < kLanesMaxP; ++q) {
        *reinterpret_cast<L*>(smem + (size_t)((warp * kLanesMaxP + q) * 32 + lane) * kLongStateBytes) = part[q];
        s_pst[(warp * kLanesMaxP + q) * 32 + lane] = (uint8_t)pst[q];
    }
    __syncthreads();
    for (int q = warp; q < p.P; q += kLanesWarps) {
        L acc;
        uint32_t acc_st = 0;
#pragma unroll
        for (int w2 = 0; w2 < kLanesWarps; ++w2) {  // earlier children first
            acc.merge(*reinterpret_cast<const L*>(smem + (size_t)((w2 * kLanesMaxP + q) * 32 + lane) * kLongStateBytes));
            acc_st |= s_pst[(w2 * kLanesMaxP + q) * 32 + lane];
        }
        if (lane < rows) {
            const int64_t o = row0 + lane;
            if (p.SS == 1) {
                up_long_finish<METHOD, NANDEF, RANGE>(p, m, o, q, acc, acc_st);
            } else {
                unsigned char* base = p.scratch + (size_t)blockIdx.y * p.scratch_stride;
                const int64_t n_states = p.O * p.SS * p.row_out;
                const int64_t slot = (o * p.SS + ss) * p.row_out + q;
                *reinterpret_cast<L*>(base + slot * kLongStateBytes) = acc;
                base[n_states * kLongStateBytes + slot] = (uint8_t)acc_st;
            }
        }
    }
}

template <typename G, bool NANDEF, bool RANGE, int STATUS>
__device__ __forceinline__ void up_lanes_dispatch(const UpLongParams& p, const UpMeasure& m, unsigned char* smem, uint32_t rg, int32_t ss) {
    switch (m.method) {
        case OLAP_SUM: up_lanes_body<G, OLAP_SUM, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
        case OLAP_AVERAGE: up_lanes_body<G, OLAP_AVERAGE, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
        case OLAP_HIGHEST: up_lanes_body<G, OLAP_HIGHEST, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
        case OLAP_LOWEST: up_lanes_body<G, OLAP_LOWEST, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
        case OLAP_FIRST: up_lanes_body<G, OLAP_FIRST, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
        case OLAP_LAST: up_lanes_body<G, OLAP_LAST, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
        case OLAP_COUNT: up_lanes_body<G, OLAP_COUNT, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
        default: up_lanes_body<G, OLAP_PRODUCT, NANDEF, RANGE, STATUS>(p, m, smem, rg, ss); break;
    }
}

template <typename G, bool RANGE>
__global__ void __launch_bounds__(kLanesWarps * 32, 3) drillup_lanes_kernel(const __grid_constant__ UpLongParams p) {
    extern __shared__ __align__(128) unsigned char smem_lanes[];
    const UpMeasure m = p.meas ? p.meas[blockIdx.y] : p.meas_inline[blockIdx.y];
    const uint32_t rg = p.div_ss.div(blockIdx.x), ss = blockIdx.x - rg * (uint32_t)p.SS;
    const int status = m.st_in ? ST_LOAD : (m.derive ? ST_DERIVE : ST_NONE);
    if (m.nan_default) {
        if (status == ST_LOAD) up_lanes_dispatch<G, true, RANGE, ST_LOAD>(p, m, smem_lanes, rg, (int32_t)ss);
        else if (status == ST_DERIVE) up_lanes_dispatch<G, true, RANGE, ST_DERIVE>(p, m, smem_lanes, rg, (int32_t)ss);
        else up_lanes_dispatch<G, true, RANGE, ST_NONE>(p, m, smem_lanes, rg, (int32_t)ss);
    } else {
        if (status == ST_LOAD) up_lanes_dispatch<G, false, RANGE, ST_LOAD>(p, m, smem_lanes, rg, (int32_t)ss);
        else if (status == ST_DERIVE) up_lanes_dispatch<G, false, RANGE, ST_DERIVE>(p, m, smem_lanes, rg, (int32_t)ss);
        else up_lanes_dispatch<G, false, RANGE, ST_NONE>(p, m, smem_lanes, rg, (int32_t)ss);
    }
}

template <typename G, bool RANGE>
static int launch_lanes_geo(const UpLongParams& p0, bool loaded, int64_t O, int64_t P, int n, const LanesDecision& d) {
    UpLongParams p = p0;
    p.lanes_stage = G::stage_bytes(loaded);
    const size_t smem = (size_t)kLanesWarps * G::kStages * p.lanes_stage;
    static_assert((size_t)kLanesWarps * G::kStages * G::stage_bytes(false) >= (size_t)kLanesWarps * kLanesMaxP * 32 * (kLongStateBytes + 1),
                  "the warps' lanes are folded through the staging memory");
    static bool attr_set = false;
    auto kern = drillup_lanes_kernel<G, RANGE>;
    if (!attr_set) {
        OLAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kLanesWarps * G::kStages * G::stage_bytes(true))));
        attr_set = true;
    }
    mark_kernels_begin();
    kern<<<dim3((unsigned)(ceil_div(O, kLanesRows) * d.SS), (unsigned)n), kLanesWarps * 32, smem, g.stream>>>(p);
    ++g_launches;
    if (d.SS > 1) {
        const int64_t blocks = O * P;  // one CTA per output cell
        if (blocks > 0x7fffffffLL) return fail(OLAP_E_UNSUPPORTED, "drillUp: grid too large");
        drillup_long_merge_kernel<RANGE><<<dim3((unsigned)blocks, (unsigned)n), kLongMergeThreads, 0, g.stream>>>(p);
        ++g_launches;
    }
    return OLAP_OK;
}

int launch_up_lanes(const UpMeasure* d_meas, const UpMeasure* h_meas, int n, bool contiguous, const int32_t* d_pstart,
                    const int32_t* d_children, const uint8_t* d_lists, int64_t O, int64_t C, int64_t P,
                    const LanesDecision& d, unsigned char* d_scratch) {
    UpLongParams p{};
    p.meas = d_meas;
    if (!d_meas) for (int k = 0; k < n; ++k) p.meas_inline[k] = h_meas[k];
    p.pstart = d_pstart;
    p.children = d_children;
    p.map8 = d_lists;
    p.lanes_cs = d.Cs;
    bool loaded = false;
    for (int k = 0; k < n; ++k) loaded |= h_meas[k].st_in != nullptr;
    p.O = O; p.C = (int32_t)C; p.P = (int32_t)P; p.I = 1;
    p.SS = d.SS;
    p.row_out = (int32_t)P;
    p.div_i = FastDiv(1u);
    p.div_ss = FastDiv((uint32_t)d.SS);
    p.scratch = d_scratch;
    p.scratch_stride = d.scratch_stride;
    if (d.geo) return contiguous ? launch_lanes_geo<LanesGeoB, true>(p, loaded, O, P, n, d) : launch_lanes_geo<LanesGeoB, false>(p, loaded, O, P, n, d);
    return contiguous ? launch_lanes_geo<LanesGeoA, true>(p, loaded, O, P, n, d) : launch_lanes_geo<LanesGeoA, false>(p, loaded, O, P, n, d);
}

}  // namespace olap
