// Reorder (in-memory.js:178-211) with the VALUE planes moved by the TMA unit.
//
// Same tile geometry as transpose_pair_kernel (kernels_pair.cuh): a tile is the A x B block spanned by
// a contiguous INPUT run of A cells and a contiguous OUTPUT run of B cells.  What changes is who
// moves the bytes:
//   * a producer warp describes the tile to the TMA unit as a box of a <= 5-D tensor map over the
//     input plane (cp.async.bulk.tensor, SASS UTMALDG): dimension 0 is the input run, the axes of the
//     output run follow in OUTPUT order, so the box lands in shared memory as L[j][i] with j in
//     output-run order.  A ring of NST such stages is kept in flight per SM — no registers, no
//     warps waiting on loads (the pair kernel holds one tile of loads in registers: 5 warps per
//     scheduler, "no eligible warp" 70 % of the cycles, DRAM at 58 %);
//   * 640 consumer threads turn L[j][i] into S[i][j] through registers (scalar shared loads along i:
//     conflict-free; 128-bit shared stores along j);
//   * one thread hands S to the TMA unit again (UTMASTG): a box of the output tensor map whose
//     dimension 0 is the output run and whose other axes are the input-run axes in INPUT order.
// Ragged edge tiles need no code: the unit zero-fills what lies outside the tensor on the way in and
// clips on the way out.
// The status bytes cannot ride the TMA (rows of 100 / 200 bytes: boxes and strides must be multiples
// of 16 bytes), so the consumers move them exactly as the pair kernel does — 32-bit loads of 4x4
// micro-tiles prefetched one tile ahead, PRMT transposes, a padded shared tile, coalesced 32-bit stores.
#pragma once

#include <cuda.h>
#include <dlfcn.h>

#include "kernels_pair.cuh"

namespace olap {

constexpr int kTmaConsumers = 640;
constexpr int kTmaThreads = kTmaConsumers + 32;  // + the producer warp
constexpr int kTmaMaxStages = 4;

struct TmaParams {
    PairParams pp;                 // geometry, grid decomposition, status tables
    const CUtensorMap* maps;       // [2 * n_measures] in global memory: input map, output map of each measure
    int in_rank, out_rank;
    int in_dim[OLAP_MAX_DIMS], out_dim[OLAP_MAX_DIMS];        // per grid slot: which tensor-map dimension its start feeds
    int32_t in_mul[OLAP_MAX_DIMS], out_mul[OLAP_MAX_DIMS];    // ... and with which multiplier (merged dimensions)
    FastDiv div_A;                 // by the input-run length
    uint32_t nst, ns;              // load stages, store stages
    uint32_t stage_stride;         // bytes between stages (A * B * 4 rounded up to 128)
    uint32_t off_S, off_st, off_tab, off_bar;  // byte offsets in dynamic shared memory (L stages start at 0)
};

struct TmaPlan {
    bool use = false;
    TmaParams p{};
    PairPlan pair;
    size_t smem = 0;
    // tensor-map descriptions (per measure only the base address differs)
    int in_rank = 0, out_rank = 0;
    uint64_t in_len[5], out_len[5], in_stride[5], out_stride[5];  // strides in BYTES (entry 0 unused: contiguous)
    uint32_t in_box[5], out_box[5];
};

// ---- PTX wrappers -------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tma_smem(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem(bar)), "r"(count));
}
__device__ __forceinline__ void tma_mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma_smem(bar)) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TDONE_%=;\n"
        "bra TWAIT_%=;\n"
        "TDONE_%=:\n"
        "}\n" ::"r"(tma_smem(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"r"(kTmaConsumers) : "memory"); }

// global -> shared, box of a rank-R tensor map at coordinates c[0..R)
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* map, const int32_t* c, int rank, uint64_t* bar) {
    const uint32_t d = tma_smem(dst), b = tma_smem(bar);
    const uint64_t m = reinterpret_cast<uint64_t>(map);
    if (rank == 2)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(d), "l"(m), "r"(b), "r"(c[0]), "r"(c[1]) : "memory");
    else if (rank == 3)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(d), "l"(m), "r"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
    else if (rank == 4)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(d), "l"(m), "r"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(d), "l"(m), "r"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
// shared -> global
__device__ __forceinline__ void tma_store(const CUtensorMap* map, const int32_t* c, int rank, const void* src) {
    const uint32_t s = tma_smem(src);
    const uint64_t m = reinterpret_cast<uint64_t>(map);
    if (rank == 2)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(s), "r"(c[0]), "r"(c[1]) : "memory");
    else if (rank == 3)
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(s), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
    else if (rank == 4)
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(m), "r"(s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// tile index -> coordinates of its box in both tensor maps, and the element offsets the status path needs
__device__ __forceinline__ void tma_decode(const TmaParams& p, uint32_t tile, int32_t* cin, int32_t* cout, int64_t& src,
                                           int64_t& dst, uint32_t& a_eff, uint32_t& b_eff) {
    const PairParams& g = p.pp;
#pragma unroll
    for (int d = 0; d < 5; ++d) cin[d] = cout[d] = 0;
    uint32_t rest = tile;
    a_eff = g.A;
    b_eff = g.B;
    src = dst = 0;
    for (int a = g.n_axes - 1; a >= 0; --a) {
        const uint32_t q = g.div_boxes[a].div(rest);
        const uint32_t bi = rest - q * g.boxes[a];
        rest = q;
        const uint32_t start = bi * g.bsize[a];
        src += (int64_t)start * g.src_stride[a];
        dst += (int64_t)start * g.dst_stride[a];
        const uint32_t rem = g.len[a] - start, e = g.bsize[a] < rem ? g.bsize[a] : rem;
        if (a == g.in_axis) a_eff = e * g.in_mult;
        if (a == g.out_axis) b_eff = e * g.out_mult;
        // no dynamic indexing of the coordinate arrays (they must stay in registers)
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            if (p.in_dim[a] == d) cin[d] += (int32_t)start * p.in_mul[a];
            if (p.out_dim[a] == d) cout[d] += (int32_t)start * p.out_mul[a];
        }
    }
}

// status bytes of one 4x4 micro-tile, loaded ahead (same addressing as pair_load)
struct TmaStatusRegs {
    uint32_t b0, b1, b2, b3;
    uint32_t sidx;  // index in the padded shared status tile, or 0xffffffff: nothing loaded
};
__device__ __forceinline__ void tma_status_load(const PairParams& g, uint32_t mt, const uint8_t* st_src, const uint32_t* s_src_row,
                                                uint32_t n_ig, uint32_t n_jq, TmaStatusRegs& r) {
    const uint32_t jg = g.div_nIg.div(mt), ig = mt - jg * g.nIg;
    r.sidx = 0xffffffffu;
    if (ig < n_ig && jg < n_jq) {
        const uint4 ro = *reinterpret_cast<const uint4*>(s_src_row + 4 * jg);
        r.b0 = ld_stream_u32(st_src + (((size_t)ro.x + ig) << 2));
        r.b1 = ld_stream_u32(st_src + (((size_t)ro.y + ig) << 2));
        r.b2 = ld_stream_u32(st_src + (((size_t)ro.z + ig) << 2));
        r.b3 = ld_stream_u32(st_src + (((size_t)ro.w + ig) << 2));
        r.sidx = ig * g.PB + 4 * jg;
    }
}
__device__ __forceinline__ void tma_status_stash(const PairParams& g, const TmaStatusRegs& r, uint8_t* s_st) {
    if (r.sidx == 0xffffffffu) return;
    const uint32_t plane = g.nIg * g.PB;
    const uint32_t p01 = pair_prmt(r.b0, r.b1, 0x5140), q01 = pair_prmt(r.b0, r.b1, 0x7362);
    const uint32_t p23 = pair_prmt(r.b2, r.b3, 0x5140), q23 = pair_prmt(r.b2, r.b3, 0x7362);
    *reinterpret_cast<uint32_t*>(s_st + r.sidx) = pair_prmt(p01, p23, 0x5410);
    *reinterpret_cast<uint32_t*>(s_st + r.sidx + plane) = pair_prmt(p01, p23, 0x7632);
    *reinterpret_cast<uint32_t*>(s_st + r.sidx + 2 * plane) = pair_prmt(q01, q23, 0x5410);
    *reinterpret_cast<uint32_t*>(s_st + r.sidx + 3 * plane) = pair_prmt(q01, q23, 0x7632);
}

template <bool STATUS>
__device__ __forceinline__ void tma_body(const TmaParams& p, const GatherMeasure& m, unsigned char* smem, uint32_t n_boxes) {
    const PairParams& g = p.pp;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_bar);  // [nst] a stage holds its tile
    uint64_t* empty = full + kTmaMaxStages;                          // [nst] a stage may be refilled
    const CUtensorMap* map_in = p.maps + 2 * blockIdx.y;
    const CUtensorMap* map_out = map_in + 1;
    const uint32_t A = g.A, B = g.B;
    const uint32_t tile_bytes = A * B * 4u;

    if (threadIdx.x >= kTmaConsumers) {
        // ---------------- producer warp: one lane keeps the ring of load stages full ----------------
        if (threadIdx.x == kTmaConsumers) {
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < n_boxes; tile += gridDim.x, ++it) {
                const uint32_t s = it % p.nst, ph = (it / p.nst) & 1u;
                tma_mbar_wait(&empty[s], ph ^ 1u);  // fresh barriers: the first pass falls through
                int32_t cin[5], cout[5];
                int64_t sb, db;
                uint32_t ae, be;
                tma_decode(p, tile, cin, cout, sb, db, ae, be);
                tma_mbar_expect(&full[s], tile_bytes);
                tma_load(smem + (size_t)s * p.stage_stride, map_in, cin, p.in_rank, &full[s]);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const uint32_t tid = threadIdx.x;
    uint8_t* s_st = smem + p.off_st;
    const uint32_t* s_src_row = reinterpret_cast<const uint32_t*>(smem + p.off_tab);
    const uint32_t* s_dst_row = s_src_row + B;
    constexpr int NM = 2;  // micro-tiles of status per thread (2 * 640 >= 25 * 50)
    TmaStatusRegs sr[NM];
    int32_t cin[5], cout[5];
    int64_t sb, db;
    uint32_t a_eff, b_eff;
    uint32_t tile = blockIdx.x;
    if (tile >= n_boxes) return;
    tma_decode(p, tile, cin, cout, sb, db, a_eff, b_eff);
    if (STATUS) {
#pragma unroll
        for (int q = 0; q < NM; ++q) tma_status_load(g, tid + q * kTmaConsumers, m.st_in + sb, s_src_row, a_eff >> 2, b_eff >> 2, sr[q]);
    }
    const uint32_t n_items = A * (B >> 2);  // (i, jq): 4 cells of output run i
    for (uint32_t it = 0;; ++it) {
        const uint32_t s = it % p.nst, ph = (it / p.nst) & 1u, q_s = it % p.ns;
        const float* L = reinterpret_cast<const float*>(smem + (size_t)s * p.stage_stride);
        float* S = reinterpret_cast<float*>(smem + p.off_S + (size_t)q_s * p.stage_stride);
        if (STATUS) {
#pragma unroll
            for (int q = 0; q < NM; ++q) tma_status_stash(g, sr[q], s_st);
        }
        // the store that used this S stage NS tiles ago has finished reading it
        if (tid == 0) {
            if (p.ns == 1) tma_store_wait_read<0>();
            else tma_store_wait_read<1>();
        }
        tma_mbar_wait(&full[s], ph);
        tma_consumer_sync();  // S stage free for everybody, status tile complete
        // L[j][i] -> S[i][j]: lanes walk i (conflict-free scalar loads), every item is a 128-bit store along j
        for (uint32_t x = tid; x < n_items; x += kTmaConsumers) {
            const uint32_t jq = p.div_A.div(x), i = x - jq * A;
            const float* col = L + (size_t)(4 * jq) * A + i;
            const float4 v = make_float4(col[0], col[A], col[2 * A], col[3 * A]);
            *reinterpret_cast<float4*>(S + (size_t)i * B + 4 * jq) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my shared stores -> visible to the TMA unit
        tma_consumer_sync();
        const int64_t db_cur = db;
        const uint32_t a_cur = a_eff, nj_cur = b_eff >> 2;
        if (tid == 0) {
            tma_mbar_arrive(&empty[s]);  // every consumer is past its last read of L[s]
            tma_store(map_out, cout, p.out_rank, S);
        }
        tile += gridDim.x;
        const bool more = tile < n_boxes;
        if (more) {
            tma_decode(p, tile, cin, cout, sb, db, a_eff, b_eff);
            if (STATUS) {  // next tile's status bytes fly while this tile's leave
#pragma unroll
                for (int q = 0; q < NM; ++q) tma_status_load(g, tid + q * kTmaConsumers, m.st_in + sb, s_src_row, a_eff >> 2, b_eff >> 2, sr[q]);
            }
        }
        if (STATUS) {
            // status rows of the finished tile: coalesced 32-bit stores along the output runs
            uint8_t* st_dst = m.st_out + db_cur;
            const uint32_t n_it = A * g.nJq;
            for (uint32_t x = tid; x < n_it; x += kTmaConsumers) {
                const uint32_t i = g.div_nJq.div(x), jq = x - i * g.nJq;
                if (i < a_cur && jq < nj_cur) {
                    const uint32_t sidx = ((i & 3u) * g.nIg + (i >> 2)) * g.PB + 4 * jq;
                    *reinterpret_cast<uint32_t*>(st_dst + (((size_t)s_dst_row[i] + jq) << 2)) = *reinterpret_cast<const uint32_t*>(s_st + sidx);
                }
            }
            tma_consumer_sync();  // the status tile may be overwritten
        }
        if (!more) break;
    }
    if (tid == 0) tma_store_wait_all();  // shared memory must outlive the last stores
}

__global__ void __launch_bounds__(kTmaThreads, 1) transpose_tma_kernel(const __grid_constant__ TmaParams p, uint32_t n_boxes) {
    extern __shared__ __align__(128) unsigned char smem_tma[];
    const PairParams& g = p.pp;
    uint32_t* s_src_row = reinterpret_cast<uint32_t*>(smem_tma + p.off_tab);
    uint32_t* s_dst_row = s_src_row + g.B;
    const GatherMeasure m = g.meas[blockIdx.y];
    for (uint32_t i = threadIdx.x; i < g.B; i += blockDim.x) s_src_row[i] = __ldg(g.src_row + i);
    for (uint32_t i = threadIdx.x; i < g.A; i += blockDim.x) s_dst_row[i] = __ldg(g.dst_row + i);
    if (threadIdx.x == 0) {
        uint64_t* full = reinterpret_cast<uint64_t*>(smem_tma + p.off_bar);
        for (uint32_t s = 0; s < p.nst; ++s) {
            tma_mbar_init(&full[s], 1);
            tma_mbar_init(&full[kTmaMaxStages + s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (m.st_in) tma_body<true>(p, m, smem_tma, n_boxes);
    else tma_body<false>(p, m, smem_tma, n_boxes);
}
#endif  // __CUDACC__

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmaEncodeFn tma_encode_fn() {
    static TmaEncodeFn fn = [] {
        void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        return drv ? reinterpret_cast<TmaEncodeFn>(dlsym(drv, "cuTensorMapEncodeTiled")) : nullptr;
    }();
    return fn;
}

// One side of the tile as a tensor map: dimension 0 = the run of `run_group` (merged axes), then
// the axes of the OTHER group that have an extent > 1 inside the tile, ordered by their stride on the
// other side (so that the box lands / leaves in the order the transposition wants), then the
// extent-1 axes (merged where they are contiguous).  Returns false when it needs more than 5 dimensions.
inline bool tma_side(const std::vector<PairAxis>& axes, bool input_side, int run_group, int& rank, uint64_t* len,
                     uint64_t* stride_bytes, uint32_t* box, int* dim_of_slot, int32_t* mul_of_slot) {
    auto st_here = [&](const PairAxis& a) { return input_side ? a.src_stride : a.dst_stride; };
    auto st_there = [&](const PairAxis& a) { return input_side ? a.dst_stride : a.src_stride; };
    std::vector<int> order(axes.size());
    for (size_t i = 0; i < axes.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return st_here(axes[x]) < st_here(axes[y]); });
    struct Dim { uint64_t len; int64_t stride; uint32_t box; std::vector<std::pair<int, int64_t>> parts; bool other = false; int64_t key = 0; bool closed = false; };
    std::vector<Dim> dims;
    // dimension 0: the axes of the run, innermost first (contiguous on this side by construction)
    Dim d0{1, 1, 1, {}, false, 0, false};
    size_t pos = 0;
    for (; pos < order.size() && axes[order[pos]].group == run_group; ++pos) {
        const PairAxis& a = axes[order[pos]];
        if (st_here(a) != (int64_t)d0.len) return false;
        d0.parts.push_back({order[pos], (int64_t)d0.len});
        d0.box *= (uint32_t)a.box;
        if (a.box != a.len) d0.closed = a.len % a.box != 0;  // a ragged split must not run into the next axis
        d0.len *= (uint64_t)a.len;
    }
    if (d0.parts.empty()) return false;
    dims.push_back(d0);
    for (; pos < order.size(); ++pos) {
        const PairAxis& a = axes[order[pos]];
        if (a.group == run_group) return false;  // the run is not the innermost block on this side
        Dim& last = dims.back();
        const bool contiguous = st_here(a) == last.stride * (int64_t)last.len;
        if (a.box == 1 && contiguous && !last.other && !last.closed) {
            // extent 1: rides on the previous dimension as a coordinate offset
            last.parts.push_back({order[pos], (int64_t)last.len});
            last.len *= (uint64_t)a.len;
            continue;
        }
        Dim d{(uint64_t)a.len, st_here(a), (uint32_t)a.box, {{order[pos], 1}}, a.box > 1, st_there(a), false};
        dims.push_back(d);
    }
    // order: dim 0, the other group's axes by their stride on the other side, the rest
    std::stable_sort(dims.begin() + 1, dims.end(), [](const Dim& x, const Dim& y) {
        if (x.other != y.other) return x.other;
        return x.other && x.key < y.key;
    });
    if (dims.size() > 5) return false;
    rank = (int)dims.size();
    if (rank < 2) return false;
    for (int d = 0; d < rank; ++d) {
        len[d] = dims[d].len;
        stride_bytes[d] = (uint64_t)dims[d].stride * 4u;
        box[d] = dims[d].box;
        if (dims[d].len > 0xffffffffull || box[d] > 256 || (d > 0 && (stride_bytes[d] % 16 || stride_bytes[d] >= (1ull << 40)))) return false;
        for (auto& part : dims[d].parts) {
            if (part.second > 0x7fffffff) return false;
            dim_of_slot[axes[part.first].slot] = d;
            mul_of_slot[axes[part.first].slot] = (int32_t)part.second;
        }
    }
    return (box[0] * 4u) % 16 == 0;
}

inline TmaPlan transpose_tma_plan(const std::vector<GDim>& dims) {
    TmaPlan plan;
    // Opt-in (OLAP_TRANSPOSE_TMA=1): measured SLOWER than the pair kernel on the 6-D reversal of config 3
    // (1.88 - 2.37 ms against 1.77 ms without status plane, 3.1 - 3.8 ms against 2.36 ms with it) — the time per
    // tile is set by the SM (TMA unit on 400-byte box rows + the shared-memory transposition), not by load
    // latency: halving the CTAs doubles the time, deeper rings do not help (profiles/README.md).  The knobs
    // are read at every call so that a test can switch the path on.
    auto env_int = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
    const int knob = env_int("OLAP_TRANSPOSE_TMA");
    const int64_t want_in = env_int("OLAP_TMA_IN"), want_out = env_int("OLAP_TMA_OUT");
    if (!knob) return plan;
    plan.pair = transpose_pair_plan_for(dims, want_in, want_out);
    if (!plan.pair.geometry) return plan;
    const PairParams& g = plan.pair.p;
    if (g.nIg * g.nJq > 2u * kTmaConsumers) return plan;  // status micro-tiles per thread
    TmaParams& p = plan.p;
    for (int a = 0; a < OLAP_MAX_DIMS; ++a) { p.in_dim[a] = p.out_dim[a] = -1; p.in_mul[a] = p.out_mul[a] = 0; }
    if (!tma_side(plan.pair.axes, true, 1, plan.in_rank, plan.in_len, plan.in_stride, plan.in_box, p.in_dim, p.in_mul)) return plan;
    if (!tma_side(plan.pair.axes, false, 2, plan.out_rank, plan.out_len, plan.out_stride, plan.out_box, p.out_dim, p.out_mul)) return plan;
    // the boxes must be exactly the A x B tile
    uint64_t vin = 1, vout = 1;
    for (int d = 0; d < plan.in_rank; ++d) vin *= plan.in_box[d];
    for (int d = 0; d < plan.out_rank; ++d) vout *= plan.out_box[d];
    if (plan.in_box[0] != g.A || plan.out_box[0] != g.B || vin != (uint64_t)g.A * g.B || vout != vin) return plan;
    p.pp = g;
    p.div_A = FastDiv(g.A);
    p.in_rank = plan.in_rank;
    p.out_rank = plan.out_rank;
    const size_t tile_bytes = (size_t)g.A * g.B * 4;
    p.stage_stride = (uint32_t)((tile_bytes + 127) & ~(size_t)127);
    const size_t st_bytes = ((size_t)g.A * g.PB + 127) & ~(size_t)127;  // 4 planes of nIg rows of PB bytes
    const size_t tab_bytes = (((size_t)g.A + g.B) * 4 + 127) & ~(size_t)127;
    const size_t fixed = st_bytes + tab_bytes + 2 * kTmaMaxStages * 8 + 128;
    const size_t budget = 227 * 1024;
    const int nst_knob = env_int("OLAP_TMA_STAGES"), ns_knob = env_int("OLAP_TMA_STORE_STAGES");
    const size_t stages = (budget - fixed) / p.stage_stride;
    if (stages < 2) return plan;
    p.ns = stages >= 4 ? 2 : 1;
    if (ns_knob) p.ns = (uint32_t)ns_knob;
    p.nst = (uint32_t)std::min<size_t>(kTmaMaxStages, stages - p.ns);
    if (nst_knob) p.nst = (uint32_t)nst_knob;
    if (p.nst < 1 || p.ns < 1 || p.ns > 2 || p.nst > kTmaMaxStages || (size_t)(p.nst + p.ns) * p.stage_stride + fixed > budget) return plan;
    p.off_S = p.nst * p.stage_stride;
    p.off_st = p.off_S + p.ns * p.stage_stride;
    p.off_tab = (uint32_t)(p.off_st + st_bytes);
    p.off_bar = (uint32_t)(p.off_tab + tab_bytes);
    plan.smem = p.off_bar + 2 * kTmaMaxStages * 8;
    plan.use = true;
    return plan;
}

// the 2 * n tensor maps of a call (input / output plane of every measure), ready to upload
inline int tma_encode_maps(const TmaPlan& plan, const std::vector<GatherMeasure>& meas, std::vector<CUtensorMap>& maps) {
    static const int l2_knob = [] { const char* e = getenv("OLAP_TMA_L2"); return e ? atoi(e) : 0; }();
    const CUtensorMapL2promotion l2 = l2_knob == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                    : l2_knob == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    maps.resize(2 * meas.size());
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    for (size_t k = 0; k < meas.size(); ++k) {
        for (int side = 0; side < 2; ++side) {
            const int rank = side ? plan.out_rank : plan.in_rank;
            const uint64_t* len = side ? plan.out_len : plan.in_len;
            const uint64_t* str = side ? plan.out_stride : plan.in_stride;
            const uint32_t* box = side ? plan.out_box : plan.in_box;
            cuuint64_t gdim[5], gstr[4];
            cuuint32_t gbox[5];
            for (int d = 0; d < rank; ++d) { gdim[d] = len[d]; gbox[d] = box[d]; }
            for (int d = 1; d < rank; ++d) gstr[d - 1] = str[d];
            void* base = side ? (void*)meas[k].out : (void*)meas[k].in;
            const CUresult r = tma_encode_fn()(&maps[2 * k + side], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base, gdim, gstr, gbox,
                                               ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(OLAP_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for the %s plane", (int)r, side ? "output" : "input");
        }
    }
    return OLAP_OK;
}

#ifdef __CUDACC__
inline int launch_transpose_tma(const GatherMeasure* d_meas, const uint32_t* d_src_row, const uint32_t* d_dst_row,
                                const CUtensorMap* d_maps, int n, TmaPlan& plan) {
    plan.p.pp.meas = d_meas;
    plan.p.pp.src_row = d_src_row;
    plan.p.pp.dst_row = d_dst_row;
    plan.p.maps = d_maps;
    static bool attr_set = false;
    if (!attr_set) {
        OLAP_CUDA(cudaFuncSetAttribute(transpose_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    static const int ctas_knob = [] { const char* e = getenv("OLAP_TMA_CTAS"); return e ? atoi(e) : 0; }();
    const int64_t ctas = std::min<int64_t>(plan.pair.n_boxes, std::max<int64_t>(1, ceil_div((int64_t)(ctas_knob ? ctas_knob : g.sm_count), n)));
    const dim3 grid((unsigned)ctas, (unsigned)n);
    mark_kernels_begin();
    transpose_tma_kernel<<<grid, kTmaThreads, plan.smem, g.stream>>>(plan.p, (uint32_t)plan.pair.n_boxes);
    OLAP_CUDA(cudaGetLastError());
    ++g_launches;
    return OLAP_OK;
}
#endif

}  // namespace olap
