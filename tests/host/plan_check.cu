// Host-only check of the launch planners (no GPU needed): transpose_plan / tile_plan must
// terminate and respect their invariants for a sweep of shapes.  Built and run by
// tests/test_host_plans.py with nvcc.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <numeric>

#include "../../olap_in_memory_b200/csrc/kernels_pair.cuh"
#include "../../olap_in_memory_b200/csrc/kernels_tile.cuh"
#include "../../olap_in_memory_b200/csrc/kernels_lanes.cuh"
#include "../../olap_in_memory_b200/csrc/host_pipe.cuh"

#include <sys/wait.h>

namespace olap {
thread_local std::string g_error;
Ctx g;
std::atomic<int64_t> g_launches{0};
int fail(int code, const char*, ...) { return code; }
void mark_kernels_begin() {}
}  // namespace olap

using namespace olap;

static int check_perm(const std::vector<int64_t>& len, const std::vector<int>& perm) {
    const int k = (int)len.size();
    std::vector<int64_t> stride(k);
    int64_t acc = 1;
    for (int i = k - 1; i >= 0; --i) { stride[i] = acc; acc *= len[i]; }
    std::vector<GDim> dims(k);
    for (int i = 0; i < k; ++i) { dims[i].len = len[perm[i]]; dims[i].linear = true; dims[i].stride = stride[perm[i]]; }
    TransposePlan plan = transpose_plan(dims);
    if (!plan.use) return 0;
    const TransposeParams& p = plan.p;
    int64_t cells = 1, boxes = 1;
    for (int a = 0; a < p.n_axes; ++a) {
        if (p.bsize[a] < 1 || p.bsize[a] > p.len[a]) { printf("bad extent\n"); return 1; }
        cells *= p.bsize[a];
        boxes *= p.boxes[a];
    }
    if (cells != p.box_cells || cells > 8192 || boxes != plan.n_boxes) { printf("bad cells %lld\n", (long long)cells); return 1; }
    if (plan.smem > 200 * 1024 || plan.rd_tab.size() != p.runs_in || plan.wr_tab.size() != p.runs_out) { printf("smem/tables\n"); return 1; }
    // entry-0 strides must be the contiguous ones
    if (p.rd[0].g_stride != 1 && p.rd[0].b > 1) { printf("rd[0] not contiguous\n"); return 1; }
    if (p.wr[0].g_stride != 1 && p.wr[0].b > 1) { printf("wr[0] not contiguous\n"); return 1; }
    uint64_t rin = 1, rout = 1;
    for (int q = 1; q < kMaxBoxDims; ++q) { rin *= p.rd[q].b; rout *= p.wr[q].b; }
    if (rin * p.rd[0].b != p.box_cells || rout * p.wr[0].b != p.box_cells) { printf("runs\n"); return 1; }
    return 0;
}

// Run transpose_pair_kernel's own phase code on the CPU (PairHostMem) and compare with the
// plain permutation: checks the planner's tables, the grid order, ragged tiles and the 4x4
// micro-tile addressing without a GPU.  Returns -1 when the pair planner declines the shape.
static int n_pair2 = 0;
static int emulate_pair(const std::vector<int64_t>& len, const std::vector<int>& perm) {
    const int k = (int)len.size();
    std::vector<int64_t> stride(k), out_len(k), out_stride(k);
    int64_t acc = 1;
    for (int i = k - 1; i >= 0; --i) { stride[i] = acc; acc *= len[i]; }
    const int64_t N = acc;
    std::vector<GDim> dims(k);
    for (int i = 0; i < k; ++i) { dims[i].len = len[perm[i]]; dims[i].linear = true; dims[i].stride = stride[perm[i]]; out_len[i] = len[perm[i]]; }
    PairPlan plan = transpose_pair_plan(dims);
    if (!plan.use) return -1;
    PairParams& p = plan.p;
    if (p.A % 4 || p.B % 4 || (p.PB / 4) % 2 == 0 || p.PB < p.B || plan.smem > 200 * 1024) { printf("pair geometry\n"); return 1; }
    if (N > (1 << 24)) return 0;  // geometry only
    p.src_row = plan.src_row.data();
    p.dst_row = plan.dst_row.data();
    // 16-byte aligned planes
    std::vector<float4> in4((N + 3) / 4 + 1), out4((N + 3) / 4 + 1);
    std::vector<uint32_t> sti((N + 3) / 4 + 1), sto((N + 3) / 4 + 1, 0xeeeeeeeeu);
    float* in = reinterpret_cast<float*>(in4.data());
    float* out = reinterpret_cast<float*>(out4.data());
    uint8_t* st_in = reinterpret_cast<uint8_t*>(sti.data());
    uint8_t* st_out = reinterpret_cast<uint8_t*>(sto.data());
    for (int64_t i = 0; i < N; ++i) { in[i] = (float)i; out[i] = -1.0f; st_in[i] = (uint8_t)(i * 7 + 3); }
    const uint32_t split = p.split;
    if (split == 2) ++n_pair2;
    std::vector<float4> smem0(plan.smem / 16 + 1), smem1(plan.smem / 16 + 1);
    unsigned char* sm[2] = {reinterpret_cast<unsigned char*>(smem0.data()), reinterpret_cast<unsigned char*>(smem1.data())};
    float* s_val_of[2] = {reinterpret_cast<float*>(sm[0]), reinterpret_cast<float*>(sm[1])};
    uint8_t* s_st_of[2] = {sm[0] + p.st_offset, sm[1] + p.st_offset};
    uint32_t* s_src_row = reinterpret_cast<uint32_t*>(sm[0] + p.tab_offset);
    uint32_t* s_dst_row = s_src_row + p.B;
    for (uint32_t i = 0; i < p.B; ++i) s_src_row[i] = p.src_row[i];
    for (uint32_t i = 0; i < p.A; ++i) s_dst_row[i] = p.dst_row[i];
    for (int64_t box = 0; box < plan.n_boxes; ++box) {
        int64_t sb, db;
        uint32_t a_eff, b_eff;
        pair_decode(p, (uint32_t)box, sb, db, a_eff, b_eff);
        if (a_eff % 4 || b_eff % 4 || sb % 4 || db % 4) { printf("pair alignment\n"); return 1; }
        if (split == 2) {
            const uint32_t mt_cta = p.nIg * p.nJqLoc, nm = (mt_cta + kPairThreads - 1) / kPairThreads;
            const uint32_t threads = ((mt_cta + nm - 1) / nm + 31) / 32 * 32;
            if (threads > (uint32_t)kPairThreads || nm > 3) { printf("pair2 threads\n"); return 1; }
            for (uint32_t rank = 0; rank < 2; ++rank)
                for (uint32_t tid = 0; tid < threads; ++tid)
                    for (uint32_t q = 0; q < nm; ++q) {
                        PairRegs2 r;
                        pair2_load<true, PairHostMem>(p, rank, tid + q * threads, in + sb, st_in + sb, s_src_row, a_eff >> 2, b_eff >> 2, r);
                        pair2_stash<true>(p, r, s_val_of[0], s_val_of[1], s_st_of[0], s_st_of[1]);
                    }
            for (uint32_t rank = 0; rank < 2; ++rank)
                for (uint32_t tid = 0; tid < threads; ++tid)
                    pair2_phase2<true, PairHostMem>(p, rank, tid, out + db, st_out + db, s_val_of[rank], s_st_of[rank], s_dst_row,
                                                    a_eff, b_eff >> 2, threads);
            continue;
        }
        float* s_val = s_val_of[0];
        uint8_t* s_st = s_st_of[0];
        const uint32_t n_mt = p.nIg * p.nJq, nm = (n_mt + kPairThreads - 1) / kPairThreads;
        const uint32_t threads = ((n_mt + nm - 1) / nm + 31) / 32 * 32;
        if (threads > (uint32_t)kPairThreads) { printf("pair threads\n"); return 1; }
        for (uint32_t tid = 0; tid < threads; ++tid)
            for (uint32_t q = 0; q < nm; ++q) {
                PairRegs r;
                pair_load<true, PairHostMem>(p, tid + q * threads, in + sb, st_in + sb, s_src_row, a_eff >> 2, b_eff >> 2, r);
                pair_stash<true>(p, r, s_val, s_st);
            }
        for (uint32_t tid = 0; tid < threads; ++tid)
            pair_phase2<true, PairHostMem>(p, tid, out + db, st_out + db, s_val, s_st, s_dst_row, a_eff, b_eff >> 2, threads);
    }
    // expected: out[new index] = in[sum coord * source stride]
    std::vector<int64_t> c(k, 0);
    for (int64_t o = 0; o < N; ++o) {
        int64_t srci = 0;
        for (int i = 0; i < k; ++i) srci += c[i] * dims[i].stride;
        if (out[o] != in[srci] || st_out[o] != st_in[srci]) { printf("pair mismatch at %lld\n", (long long)o); return 1; }
        for (int i = k - 1; i >= 0; --i) { if (++c[i] < out_len[i]) break; c[i] = 0; }
    }
    return 0;
}

// drillup_lanes_kernel's host side: the segment plan covers every child exactly once in whole tiles, and the per-tile
// lists are a grouping of the tile's children by parent (a permutation, ascending inside a parent) with its inverse in
// the copy loop's layout.
static int check_lanes(int64_t O, int64_t C, int64_t P, int n_meas, bool loaded, unsigned seed) {
    const LanesDecision d = lanes_plan(O, C, P, 1, n_meas, 148, loaded);
    if (!d.use) return 0;
    const int64_t unit = (int64_t)d.tile * kLanesWarps;
    if (d.tile != 64 && d.tile != 128) { printf("lanes: tile\n"); return 1; }
    if (d.Cs < unit || d.Cs % unit || (int64_t)d.SS * d.Cs < C || (int64_t)(d.SS - 1) * d.Cs >= C) { printf("lanes: segments %d x %d over %lld\n", d.SS, d.Cs, (long long)C); return 1; }
    if (d.SS > 1 && d.scratch_stride < O * d.SS * P * 17) { printf("lanes: scratch\n"); return 1; }
    std::vector<int32_t> map((size_t)C);
    unsigned x = seed * 2654435761u + 12345u;
    for (auto& m : map) { x = x * 1664525u + 1013904223u; m = (int32_t)((x >> 16) % (unsigned)P); }
    if (seed & 1) std::sort(map.begin(), map.end());
    const std::vector<uint8_t> t = lanes_lists(map.data(), C, d.tile);
    const size_t table = 2 * (size_t)d.tile + 16;
    const int per_lane = d.tile / 32;
    if (t.size() != (size_t)ceil_div(C, d.tile) * table || table % 16) { printf("lanes: table size\n"); return 1; }
    for (int64_t g0 = 0; g0 * d.tile < C; ++g0) {
        const uint8_t* perm = t.data() + (size_t)g0 * table;
        const uint8_t* bounds = perm + d.tile;
        const uint8_t* rank = bounds + 16;
        const int n = (int)std::min<int64_t>(d.tile, C - g0 * d.tile);
        if (bounds[0] != 0 || bounds[P] != n) { printf("lanes: bounds\n"); return 1; }
        for (int q = 0; q < 8; ++q) if (bounds[q] > bounds[q + 1] || (q >= P && bounds[q] != n)) { printf("lanes: bounds order\n"); return 1; }
        std::vector<int> seen((size_t)n, 0);
        for (int q = 0; q < P; ++q)
            for (int k = bounds[q]; k < bounds[q + 1]; ++k) {
                const int c = perm[k];
                if (c >= n || seen[(size_t)c]++ || map[(size_t)(g0 * d.tile + c)] != q) { printf("lanes: list of parent %d\n", q); return 1; }
                if (k > bounds[q] && perm[k - 1] >= c) { printf("lanes: children not ascending\n"); return 1; }
                if (rank[(c & 31) * per_lane + (c >> 5)] != k) { printf("lanes: rank is not the inverse\n"); return 1; }
            }
    }
    return 0;
}

// gather_inner_flat_kernel / gather_planes_flat_kernel: the host planners (orientation, block offsets, plane offsets,
// rows per tile) and the tile index math the kernels share with this file (flat_split), replayed on the CPU against
// the plain gather.  `dims` is the gather in output order.  Returns -1 when both planners decline.
static int n_flat = 0, n_front = 0, n_planes = 0;
static int emulate_flat(const std::vector<GDim>& dims, int64_t src_size) {
    int64_t n_out = 1;
    for (const GDim& d : dims) n_out *= d.len;
    std::vector<int64_t> want((size_t)n_out), got((size_t)n_out, -1);
    {
        std::vector<int64_t> c(dims.size(), 0);
        for (int64_t o = 0; o < n_out; ++o) {
            int64_t srci = 0;
            for (size_t i = 0; i < dims.size(); ++i) srci += dims[i].linear ? c[i] * dims[i].stride : dims[i].tbl[(size_t)c[i]];
            want[(size_t)o] = srci;
            for (int i = (int)dims.size() - 1; i >= 0; --i) { if (++c[i] < dims[i].len) break; c[i] = 0; }
        }
    }
    const FlatPlan fp = flat_plan(dims, src_size);
    if (fp.use) {
        ++(fp.front ? n_front : n_flat);
        if (fp.RB < 1 || fp.RB * fp.D > kFlatCells || (fp.RB * fp.D) % 16 || ceil_div(fp.RB * fp.K, 256) > 32 || (fp.front && fp.RB % 256)) { printf("flat: tile %lld x %lld\n", (long long)fp.RB, (long long)fp.D); return 1; }
        FlatParams p{};
        p.rows = fp.rows; p.D = (uint32_t)fp.D; p.K = (uint32_t)fp.K; p.RB = (uint32_t)fp.RB;
        p.div_k = FastDiv(p.K); p.div_rb = FastDiv(p.RB);
        if (fp.rows * fp.K != n_out) { printf("flat: output size\n"); return 1; }
        for (int64_t row0 = 0; row0 < fp.rows; row0 += fp.RB) {
            const int64_t rows_t = std::min<int64_t>(fp.RB, fp.rows - row0);
            for (uint32_t j = 0; j < p.RB * p.K; ++j) {
                const FlatIdx ix = flat_split(p, fp.front, j);
                int64_t dst;
                if (fp.front) { if (ix.k >= p.K || ix.r >= rows_t) continue; dst = (int64_t)ix.k * fp.rows + row0 + ix.r; }
                else { if (j >= rows_t * fp.K) continue; dst = row0 * fp.K + j; }
                if (ix.r >= rows_t || ix.k >= p.K || got[(size_t)dst] != -1) { printf("flat: output %lld written twice / out of tile\n", (long long)dst); return 1; }
                got[(size_t)dst] = fp.const_off + (row0 + ix.r) * fp.D + fp.keep[ix.k];
            }
        }
    } else {
        const PlanesPlan pl = planes_plan(dims, src_size);
        if (!pl.use) return -1;
        ++n_planes;
        if (pl.RB % 256 || pl.RB * pl.K > kFlatCells || pl.K > 32 || pl.rows * pl.K != n_out) { printf("planes: tile\n"); return 1; }
        for (int64_t row0 = 0; row0 < pl.rows; row0 += pl.RB) {
            const int64_t rows_t = std::min<int64_t>(pl.RB, pl.rows - row0);
            for (int64_t j = 0; j < rows_t * pl.K; ++j) {
                const int64_t r = j / pl.K, k = j % pl.K, dst = row0 * pl.K + j;
                if (pl.plane[(size_t)k] % 4 || got[(size_t)dst] != -1) { printf("planes: output %lld\n", (long long)dst); return 1; }
                got[(size_t)dst] = pl.plane[(size_t)k] + row0 + r;
            }
        }
    }
    for (int64_t o = 0; o < n_out; ++o)
        if (got[(size_t)o] != want[(size_t)o]) { printf("flat: output %lld comes from %lld, want %lld\n", (long long)o, (long long)got[(size_t)o], (long long)want[(size_t)o]); return 1; }
    return 0;
}

// host_pipe.cuh: the worker-thread copy behind the pinned ring.  Every byte arrives for sizes around the slicing
// thresholds and odd alignments, and a forked child (which has none of the workers) still copies.
static int check_copy_pool() {
    setenv("OLAP_COPY_THREADS", "4", 1);
    CopyPool& pool = CopyPool::get();
    if (pool.threads() != 4) { printf("copy pool: %d threads\n", pool.threads()); return 1; }
    const size_t cap = ((size_t)34 << 20) + 64;
    std::vector<unsigned char> src(cap), dst(cap);
    unsigned x = 7;
    for (auto& b : src) { x = x * 1664525u + 1013904223u; b = (unsigned char)(x >> 24); }
    for (size_t bytes : {(size_t)0, (size_t)1, (size_t)4095, (size_t)1 << 20, ((size_t)1 << 20) + 1, ((size_t)2 << 20) - 1, ((size_t)5 << 20) + 123, (size_t)16 << 20, ((size_t)33 << 20) + 7})
        for (size_t shift : {(size_t)0, (size_t)3}) {
            std::fill(dst.begin(), dst.end(), 0);
            pool.copy(dst.data() + shift, src.data() + 1 + shift, bytes);
            if (memcmp(dst.data() + shift, src.data() + 1 + shift, bytes) || dst[shift + bytes] != 0 || (shift && dst[shift - 1] != 0)) { printf("copy pool: %zu bytes\n", bytes); return 1; }
        }
    const pid_t child = fork();
    if (child == 0) {
        std::fill(dst.begin(), dst.end(), 0);
        pool.copy(dst.data(), src.data(), (size_t)10 << 20);
        _exit(memcmp(dst.data(), src.data(), (size_t)10 << 20) ? 1 : 0);
    }
    int status = 1;
    if (child < 0 || waitpid(child, &status, 0) != child || !WIFEXITED(status) || WEXITSTATUS(status) != 0) { printf("copy pool: forked child\n"); return 1; }
    return 0;
}

static std::vector<GDim> perm_dims(const std::vector<int64_t>& len, const std::vector<int>& perm) {
    const int k = (int)len.size();
    std::vector<int64_t> stride(k);
    int64_t acc = 1;
    for (int i = k - 1; i >= 0; --i) { stride[i] = acc; acc *= len[i]; }
    std::vector<GDim> dims(k);
    for (int i = 0; i < k; ++i) { dims[i].len = len[perm[i]]; dims[i].linear = true; dims[i].stride = stride[perm[i]]; }
    return dims;
}

int main() {
    int bad = 0, n = 0, n_pair = 0;
    bad += check_copy_pool();
    ++n;
    {
        // reorders: every permutation of a few shapes whose trailing / leading axes are short
        const std::vector<std::vector<int64_t>> fshapes = {{70, 9, 10, 10}, {1000, 6, 5, 4}, {333, 7, 3}, {300, 7, 10}, {2001, 4, 3}, {40, 50, 3, 2},
                                                           {10, 5004}, {3, 4, 5000}, {2, 70, 100}, {5, 8200, 2}, {4100, 32}, {32, 4100}, {9000, 33}};
        for (const auto& len : fshapes) {
            std::vector<int> perm(len.size());
            std::iota(perm.begin(), perm.end(), 0);
            int64_t size = 1;
            for (int64_t l : len) size *= l;
            do {
                const int r = emulate_flat(perm_dims(len, perm), size);
                if (r > 0) { printf("  shape of %zu axes, first %lld\n", len.size(), (long long)len[0]); ++bad; }
                ++n;
            } while (std::next_permutation(perm.begin(), perm.end()));
        }
        // dices: tables on the trailing axes, leading axes untouched; a table on a leading axis must be declined
        auto table_dim = [](std::vector<int64_t> items, int64_t stride) { GDim d; d.len = (int64_t)items.size(); d.linear = false; for (int64_t i : items) d.tbl.push_back(i * stride); return d; };
        auto lin_dim = [](int64_t len, int64_t stride) { GDim d; d.len = len; d.linear = true; d.stride = stride; return d; };
        bad += emulate_flat({lin_dim(1000, 10), table_dim({0, 2, 4, 6, 8}, 1)}, 10000) != 0;
        bad += emulate_flat({lin_dim(77, 130), lin_dim(13, 10), table_dim({9, 0, 4}, 1)}, 77 * 130) != 0;
        bad += emulate_flat({lin_dim(500, 120), table_dim({1, 3, 8}, 12), table_dim({0, 5, 11, 2}, 1)}, 500 * 120) != 0;
        bad += emulate_flat({lin_dim(90, 400), lin_dim(4, 100), table_dim({9, 0}, 10), lin_dim(10, 1)}, 90 * 400) != 0;
        bad += emulate_flat({table_dim({0, 2, 4}, 1000), lin_dim(100, 10), table_dim({1, 2}, 1)}, 5000) != -1;
        n += 5;
        printf("flat gathers emulated: %d block-local, %d to the front, %d planes to innermost\n", n_flat, n_front, n_planes);
        if (n_flat < 10 || n_front < 5 || n_planes < 5) { printf("flat planners declined almost everything\n"); ++bad; }
    }
    {
        int taken = 0;
        unsigned seed = 0;
        for (int64_t O : {64, 70, 1196, 100000})
            for (int64_t C : {2048, 2052, 40961, 100000, 333333})
                for (int64_t P : {1, 2, 5, 8})
                    for (int n_meas : {1, 3})
                        for (int loaded = 0; loaded < 2; ++loaded) {
                            setenv("OLAP_LANES_LOADED", "1", 1);
                            setenv("OLAP_LANES_GEO", (seed & 2) ? "1" : "0", 1);
                            const LanesDecision d = lanes_plan(O, C, P, 1, n_meas, 148, loaded != 0);
                            taken += d.use;
                            bad += check_lanes(O, C, P, n_meas, loaded != 0, seed++);
                            ++n;
                        }
        unsetenv("OLAP_LANES_LOADED");
        unsetenv("OLAP_LANES_GEO");
        // a status plane that has to be read travels in 4-byte words: odd row lengths are declined; so are short rows,
        // many parents, inner runs
        if (lanes_plan(1000, 100001, 4, 1, 1, 148, true).use || lanes_plan(10, 100000, 4, 1, 1, 148, false).use ||
            lanes_plan(1000, 100000, 9, 1, 1, 148, false).use || lanes_plan(1000, 100000, 4, 2, 1, 148, false).use) { printf("lanes: should decline\n"); ++bad; }
        printf("lanes plans taken: %d\n", taken);
        if (taken < 100) { printf("lanes planner declined almost everything\n"); ++bad; }
    }
    {
        const std::vector<std::vector<int64_t>> pshapes = {
            {20, 20, 20, 10, 10, 10}, {100, 100, 100, 10, 10, 10}, {64, 64}, {128, 36}, {36, 128}, {100, 104}, {3652, 32, 32},
            {12, 32, 32, 44}, {8, 8, 8, 8, 8}, {4, 100, 4, 100}, {1000, 1000}, {72, 200, 12}, {10, 10, 10, 10, 10, 10}, {332, 100},
            {52, 7, 92}, {44, 4, 25, 8}};
        for (const auto& len : pshapes) {
            std::vector<int> perm(len.size());
            std::iota(perm.begin(), perm.end(), 0);
            int count = 0;
            do {
                const int r = emulate_pair(len, perm);
                if (r > 0) ++bad;
                if (r == 0) ++n_pair;
                ++n;
            } while (std::next_permutation(perm.begin(), perm.end()) && ++count < 130);
        }
        printf("pair transposes emulated: %d (of which 2-CTA cluster tiles: %d)\n", n_pair, n_pair2);
        if (n_pair < 20) { printf("pair planner declined almost everything\n"); ++bad; }
    }
    const std::vector<std::vector<int64_t>> shapes = {
        {100, 100, 100, 10, 10, 10}, {3, 5000}, {5000, 3}, {7, 11, 13}, {2, 2, 2, 2, 2, 2, 2, 2}, {1000, 1000},
        {64, 64, 8}, {9, 300, 11}, {37, 50, 3, 70}, {3, 3}, {129, 3, 257}, {10, 10, 10, 10, 10, 10, 10, 10, 10},
        {1, 7, 1, 9}, {4096, 17}, {17, 4096}, {6, 6, 6, 6}, {31, 33}, {2, 100000}};
    for (const auto& len : shapes) {
        std::vector<int> perm(len.size());
        std::iota(perm.begin(), perm.end(), 0);
        int count = 0;
        do {
            bad += check_perm(len, perm);
            ++n;
        } while (std::next_permutation(perm.begin(), perm.end()) && ++count < 200);
    }
    for (int64_t C : {1, 2, 10, 29, 3652, 50000})
        for (int64_t P : {1, 2, 120})
            for (int64_t I : {1, 2, 3, 8, 31})
                for (int64_t O : {1, 5, 100000})
                    for (int st = 0; st < 2; ++st) {
                        TileDecision t = tile_plan(O, C, P, I, st);
                        ++n;
                        if (t.use && (t.R < 1 || t.smem > 208 * 1024 || t.G < 1 || t.G > 256)) { printf("tile plan\n"); ++bad; }
                    }
    printf("plan_check: %d plans, %d bad\n", n, bad);
    return bad ? 1 : 0;
}
