// drillup_lanes_kernel — drillUp (in-memory.js:265-334) of MANY long rows with few parents and no inner run, ANY
// map: [O >= 64, C, 1] -> [O, P <= 8, 1] with rows too long for a shared-memory tile (customers -> segment).  The
// map is the same for every row, so the kernel puts the ROW on the lane: all 32 lanes of a warp (32 different rows)
// look at the same child at the same time, which makes the parent warp-uniform, and the child lists per parent
// are walked by uniform loops that feed ONE register accumulator per parent — no divergence, no random gather.
// The long kernel reduces this shape through per-thread cell lists that hit random shared-memory banks
// (0.33-0.41 of peak).
//   * a CTA owns 32 rows x a segment of the children; each of its 4 warps owns a contiguous chunk of the segment
//     and runs its OWN cp.async pipeline over tiles of 32 rows x 128 children (512 contiguous bytes per row and
//     tile): no CTA barrier in the loop; 12 such pipelines per SM, each holding ONE tile (while some warps wait
//     for their tile the others fold theirs; the double-buffered 64-children geometry measured 3 % slower);
//   * a tile sits row-major with an odd pitch (129 words), copied 4 bytes at a time, and the copy itself SORTS the
//     columns: child c lands in column rank(c), its position in the tile's parent-grouped order (host table,
//     one byte per child).  A parent's children are then consecutive columns: the fold is a plain walk along
//     the row (lanes along rows, same column: no bank conflicts, no index loads), eight loads in flight, as two
//     chains (first half, second half, merged in order) so that two double adds are in flight per lane;
//   * a derived status plane costs a counter of the set cells and no bytes.  A plane that has to be READ stays
//     in child order (4-byte copies cannot sort bytes) and is OR-ed through the inverse table: that path is
//     built and tested (OLAP_LANES_LOADED=1) but measured no faster than the long kernel (0.27-0.28 against
//     0.25-0.30 ms), so by default such sources stay with the long kernel;
//   * once per CTA the warps' lanes are folded IN ORDER (Lane::merge: first / last stay exact); with several
//     segments per row the states go to the scratch array of the long kernel and its merge kernel folds them.
// Measured on customers -> segment [1196, 100000, 1] -> 8 (random map, derived status): sum 0.167 ms against
// 0.246 ms (long kernel), average 0.175 against 0.298; ncu: 61 M warp instructions = 16 per 32 cells (copy 3,
// load + convert + add + count 5, the rest is list bookkeeping of 16-children lists), IPC 1.6: instruction /
// latency bound at 0.44 of the HBM peak, not memory bound (profiles/ncu_lanes_r02.txt).
// The kernel lives in its own translation unit (lanes.cu).
#pragma once

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace olap {

struct UpMeasure;

constexpr int kLanesRows = 32, kLanesMaxP = 8, kLanesWarps = 4;
// Geometry of a warp's pipeline: TILE children per tile (a multiple of 32), STAGES tiles staged per warp.
template <int TILE, int STAGES>
struct LanesGeo {
    static constexpr int kTile = TILE, kStages = STAGES;
    static constexpr uint32_t kPitchV = (TILE + 1) * 4;   // bytes between rows of a value tile (odd number of words)
    static constexpr uint32_t kPitchS = TILE + 4;         // ... of a status tile (odd number of words)
    static constexpr uint32_t kValBytes = kLanesRows * kPitchV;
    static constexpr uint32_t kListBytes = TILE + 16;        // staged with the tile: positions grouped by parent + bounds
    static constexpr uint32_t kTableBytes = 2 * TILE + 16;  // per tile in the host table: + the ranks
    static constexpr uint32_t kStBytes = kLanesRows * kPitchS;
    static constexpr uint32_t stage_bytes(bool loaded) { return ((kValBytes + kListBytes + (loaded ? kStBytes : 0u)) + 15u) & ~15u; }
    static_assert(TILE % 32 == 0 && TILE <= 128, "up to four children per lane and tile (ranks travel in one word)");
};
typedef LanesGeo<64, 2> LanesGeoA;   // 12 double-buffered pipelines per SM
typedef LanesGeo<128, 1> LanesGeoB;  // 12 single-buffered pipelines per SM, lists twice as long
constexpr int kLanesLoadedDefault = 0;  // sources whose status plane has to be read stay with the long kernel (measured: profiles/)
constexpr int kLanesStateBytes = 16;  // == kLongStateBytes (asserted in lanes.cu)

struct LanesDecision {
    bool use = false;
    int geo = 0;  // 0: LanesGeoA, 1: LanesGeoB
    int32_t tile = 64;
    int32_t Cs = 0, SS = 1;
    int64_t scratch_stride = 0;
};

// `loaded_status`: some measure brings a status plane that has to be read (it is copied in words of 4 status
// bytes: rows must start on them; and one CTA less fits an SM).
inline LanesDecision lanes_plan(int64_t O, int64_t C, int64_t P, int64_t I, int n_meas, int sm_count, bool loaded_status) {
    LanesDecision d;
    // OLAP_LANES=0 sends the shape back to the long kernel (read at every call so that a test can compare the two)
    const char* knob_env = getenv("OLAP_LANES");
    if (knob_env && atoi(knob_env) == 0) return d;
    // a plane that has to be read: OLAP_LANES_LOADED=1 takes those too (measured against the long kernel in profiles/)
    if (loaded_status) {
        const char* e = getenv("OLAP_LANES_LOADED");
        if (!(e ? atoi(e) : kLanesLoadedDefault)) return d;
    }
    if (I != 1 || P > kLanesMaxP || P < 1 || O < 64 || C < 2048 || (loaded_status && (C & 3)) || C / P < 64 || C > 0x7fffffffLL || O > 0x3fffffffLL) return d;
    d.geo = 1;
    if (const char* e = getenv("OLAP_LANES_GEO")) d.geo = atoi(e) ? 1 : 0;  // tuning, tests
    d.tile = d.geo ? LanesGeoB::kTile : LanesGeoA::kTile;
    const int64_t groups = ceil_div(O, kLanesRows);
    const int64_t slots = (int64_t)sm_count * (loaded_status ? 2 : 3);
    const int64_t unit = (int64_t)d.tile * kLanesWarps;
    // segments per row: the fewest CTA waves of the shortest segments (+1 tile: the pipeline fills once per CTA)
    int64_t best_cost = INT64_MAX, best_cs = 0, best_ss = 1;
    int64_t max_ss = std::min<int64_t>(ceil_div(C, unit), 2048), min_ss = 1;
    if (const char* e = getenv("OLAP_LANES_SS")) min_ss = max_ss = std::max<int64_t>(1, std::min<int64_t>(atoi(e), max_ss));  // tests, tuning
    for (int64_t ss = min_ss; ss <= max_ss; ++ss) {
        const int64_t cs = ceil_div(ceil_div(C, ss), unit) * unit;
        const int64_t ss_eff = ceil_div(C, cs);
        if (groups * ss_eff > 0x7fffffffLL) break;
        const int64_t waves = ceil_div(groups * ss_eff * n_meas, slots);
        const int64_t cost = waves * (cs / unit + 1);
        if (cost < best_cost) { best_cost = cost; best_cs = cs; best_ss = ss_eff; }
    }
    if (!best_cs) return d;
    d.Cs = (int32_t)best_cs;
    d.SS = (int32_t)best_ss;
    if (d.SS > 1) {
        d.scratch_stride = (O * d.SS * P * (kLanesStateBytes + 1) + 255) & ~(int64_t)255;
        if (d.scratch_stride * n_meas > ((int64_t)1 << 30)) return d;
    }
    d.use = true;
    return d;
}

// Per tile of `tile` children (2 * tile + 16 bytes): their positions inside the tile grouped by parent (parents
// ascending, children ascending within a parent); 16 bytes of list bounds (entry q = where parent q starts,
// entry P = the number of children of the tile); and the inverse, the rank of every child in that order, laid out
// for the copy loop (byte j of lane l's group of tile / 32 bytes = rank of child l + 32 j).  `map` == nullptr:
// every child rolls up to parent 0.
inline std::vector<uint8_t> lanes_lists(const int32_t* map, int64_t C, int tile) {
    const int64_t tiles = ceil_div(C, tile);
    const size_t table = 2 * (size_t)tile + 16;
    const int per_lane = tile / 32;
    std::vector<uint8_t> t((size_t)tiles * table, 0);
    for (int64_t g0 = 0; g0 < tiles; ++g0) {
        uint8_t* perm = t.data() + (size_t)g0 * table;
        uint8_t* bounds = perm + tile;
        uint8_t* rank = bounds + 16;
        const int64_t c0 = g0 * tile;
        const int n = (int)std::min<int64_t>(tile, C - c0);
        int count[kLanesMaxP + 1] = {0};
        for (int k = 0; k < n; ++k) ++count[(map ? map[c0 + k] : 0) + 1];
        for (int q = 0; q < kLanesMaxP; ++q) count[q + 1] += count[q];
        for (int q = 0; q <= kLanesMaxP; ++q) bounds[q] = (uint8_t)count[q];
        for (int k = 0; k < n; ++k) {
            const int pos = count[map ? map[c0 + k] : 0]++;
            perm[pos] = (uint8_t)k;
            rank[(k & 31) * per_lane + (k >> 5)] = (uint8_t)pos;
        }
    }
    return t;
}

// lanes.cu
int launch_up_lanes(const UpMeasure* d_meas, const UpMeasure* h_meas, int n, bool contiguous, const int32_t* d_pstart,
                    const int32_t* d_children, const uint8_t* d_lists, int64_t O, int64_t C, int64_t P,
                    const LanesDecision& d, unsigned char* d_scratch);

}  // namespace olap
