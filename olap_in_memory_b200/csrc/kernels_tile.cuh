// Shared-memory staged kernels (filled in after the first measurements):
//  - drillUp along the innermost / a short inner axis: rows of C*I contiguous cells are
//    staged in shared memory with bulk async copies and reduced per parent;
//  - reorder whose innermost axis moves: tiled transpose.
#pragma once
#include "common.cuh"
#include "kernels_drillup.cuh"
#include "kernels_gather.cuh"

namespace olap {

struct TileDecision {
    bool use = false;
};
inline TileDecision tile_plan(int64_t, int64_t, int64_t, int64_t, int) { return TileDecision{}; }
inline int launch_up_tile(const UpMeasure*, int, bool, const int32_t*, const int32_t*, int64_t, int64_t, int64_t, int64_t,
                          const TileDecision&) {
    return fail(OLAP_E_UNSUPPORTED, "tile path not built");
}

struct TransposePlan {
    bool use = false;
};
inline TransposePlan transpose_plan(const std::vector<GDim>&) { return TransposePlan{}; }
inline int launch_transpose(const std::vector<GatherMeasure>&, int, const TransposePlan&) {
    return fail(OLAP_E_UNSUPPORTED, "transpose path not built");
}

}  // namespace olap
