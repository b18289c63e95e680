// Host-only check of the launch planners (no GPU needed): transpose_plan / tile_plan must
// terminate and respect their invariants for a sweep of shapes.  Built and run by
// tests/test_host_plans.py with nvcc.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <numeric>

#include "../../olap_in_memory_b200/csrc/kernels_tile.cuh"

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

int main() {
    int bad = 0, n = 0;
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
