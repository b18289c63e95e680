// drillUp of a SHARDED dimension, pull model (in-memory.js:265-334).
//
// The cube is split by rows of its flattened leading dimensions across the GPUs of one box;
// a rollup of one of those dimensions needs, for every output row, child rows that live on
// several GPUs.  The rank that owns an output row reads the child rows straight out of the
// peers' stores (CUDA-IPC-mapped pointers, 128-bit loads over NVLink 5 / NVSwitch) and reduces
// them with the very lanes of the single-GPU kernel, children in ascending GLOBAL row order:
// the result is bit-equal to the unsharded rollup for every method (double accumulation in the
// reference's order, first / last exact, `average` divided once) — no partial planes, no receive
// buffers, no separate exchange or combine pass.  Every cell crosses NVLink at most once
// ((W-1)/W of the input bytes); the local HBM only sees the local children and the output.
//
// Loads of peer memory bypass the local L2 (B300_MICROARCH.md "NVLink"); one thread keeps up
// to U = 8 child vectors in flight, 2 048 threads per SM, far above the 1.5 MB of in-flight
// bytes that 770 GB/s x 2 us needs.
#pragma once
#include "kernels_drillup.cuh"

namespace olap {

struct PullMeasure {
    float* out;
    uint8_t* st_out;  // nullable: no status plane, or a plane another measure of the call writes
    int method;
    int nan_default;
    int derive;       // the sources' status planes are derived (olap_store::derived): never read, recomputed from the values
};

struct UpPullParams {
    const PullMeasure* meas;
    const int32_t* row_start;     // [rows + 1] CSR over the child rows of my output rows
    const int32_t* child_rank;    // [n_children] rank that holds the child row
    const int64_t* child_off;     // [n_children] first cell of the child row inside that rank's store
    const float* const* base_v;   // [n_measures * n_ranks] first cell of rank r's store k, in MY address space
    const uint8_t* const* base_s; // same for the status planes (entries unused when st_out is null)
    int n_ranks;
    int64_t rows;                 // my output rows
    int64_t inner;                // cells per row
    int64_t IV;                   // inner / VEC
    int64_t row0;                 // first output row of this launch (row chunks of 65 535 * blockDim.y)
};

// plain (coherent) loads: the source may be another GPU's memory
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_peer_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <int METHOD>
__device__ __noinline__ float exact_redo_pull(const UpPullParams& p, int m, int64_t cell, int32_t k0, int32_t k1) {
    Acc<METHOD> a;
    for (int32_t k = k0; k < k1; ++k) {
        const float* src = p.base_v[(size_t)m * p.n_ranks + p.child_rank[k]] + p.child_off[k] + cell;
        a.step(*reinterpret_cast<const volatile float*>(src), 1);
    }
    return a.result(1);
}

template <int METHOD, bool NANDEF, int VEC, int STATUS>
__device__ __forceinline__ void up_pull_body(const UpPullParams& p, const PullMeasure& m, int mi, int64_t row, int64_t iv) {
    constexpr int U = 8;
    const int32_t k0 = p.row_start[row], k1 = p.row_start[row + 1];
    const int64_t cell = iv * VEC;
    const float* const* base_v = p.base_v + (size_t)mi * p.n_ranks;
    const uint8_t* const* base_s = p.base_s + (size_t)mi * p.n_ranks;
    Lane<METHOD, NANDEF> lane[VEC];
    uint32_t st = 0;
    const float unset = NANDEF ? canon_nan() : 0.0f;  // what an unset cell holds: every lane skips it
    for (int32_t k = k0; k < k1; k += U) {
        Cells<VEC> c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) c[u].v[e] = unset;
            c[u].st = 0;
            if (k + u < k1) {
                const int32_t r = p.child_rank[k + u];
                const int64_t off = p.child_off[k + u] + cell;
                if (VEC == 4) {
                    const float4 t = ld_peer4(base_v[r] + off);
                    c[u].v[0] = t.x; c[u].v[1 % VEC] = t.y; c[u].v[2 % VEC] = t.z; c[u].v[3 % VEC] = t.w;
                    if (STATUS == ST_LOAD) c[u].st = ld_peer_u32(base_s[r] + off);
                } else {
                    c[u].v[0] = *reinterpret_cast<const volatile float*>(base_v[r] + off);
                    if (STATUS == ST_LOAD) c[u].st = *reinterpret_cast<const volatile uint8_t*>(base_s[r] + off);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) lane[e].step(c[u].v[e]);
            if (STATUS == ST_DERIVE) {
                if (k + u < k1) st |= derived_status<VEC, NANDEF>(c[u].v);  // a slot past the last child contributes nothing
            } else {
                st |= c[u].st;
            }
        }
    }
    const int64_t out_off = row * p.inner + cell;
    float r[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        r[e] = lane_poisoned(lane[e]) ? exact_redo_pull<METHOD>(p, mi, cell + e, k0, k1) : lane[e].result();
    store_cells<VEC>(m.out + out_off, r);
    if (STATUS) store_status<VEC>(m.st_out + out_off, k0 == k1 ? unset_status<VEC>() : st);  // no child: not set
}

template <bool NANDEF, int VEC, int STATUS>
__device__ __forceinline__ void up_pull_dispatch(const UpPullParams& p, const PullMeasure& m, int mi, int64_t row, int64_t iv) {
    switch (m.method) {
        case OLAP_SUM: up_pull_body<OLAP_SUM, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
        case OLAP_AVERAGE: up_pull_body<OLAP_AVERAGE, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
        case OLAP_HIGHEST: up_pull_body<OLAP_HIGHEST, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
        case OLAP_LOWEST: up_pull_body<OLAP_LOWEST, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
        case OLAP_FIRST: up_pull_body<OLAP_FIRST, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
        case OLAP_LAST: up_pull_body<OLAP_LAST, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
        case OLAP_COUNT: up_pull_body<OLAP_COUNT, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
        default: up_pull_body<OLAP_PRODUCT, NANDEF, VEC, STATUS>(p, m, mi, row, iv); break;
    }
}

// ---- staged variant (inner % 4 == 0): the children of an output vector are fetched with cp.async (LDGSTS)
// into the thread's own shared-memory slots — up to `chunk` (<= 16) child vectors in flight per thread without a
// single register, 3 CTAs per SM: ~190 KB in flight per SM instead of ~65 KB.  The register version is bound by
// latency x occupancy, not by the links: 2 x B200 moved 30 GB of values in 61 ms whether or not 7.5 GB of status
// bytes travelled with them.  A thread reads only its own slots: no CTA barrier anywhere.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

template <int METHOD, bool NANDEF, int STATUS>
__device__ __forceinline__ void up_pull_staged_body(const UpPullParams& p, const PullMeasure& m, int mi, int64_t row, int64_t iv,
                                                    float4* s_val, uint32_t* s_st, int chunk) {
    const int32_t k0 = p.row_start[row], k1 = p.row_start[row + 1];
    const int64_t cell = iv * 4;
    const float* const* base_v = p.base_v + (size_t)mi * p.n_ranks;
    const uint8_t* const* base_s = p.base_s + (size_t)mi * p.n_ranks;
    const int slot = threadIdx.y * blockDim.x + threadIdx.x, slots = blockDim.x * blockDim.y;
    Lane<METHOD, NANDEF> lane[4];
    uint32_t st = 0;
    for (int32_t k = k0; k < k1; k += chunk) {
        const int n = min(chunk, k1 - k);
        for (int u = 0; u < n; ++u) {
            const int32_t r = p.child_rank[k + u];
            const int64_t off = p.child_off[k + u] + cell;
            cp_async16(s_val + u * slots + slot, base_v[r] + off);
            if (STATUS == ST_LOAD) cp_async4(s_st + u * slots + slot, base_s[r] + off);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int u = 0; u < n; ++u) {
            const float4 t = s_val[u * slots + slot];
            const float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) lane[e].step(v[e]);
            if (STATUS == ST_LOAD) st |= s_st[u * slots + slot];
            if (STATUS == ST_DERIVE) st |= derived_status<4, NANDEF>(v);
        }
    }
    const int64_t out_off = row * p.inner + cell;
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
        r[e] = lane_poisoned(lane[e]) ? exact_redo_pull<METHOD>(p, mi, cell + e, k0, k1) : lane[e].result();
    store_cells<4>(m.out + out_off, r);
    if (STATUS != ST_NONE) store_status<4>(m.st_out + out_off, k0 == k1 ? unset_status<4>() : st);
}

template <bool NANDEF, int STATUS>
__device__ __forceinline__ void up_pull_staged_dispatch(const UpPullParams& p, const PullMeasure& m, int mi, int64_t row, int64_t iv,
                                                        float4* s_val, uint32_t* s_st, int chunk) {
    switch (m.method) {
        case OLAP_SUM: up_pull_staged_body<OLAP_SUM, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
        case OLAP_AVERAGE: up_pull_staged_body<OLAP_AVERAGE, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
        case OLAP_HIGHEST: up_pull_staged_body<OLAP_HIGHEST, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
        case OLAP_LOWEST: up_pull_staged_body<OLAP_LOWEST, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
        case OLAP_FIRST: up_pull_staged_body<OLAP_FIRST, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
        case OLAP_LAST: up_pull_staged_body<OLAP_LAST, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
        case OLAP_COUNT: up_pull_staged_body<OLAP_COUNT, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
        default: up_pull_staged_body<OLAP_PRODUCT, NANDEF, STATUS>(p, m, mi, row, iv, s_val, s_st, chunk); break;
    }
}

// dynamic shared memory: float4 s_val[chunk][256], then uint32 s_st[chunk][256] (ST_LOAD only)
__global__ void __launch_bounds__(256, 3) drillup_pull_staged_kernel(const __grid_constant__ UpPullParams p, int chunk) {
    extern __shared__ __align__(16) unsigned char smem_pull[];
    const int64_t row = p.row0 + (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
    const int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= p.rows || iv >= p.IV) return;
    float4* s_val = reinterpret_cast<float4*>(smem_pull);
    uint32_t* s_st = reinterpret_cast<uint32_t*>(smem_pull + (size_t)chunk * 256 * 16);
    const int mi = blockIdx.z;
    const PullMeasure m = p.meas[mi];
    const int status = m.st_out ? (m.derive ? ST_DERIVE : ST_LOAD) : ST_NONE;
    if (m.nan_default) {
        if (status == ST_LOAD) up_pull_staged_dispatch<true, ST_LOAD>(p, m, mi, row, iv, s_val, s_st, chunk);
        else if (status == ST_DERIVE) up_pull_staged_dispatch<true, ST_DERIVE>(p, m, mi, row, iv, s_val, s_st, chunk);
        else up_pull_staged_dispatch<true, ST_NONE>(p, m, mi, row, iv, s_val, s_st, chunk);
    } else {
        if (status == ST_LOAD) up_pull_staged_dispatch<false, ST_LOAD>(p, m, mi, row, iv, s_val, s_st, chunk);
        else if (status == ST_DERIVE) up_pull_staged_dispatch<false, ST_DERIVE>(p, m, mi, row, iv, s_val, s_st, chunk);
        else up_pull_staged_dispatch<false, ST_NONE>(p, m, mi, row, iv, s_val, s_st, chunk);
    }
}

// grid = (column blocks of a row, row blocks, measures), block = (bx, by): thread (tx, ty) owns output
// vector iv = blockIdx.x * bx + tx of row  row0 + blockIdx.y * by + ty.
template <int VEC>
__global__ void __launch_bounds__(256, 2) drillup_pull_kernel(const __grid_constant__ UpPullParams p) {
    const int64_t row = p.row0 + (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
    const int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= p.rows || iv >= p.IV) return;
    const int mi = blockIdx.z;
    const PullMeasure m = p.meas[mi];
    const int status = m.st_out ? (m.derive ? ST_DERIVE : ST_LOAD) : ST_NONE;
    if (m.nan_default) {
        if (status == ST_LOAD) up_pull_dispatch<true, VEC, ST_LOAD>(p, m, mi, row, iv);
        else if (status == ST_DERIVE) up_pull_dispatch<true, VEC, ST_DERIVE>(p, m, mi, row, iv);
        else up_pull_dispatch<true, VEC, ST_NONE>(p, m, mi, row, iv);
    } else {
        if (status == ST_LOAD) up_pull_dispatch<false, VEC, ST_LOAD>(p, m, mi, row, iv);
        else if (status == ST_DERIVE) up_pull_dispatch<false, VEC, ST_DERIVE>(p, m, mi, row, iv);
        else up_pull_dispatch<false, VEC, ST_NONE>(p, m, mi, row, iv);
    }
}

}  // namespace olap
