// a17  KMeansClustering (models/aug_utils.py:142-157): the Lloyd assignment pass, deterministic.
//
// Rows are split STATICALLY: CTA b owns a contiguous chunk, warp w of it a contiguous sub-chunk it walks in order, adding each
// row into its own [K, d] slab in shared memory.  The slabs are summed in warp order and written as one partial per CTA;
// kmeans_update_kernel sums the partials in CTA order.  No floating-point atomics: the centroids are a pure function of the
// inputs, whatever the scheduling.
// Distances: lane l evaluates centroid k = 32 r + l against the row staged in shared memory, sum_j (x_j - c_kj)^2 in index
// order (the reference's (x - c).square().sum(-1)); centroid rows are padded to d + 1 floats so the 32 lanes hit 32 banks.
// argmin ties -> lowest centroid id.
//
// kmeans_assign_kernel<R>: a warp stages R rows per round and every lane runs R independent distance chains per centroid (the
// round-2 ncu capture showed the R = 1 form latency-bound: 8 warps per SM, one dependent FADD -> FFMA chain of length d per
// centroid, 12 % of the warp slots, 0.25 ms per iteration at the amazon shape).  The arithmetic of a (row, centroid) pair, the
// argmin reduction and the order in which a warp adds its rows into its slab are those of R = 1, so assignments, partial sums,
// counts and the change counter are BIT-IDENTICAL for every R (checked on the host, tests/emu/kmeans_emu.cpp, and on the GPU,
// tests/test_gpu_kernels.py).  ssl_set_option("kmeans_rows_per_round", 1) selects the R = 1 instantiation.
//
// Kernel source only (no runtime API): tests/emu compiles this file for the host.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef SSL_HOST_EMU
#define SSL_DYN_SMEM_FLOAT(name) float *name = emu_dyn_smem_float()
#else
#define SSL_DYN_SMEM_FLOAT(name) extern __shared__ float name[]
#endif

namespace ssl_kmeans {

constexpr int kMaxRowsPerRound = 4;

// floats of dynamic shared memory of a launch with W warps staging R rows each
__host__ __device__ inline size_t smem_floats(int K, int dim, int W, int R) {
    return (size_t)K * (dim + 1) + (size_t)W * R * dim + (size_t)W * K * dim + (size_t)W * K;
}

template <int R>
static __global__ void kmeans_assign_kernel(const float *__restrict__ x, int64_t stride, int64_t n, int dim, int K,
                                            const float *__restrict__ cents, int64_t *__restrict__ assign,
                                            float *__restrict__ part_sum, float *__restrict__ part_cnt,
                                            int *__restrict__ changed, int64_t rows_per_cta, int64_t rows_per_warp) {
    SSL_DYN_SMEM_FLOAT(smem);
    const int W = (int)blockDim.x >> 5, warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    const int cpad = dim + 1;
    float *cs = smem;                              // [K, dim + 1]
    float *xs = cs + (size_t)K * cpad;             // [W, R, dim]
    float *slab = xs + (size_t)W * R * dim;        // [W, K, dim]
    float *cnt = slab + (size_t)W * K * dim;       // [W, K]
    for (int e = (int)threadIdx.x; e < K * dim; e += (int)blockDim.x) cs[(e / dim) * cpad + (e % dim)] = cents[e];
    for (int e = (int)threadIdx.x; e < W * K * dim; e += (int)blockDim.x) slab[e] = 0.f;
    for (int e = (int)threadIdx.x; e < W * K; e += (int)blockDim.x) cnt[e] = 0.f;
    __syncthreads();

    float *myx = xs + (size_t)warp * R * dim;
    float *myslab = slab + (size_t)warp * K * dim;
    const int64_t cta0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r0 = cta0 + (int64_t)warp * rows_per_warp;
    const int64_t cta_end = cta0 + rows_per_cta < n ? cta0 + rows_per_cta : n;
    const int64_t r1 = r0 + rows_per_warp < cta_end ? r0 + rows_per_warp : cta_end;
    int n_changed = 0;
    for (int64_t r = r0; r < r1; r += R) {
        const int nr = (r1 - r < (int64_t)R) ? (int)(r1 - r) : R;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            if (q < nr) {
                const float *xr = x + (r + q) * stride;
                for (int j = lane; j < dim; j += 32) myx[q * dim + j] = __ldg(xr + j);
            } else {
                for (int j = lane; j < dim; j += 32) myx[q * dim + j] = 0.f;      // tail of the sub-chunk: computed, never applied
            }
        }
        __syncwarp();
        float best[R];
        int best_k[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            best[q] = INFINITY;
            best_k[q] = 0x7fffffff;
        }
        for (int k = lane; k < K; k += 32) {
            const float *c = cs + (size_t)k * cpad;
            float d2[R];
#pragma unroll
            for (int q = 0; q < R; ++q) d2[q] = 0.f;
            for (int j = 0; j < dim; ++j) {
                const float cj = c[j];
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const float t = myx[q * dim + j] - cj;
                    d2[q] = fmaf(t, t, d2[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (d2[q] < best[q]) {             // k ascends per lane: strict < keeps the lowest id
                    best[q] = d2[q];
                    best_k[q] = k;
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const float ob = __shfl_xor_sync(0xffffffffu, best[q], o);
                const int ok = __shfl_xor_sync(0xffffffffu, best_k[q], o);
                if (ob < best[q] || (ob == best[q] && ok < best_k[q])) {
                    best[q] = ob;
                    best_k[q] = ok;
                }
            }
        }
        // the round's rows enter the warp's slab in row order: lane j adds element j of row q before element j of row q + 1
#pragma unroll
        for (int q = 0; q < R; ++q) {
            if (q < nr) {
                const int bk = best_k[q] >= K ? 0 : best_k[q];      // all distances NaN: torch.min returns index 0 as well
                float *dst = myslab + (size_t)bk * dim;
                for (int j = lane; j < dim; j += 32) dst[j] += myx[q * dim + j];
                if (lane == 0) {
                    cnt[warp * K + bk] += 1.f;
                    if (assign[r + q] != (int64_t)bk) ++n_changed;
                    assign[r + q] = bk;
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0 && n_changed) atomicAdd(changed, n_changed);
    __syncthreads();
    float *ps = part_sum + (size_t)blockIdx.x * K * dim;
    for (int e = (int)threadIdx.x; e < K * dim; e += (int)blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < W; ++w) s += slab[(size_t)w * K * dim + e];
        ps[e] = s;
    }
    for (int e = (int)threadIdx.x; e < K; e += (int)blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < W; ++w) s += cnt[w * K + e];
        part_cnt[(size_t)blockIdx.x * K + e] = s;
    }
}

}  // namespace ssl_kmeans
