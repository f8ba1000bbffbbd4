// full_predict + _mask_predict (lightgcn.py:58-66, base_model.py:35-36) as a register-tiled product:
//   preds[b, i] = (U[users[b]] . I[i]) * (1 - M[b, i]) - 1e8 * M[b, i]
// The score block of an evaluation batch is a [n_b, dim] x [dim, n_item] product with a tiny inner dimension (dim <= 128), so
// the bound is the write of the [n_b, n_item] score matrix.  A CTA owns a 128 x 128 tile of scores: 16 x 16 threads, 8 x 8
// outputs each (rows ty + 16 i, columns tx + 16 j: shared-memory reads are conflict-free broadcasts, global stores 64-byte
// runs), operands staged 32 inner-dimension values at a time through shared memory (k-major, padded rows: the transposing
// stores are conflict-free), every item row read once per 128 users instead of once per user.  Each score is one sequential
// fp32 FMA chain over k = 0 .. dim-1.  The mask is applied in the same launch: the reference's dense [n_b, n_item] tensor in the
// epilogue, or -- after a CTA barrier -- the training CSR rows of the tile's users scattered over the tile's columns.
//
// This header holds the kernel only (no runtime API), so that tests/emu can compile the same source for the host and run it
// thread by thread (pthreads + barrier) under the address and thread sanitizers.
#pragma once
#include <stdint.h>

namespace ssl_predict {

constexpr int TM = 128;      // users per tile
constexpr int TN = 128;      // items per tile
constexpr int TK = 32;       // inner-dimension values staged per round
constexpr int NT = 256;      // threads: 16 (columns) x 16 (rows)
constexpr int PAD = 1;

static __global__ void __launch_bounds__(NT)
predict_tile_kernel(const float *__restrict__ ut, int64_t us, const float *__restrict__ itab, int64_t is,
                    const int64_t *__restrict__ users, int64_t n_b, int64_t n_item, int dim,
                    const int64_t *__restrict__ mask_dense, const int32_t *__restrict__ trn_rowptr,
                    const int32_t *__restrict__ trn_cols, float *preds) {
    __shared__ float a_s[TK][TM + PAD];
    __shared__ float b_s[TK][TN + PAD];
    __shared__ int64_t u_s[TM];
    const int tid = (int)threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
    if (tid < TM) u_s[tid] = (m0 + tid < n_b) ? users[m0 + tid] : (int64_t)-1;
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < dim; k0 += TK) {
        // a warp loads 32 consecutive k of one row (128 contiguous bytes) and stores them down a padded column
#pragma unroll 4
        for (int q = 0; q < TM * TK / NT; ++q) {
            const int idx = tid + NT * q;
            const int r = idx / TK, k = idx % TK;
            const int64_t u = u_s[r];
            a_s[k][r] = (u >= 0 && k0 + k < dim) ? ut[u * us + k0 + k] : 0.f;
        }
#pragma unroll 4
        for (int q = 0; q < TN * TK / NT; ++q) {
            const int idx = tid + NT * q;
            const int r = idx / TK, k = idx % TK;
            const int64_t it = n0 + r;
            b_s[k][r] = (it < n_item && k0 + k < dim) ? itab[it * is + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < TK; ++k) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = a_s[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = b_s[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t b = m0 + ty + 16 * i;
        if (b >= n_b) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t it = n0 + tx + 16 * j;
            if (it >= n_item) continue;
            float m = 0.f;
            if (mask_dense) m = (float)mask_dense[b * n_item + it];
            preds[b * n_item + it] = acc[i][j] * (1.f - m) - 1e8f * m;          // base_model.py:36
        }
    }

    if (mask_dense == nullptr && trn_rowptr != nullptr) {
        __syncthreads();      // the tile's scores are written and visible to the whole CTA
        const int lane = tid & 31, warp = tid >> 5;
        const int64_t n1 = (n0 + TN < n_item) ? n0 + TN : n_item;
        for (int r = warp; r < TM; r += NT / 32) {
            const int64_t u = u_s[r];
            if (u < 0) continue;
            float *row = preds + (m0 + r) * n_item;
            const int e1 = trn_rowptr[u + 1];
            for (int e = trn_rowptr[u] + lane; e < e1; e += 32) {
                const int64_t it = trn_cols[e];
                if (it >= n0 && it < n1) row[it] = row[it] * 0.f - 1e8f;      // s * (1 - 1) - 1e8 * 1
            }
        }
    }
}

}  // namespace ssl_predict
