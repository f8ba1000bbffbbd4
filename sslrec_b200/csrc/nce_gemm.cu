// ssl_softmax_gemm: the InfoNCE contraction without the [B, N_side] logits.
//
//   for every row r of R (resident tile of 128 rows) and every row c of C (streamed, 64 per tile):
//       e = exp2(R_r . C_c - offset) * colscale[c];   rowsum[r] += e;   O[r, :] += e * C_c
//
// Forward of a term : R = anchors a^ * log2e/tau, C = normalised table  -> log-sum-exp pieces and
//                     the softmax-weighted table average (the anchor gradient).
// Backward of a term: R = table tile, C = anchors, colscale = g*ln2/rowsum -> the dense table
//                     gradient.  |cos| <= 1 so offset = log2e/tau bounds every exponent by 0: no
//                     running max / rescale is needed (the reference subtracts no max either,
//                     loss_utils.py:37, it just overflows where this does not).
//
// This is a dense fp32 contraction (2 x 2*B*N*d flop per launch) executed on the FP32 FMA pipe:
// BASELINE.json keeps tensor cores off this path, and the 1e-5 loss tolerance excludes plain
// tf32.  Roofline: FP32 FMA throughput, not HBM.
//
// Tiling.  256 threads = 16 (tx) x 16 (ty).  GEMM1: S[128 x 64] = R_tile C_tile^T with both
// operands K-major in shared memory (R transposed once per CTA; C arrives already K-major from
// the producer's tile copy), 8 x 4 outputs per thread, 3 LDS.128 per 32 FFMA.  The exponentials
// are written transposed (E_T[c][r], pitch 132 -> conflict-free STS.128) and GEMM2:
// O[128 x dim] += E C_tile reads E_T and the row-major C tile, again 3 LDS.128 per 32 FFMA.
// The two C copies are fetched with 1-D bulk async copies (TMA engine, cp.async.bulk +
// mbarrier complete_tx); the K-major copy of tile t+1 lands while GEMM2 of tile t runs and the
// row-major copy while GEMM1 of tile t+1 runs, so one buffer each suffices and two CTAs fit
// per SM at dim <= 64.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, EPITCH = BM + 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
__global__ void __launch_bounds__(256, (D <= 64) ? 2 : 1)
softmax_gemm_kernel(const float *__restrict__ R, int64_t n_r, const float *__restrict__ C, const float *__restrict__ C_t,
                    int64_t n_c, int dim, const float *__restrict__ colscale, float offset, int n_split,
                    float *__restrict__ rowsum_part, float *__restrict__ o_part) {
    constexpr int CPT = D / 16;
    extern __shared__ __align__(128) float smem[];
    float *Rs_T = smem;                 // [D][BM]
    float *Cs_T = Rs_T + D * BM;        // [D][BN]   K-major, permuted columns
    float *Cs = Cs_T + D * BN;          // [BN][dim] row-major
    float *E_T = Cs + BN * D;           // [BN][EPITCH]
    uint64_t *bars = reinterpret_cast<uint64_t *>(E_T + BN * EPITCH);

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int rt = blockIdx.x / n_split, sp = blockIdx.x % n_split;
    const int64_t n_ct = (n_c + BN - 1) / BN;
    const int64_t t0 = n_ct * sp / n_split, t1 = n_ct * (sp + 1) / n_split;
    const int64_t row0 = (int64_t)rt * BM;

    for (int i = tid; i < D * BM + D * BN + BN * D; i += 256) smem[i] = 0.f;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t tile_bytes = (uint32_t)(BN * dim * sizeof(float));
    if (tid == 0 && t0 < t1) {
        mbar_expect_tx(&bars[0], tile_bytes);
        bulk_g2s(Cs_T, C_t + (size_t)t0 * dim * BN, tile_bytes, &bars[0]);
        mbar_expect_tx(&bars[1], tile_bytes);
        bulk_g2s(Cs, C + (size_t)t0 * BN * dim, tile_bytes, &bars[1]);
    }
    // resident tile, transposed to K-major (once per CTA)
    {
        const int quads = dim >> 2;
        for (int i = tid; i < BM * quads; i += 256) {
            const int r = i / quads, q = i % quads;
            if (row0 + r < n_r) {
                const float4 v = ssl::ldg4(R + (row0 + r) * dim + q * 4);
                Rs_T[(q * 4 + 0) * BM + r] = v.x;
                Rs_T[(q * 4 + 1) * BM + r] = v.y;
                Rs_T[(q * 4 + 2) * BM + r] = v.z;
                Rs_T[(q * 4 + 3) * BM + r] = v.w;
            }
        }
    }
    __syncthreads();

    float o[8][CPT];
    float rowsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        rowsum[i] = 0.f;
#pragma unroll
        for (int c = 0; c < CPT; ++c) o[i][c] = 0.f;
    }

    uint32_t parity = 0;
    for (int64_t t = t0; t < t1; ++t, parity ^= 1) {
        // ---------------- GEMM1: S = R_tile . C_tile^T ----------------
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
        mbar_wait(&bars[0], parity);
#pragma unroll 8
        for (int k = 0; k < dim; ++k) {
            const float4 r0 = *reinterpret_cast<const float4 *>(Rs_T + k * BM + ty * 8);
            const float4 r1 = *reinterpret_cast<const float4 *>(Rs_T + k * BM + ty * 8 + 4);
            const float4 c = *reinterpret_cast<const float4 *>(Cs_T + k * BN + tx * 4);
            const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(rr[i], cc[j], s[i][j]);
        }
        __syncthreads();   // everyone is done with Cs_T (and with E_T / Cs of the previous tile)
        if (tid == 0 && t + 1 < t1) {
            mbar_expect_tx(&bars[0], tile_bytes);
            bulk_g2s(Cs_T, C_t + (size_t)(t + 1) * dim * BN, tile_bytes, &bars[0]);
        }
        // ---------------- exponentials, row sums, transposed store ----------------
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cl = tx + 16 * j;                 // logical column of physical slot 4*tx + j
            const int64_t cg = t * BN + cl;
            const bool valid = cg < n_c;
            const float cs = valid ? (colscale ? __ldg(colscale + cg) : 1.f) : 0.f;
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                e[i] = valid ? ex2(s[i][j] - offset) * cs : 0.f;
                rowsum[i] += e[i];
            }
            *reinterpret_cast<float4 *>(E_T + cl * EPITCH + ty * 8) = make_float4(e[0], e[1], e[2], e[3]);
            *reinterpret_cast<float4 *>(E_T + cl * EPITCH + ty * 8 + 4) = make_float4(e[4], e[5], e[6], e[7]);
        }
        __syncthreads();   // E_T complete
        // ---------------- GEMM2: O += E . C_tile ----------------
        mbar_wait(&bars[1], parity);
#pragma unroll 4
        for (int j = 0; j < BN; ++j) {
            const float4 e0 = *reinterpret_cast<const float4 *>(E_T + j * EPITCH + ty * 8);
            const float4 e1 = *reinterpret_cast<const float4 *>(E_T + j * EPITCH + ty * 8 + 4);
            const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
            float cv[CPT];
            const float *cp = Cs + j * dim + tx * CPT;
            if constexpr (CPT == 2) {
                const float2 v = *reinterpret_cast<const float2 *>(cp);
                cv[0] = v.x; cv[1] = v.y;
            } else {
#pragma unroll
                for (int q = 0; q < CPT / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4 *>(cp + q * 4);
                    cv[q * 4 + 0] = v.x; cv[q * 4 + 1] = v.y; cv[q * 4 + 2] = v.z; cv[q * 4 + 3] = v.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < CPT; ++c) o[i][c] = fmaf(ee[i], cv[c], o[i][c]);
        }
        __syncthreads();   // everyone is done with Cs and E_T
        if (tid == 0 && t + 1 < t1) {
            mbar_expect_tx(&bars[1], tile_bytes);
            bulk_g2s(Cs, C + (size_t)(t + 1) * BN * dim, tile_bytes, &bars[1]);
        }
    }

    // ---------------- write this split's partials ----------------
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = rowsum[i];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const int64_t row = row0 + ty * 8 + i;
        if (row < n_r) {
            if (tx == 0 && rowsum_part != nullptr) rowsum_part[(size_t)sp * n_r + row] = v;
            float *dst = o_part + ((size_t)sp * n_r + row) * dim + tx * CPT;
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                if (tx * CPT + c < dim) dst[c] = o[i][c];
        }
    }
}

template <int D>
int launch(const float *R, int64_t n_r, const float *C, const float *C_t, int64_t n_c, int dim, const float *colscale,
           float offset, int n_split, float *rowsum_part, float *o_part, cudaStream_t st) {
    const size_t smem = sizeof(float) * (D * BM + D * BN + BN * D + BN * EPITCH) + 2 * sizeof(uint64_t);
    static bool configured[64] = {};      // cudaFuncSetAttribute is per device
    int dev = 0;
    SSL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        SSL_CUDA(cudaFuncSetAttribute(softmax_gemm_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int64_t grid = ((n_r + BM - 1) / BM) * n_split;
    softmax_gemm_kernel<D><<<(unsigned)grid, 256, smem, st>>>(R, n_r, C, C_t, n_c, dim, colscale, offset, n_split, rowsum_part, o_part);
    SSL_LAUNCH_CHECK("softmax_gemm_kernel");
    return SSL_OK;
}

}  // namespace

extern "C" int ssl_softmax_gemm(const float *R, int64_t n_r, const float *C, const float *C_t, int64_t n_c, int32_t dim,
                                const float *colscale, float offset, int32_t n_split, float *rowsum_part, float *o_part,
                                void *stream) {
    SSL_CHECK_ARG(R && C && C_t && o_part, "ssl_softmax_gemm: null argument");
    SSL_CHECK_ARG(dim >= 4 && dim <= SSL_MAX_DIM && dim % 4 == 0, "ssl_softmax_gemm: dim %d must be a multiple of 4 <= %d", dim, SSL_MAX_DIM);
    SSL_CHECK_ARG((n_split >= 1 && n_split <= (n_c + BN - 1) / BN) || n_c == 0, "ssl_softmax_gemm: n_split %d exceeds the number of C tiles", n_split);
    SSL_CHECK_ARG(((reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(C_t)) & 15) == 0,
                  "ssl_softmax_gemm: operands must be 16-byte aligned");
    if (n_r == 0 || n_c == 0) return SSL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dim <= 32) return launch<32>(R, n_r, C, C_t, n_c, dim, colscale, offset, n_split, rowsum_part, o_part, st);
    if (dim <= 64) return launch<64>(R, n_r, C, C_t, n_c, dim, colscale, offset, n_split, rowsum_part, o_part, st);
    return launch<128>(R, n_r, C, C_t, n_c, dim, colscale, offset, n_split, rowsum_part, o_part, st);
}
