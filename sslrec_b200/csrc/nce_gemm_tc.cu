// ssl_softmax_gemm_tf32x3: the InfoNCE contraction on the 5th-generation tensor cores (tcgen05,
// kind::tf32) with fp32-grade accuracy through 3xTF32 error compensation.
//
//   S  = R C^T   = R_hi C_hi^T + R_lo C_hi^T + R_hi C_lo^T          (x = x_hi + x_lo, x_hi = tf32(x))
//   E  = exp2(S - offset) * colscale ;  rowsum += sum_c E
//   O += E C     = E_hi C_hi   + E_lo C_hi   + E_hi C_lo
//
// Same contract as ssl_softmax_gemm (nce_gemm.cu): one launch is the forward of an InfoNCE term or,
// with the operand roles swapped, its backward.  Both parts are rounded to nearest (cvt.rna.tf32), so the
// dropped lo*lo products and the rounding of the lo parts are O(2^-22) relative and unbiased -- the fp32
// rounding level of the FFMA kernel.
//
// Structure (one CTA per SM, 256 threads, warp-specialised, all synchronisation by mbarriers):
//   warp 0 / lane 0 : TMA producer, ring 1.  2-D tensor maps (SWIZZLE_128B, 32-float boxes) over the
//                     row-major hi / lo operand arrays: the resident 128-row R tile once, then the 64-row
//                     C tiles (GEMM1's B operand, K-major; freed as soon as GEMM1 of the tile retires).
//   warp 2 / lane 0 : TMA producer, ring 2: the same tiles from the TRANSPOSED copies [d, n] (GEMM2's B
//                     operand, K-major again).  tf32 operands must be K-major here: an MN-major view of
//                     the row-major tile needs the 32-byte-atom swizzle, which the K-major GEMM1 view of
//                     the same bytes cannot share (measured: the MN-major descriptor yields zeros).
//   warp 1 / lane 0 : MMA issuer 1.  GEMM1 (M=128, N=64, K=d): both operands in shared memory -> S in TMEM.
//   warp 3 / lane 0 : MMA issuer 2.  GEMM2 (M=128, N=d, K=64): A = E read from TENSOR MEMORY, B = C^T tile ->
//                     O in TMEM, accumulated over all tiles.  Two issuing threads because tf32 MMAs are only
//                     K = 8 deep: one thread cannot issue them as fast as the tensor pipe retires them.
//   warps 4-11,12-19: two epilogue groups (256 threads each) ping-ponging over the tiles (parity); thread =
//                     (TMEM lane = row, one 32-column half of the tile).  tcgen05.ld S (which frees the S
//                     buffer for the next GEMM1 at once), ex2, row sums in registers, tf32 split of E,
//                     tcgen05.st E_hi / E_lo into their own TMEM buffers; finally O is read out once per CTA.
//   Roofline: 65536 MACs per K=8 MMA at the measured tf32 peak (cuBLAS bf16 burst / 2 = 865 TFLOP/s) is ~44 SM cycles,
//   48 MMAs per 64-column tile = ~2100 cycles; measured 2350 cycles per tile at the bench's forward shape
//   (0.339 ms, 777 TFLOP/s of tf32 MMA work = 90 % of that peak).  Tried and dropped: the resident operand's
//   hi part in TMEM (TS-form GEMM1, single S buffer) -- same speed, slower backward shape.
//   TMEM columns    : [0,128) S x2, [128,256) E_hi x2, [256,384) E_lo x2, [384,384+d) O (hi*hi),
//                     [448,448+d) O correction terms -- all 512 columns.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64;
constexpr int ST1 = 3, ST2 = 2;        // stages of ring 1 (GEMM1's B, gates the S pipeline) and ring 2 (GEMM2's B); 64 + 96 + 64 KB at d = 64
constexpr int kNumThreads = 640;       // warps 0-3: TMA, MMA1, TMA, MMA2; warps 4-11 and 12-19: two epilogue groups of 256 threads
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t COL_S = 0, COL_EHI = 128, COL_ELO = 256, COL_O = 384, COL_OC = 448;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// exactly one lane of a fully converged warp (warp-uniform control flow around it keeps the operands of
// the MMA / commit instructions in uniform registers: no per-lane "waterfall" loop around every UTCHMMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Four consecutive K-steps (one 128-byte swizzle chunk = 32 tf32) in ONE asm block: the descriptors advance by
// 32 bytes (+2 in the 16-byte-unit address field) inside the block, so the single issuing thread spends ~3
// instructions per MMA instead of ~15 (ncu round 1: the contraction was bound by the MMA issue loop, not by the
// tensor pipe).  acc_first: whether the first of the four accumulates onto D; the other three always do.
__device__ __forceinline__ void mma_ss_x4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc_first) {
    asm volatile(
        "{\n"
        ".reg .pred p, t;\n"
        ".reg .b64 a, b;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "mov.b64 a, %1;\n"
        "mov.b64 b, %2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %3, p;\n"
        "add.s64 a, a, 2;\n add.s64 b, b, 2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %3, t;\n"
        "add.s64 a, a, 2;\n add.s64 b, b, 2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %3, t;\n"
        "add.s64 a, a, 2;\n add.s64 b, b, 2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %3, t;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc_first) : "memory");
}
__device__ __forceinline__ void mma_ts_x4(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc_first) {
    asm volatile(
        "{\n"
        ".reg .pred p, t;\n"
        ".reg .b64 b;\n"
        ".reg .b32 a;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "mov.b32 a, %1;\n"
        "mov.b64 b, %2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [a], b, %3, p;\n"
        "add.s32 a, a, 8;\n add.s64 b, b, 2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [a], b, %3, t;\n"
        "add.s32 a, a, 8;\n add.s64 b, b, 2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [a], b, %3, t;\n"
        "add.s32 a, a, 8;\n add.s64 b, b, 2;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [a], b, %3, t;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc_first) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32
__host__ __device__ constexpr uint32_t instr_desc(int m, int n, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// one 32-column chunk of a tile: S -> E, row sum, tf32 split
template <bool CHECK>
__device__ __forceinline__ void exp_chunk(uint32_t (&v)[32], uint32_t (&lo)[32], float offset, const float *__restrict__ cs_ptr,
                                          int64_t col, int64_t n_c, float &rowsum) {
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
        float cs[4] = {1.f, 1.f, 1.f, 1.f};
        if (cs_ptr != nullptr) {
            const float4 c4 = __ldg(reinterpret_cast<const float4 *>(cs_ptr + col + k));
            cs[0] = c4.x; cs[1] = c4.y; cs[2] = c4.z; cs[3] = c4.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float e = ex2(__uint_as_float(v[k + u]) - offset) * cs[u];
            if (CHECK) e = (col + k + u < n_c) ? e : 0.f;
            rowsum += e;
            float ehi, elo;
            ssl::tf32_split(e, ehi, elo);
            v[k + u] = __float_as_uint(ehi);
            lo[k + u] = __float_as_uint(elo);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(kNumThreads, 1)
softmax_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
                       const __grid_constant__ CUtensorMap map_c_hi, const __grid_constant__ CUtensorMap map_c_lo,
                       const __grid_constant__ CUtensorMap map_ct_hi, const __grid_constant__ CUtensorMap map_ct_lo,
                       int64_t n_r, int64_t n_c, const float *__restrict__ colscale, float offset, int n_split,
                       float *__restrict__ rowsum_part, float *__restrict__ o_part) {
    constexpr int KCH = D / 32;                         // 128-byte K chunks per operand row (GEMM1: K = d)
    constexpr int JCH = BN / 32;                        // 128-byte K chunks of the transposed tile (GEMM2: K = 64 rows)
    constexpr uint32_t R_CHUNK = BM * 128, C_CHUNK = BN * 128, T_CHUNK = D * 128;
    constexpr uint32_t R_BYTES = KCH * R_CHUNK, C_BYTES = KCH * C_CHUNK, T_BYTES = JCH * T_CHUNK;   // one precision part
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *r_hi = smem, *r_lo = smem + R_BYTES;
    uint8_t *ring1 = smem + 2 * R_BYTES;                // stage s: C tile row-major, hi then lo (GEMM1's B, K-major)
    uint8_t *ring2 = ring1 + ST1 * 2 * C_BYTES;      // stage s: C^T tile, hi then lo          (GEMM2's B, K-major)
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring2 + ST2 * 2 * T_BYTES);
    uint64_t *full1 = bars, *empty1 = full1 + ST1, *full2 = empty1 + ST1, *empty2 = full2 + ST2;
    uint64_t *s_full = empty2 + ST2, *s_free = s_full + 2, *e_ready = s_free + 2, *e_free = e_ready + 2, *r_full = e_free + 2, *o_full = r_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_full + 1);
    float *rowsum_x = reinterpret_cast<float *>(tmem_slot + 4);     // [3][128] partial row sums of the other epilogue sub-groups

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rt = blockIdx.x / n_split, sp = blockIdx.x % n_split;
    const int64_t n_ct = (n_c + BN - 1) / BN;
    const int t0 = (int)(n_ct * sp / n_split), t1 = (int)(n_ct * (sp + 1) / n_split);
    const int n_tiles = t1 - t0;
    const int row0 = rt * BM;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r_lo));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c_lo));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ct_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ct_lo));
        for (int s = 0; s < ST1; ++s) {
            mbar_init(&full1[s], 1);
            mbar_init(&empty1[s], 1);
        }
        for (int s = 0; s < ST2; ++s) {
            mbar_init(&full2[s], 1);
            mbar_init(&empty2[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&s_free[b], 256);
            mbar_init(&e_ready[b], 256);
            mbar_init(&e_free[b], 1);
        }
        mbar_init(r_full, 1);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer, ring 1: R tile once, then the row-major C tiles =====================
        mbar_expect_tx(r_full, 2 * R_BYTES);
        for (int c = 0; c < KCH; ++c) {
            tma_load_2d(r_hi + c * R_CHUNK, &map_r_hi, c * 32, row0, r_full);
            tma_load_2d(r_lo + c * R_CHUNK, &map_r_lo, c * 32, row0, r_full);
        }
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % ST1;
            mbar_wait(&empty1[s], ((i / ST1) & 1) ^ 1);
            uint8_t *hi = ring1 + s * 2 * C_BYTES, *lo = hi + C_BYTES;
            mbar_expect_tx(&full1[s], 2 * C_BYTES);
            for (int c = 0; c < KCH; ++c) {
                tma_load_2d(hi + c * C_CHUNK, &map_c_hi, c * 32, (t0 + i) * BN, &full1[s]);
                tma_load_2d(lo + c * C_CHUNK, &map_c_lo, c * 32, (t0 + i) * BN, &full1[s]);
            }
        }
    } else if (warp == 2 && lane == 0) {
        // ===================== TMA producer, ring 2: the transposed C tiles [d, 64] =====================
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % ST2;
            mbar_wait(&empty2[s], ((i / ST2) & 1) ^ 1);
            uint8_t *hi = ring2 + s * 2 * T_BYTES, *lo = hi + T_BYTES;
            mbar_expect_tx(&full2[s], 2 * T_BYTES);
            for (int c = 0; c < JCH; ++c) {
                tma_load_2d(hi + c * T_CHUNK, &map_ct_hi, (t0 + i) * BN + c * 32, 0, &full2[s]);
                tma_load_2d(lo + c * T_CHUNK, &map_ct_lo, (t0 + i) * BN + c * 32, 0, &full2[s]);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer 1: S = R C^T (three tf32 products, small ones first) =====================
        constexpr uint32_t idesc1 = instr_desc(BM, BN, 0);      // S[128 x 64] = R (K-major, K = d) x C (K-major)
        const uint32_t r_hi_a = smem_u32(r_hi), r_lo_a = smem_u32(r_lo);
        mbar_wait(r_full, 0);
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % ST1, b = i & 1;
            mbar_wait(&full1[s], (i / ST1) & 1);
            mbar_wait(&s_free[b], ((i >> 1) & 1) ^ 1);           // the epilogue of tile i-2 has read S out of this buffer
            tc_fence_after();
            const uint32_t c_hi_a = smem_u32(ring1 + s * 2 * C_BYTES), c_lo_a = c_hi_a + C_BYTES;
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const uint32_t ra = (part == 0) ? r_lo_a : r_hi_a;
                const uint32_t cb = (part == 1) ? c_lo_a : c_hi_a;
#pragma unroll
                for (int c = 0; c < KCH; ++c)
                    if (elect_one())
                        mma_ss_x4(tmem + COL_S + b * BN, smem_desc(ra + c * R_CHUNK, 16, 1024), smem_desc(cb + c * C_CHUNK, 16, 1024), idesc1,
                                  (part > 0 || c > 0) ? 1u : 0u);
            }
            if (elect_one()) {
                tc_commit(&s_full[b]);
                tc_commit(&empty1[s]);                          // GEMM1 was the only reader of this stage
            }
            __syncwarp();
        }
    } else if (warp == 3) {
        // ===================== MMA issuer 2: O += E C  (A = E from TMEM, B = C^T tile) =====================
        constexpr uint32_t idesc2 = instr_desc(BM, D, 0);       // O[128 x d] += E (TMEM, K = 64) x C^T (K-major)
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j % ST2, b = j & 1;
            mbar_wait(&e_ready[b], (j >> 1) & 1);
            mbar_wait(&full2[s], (j / ST2) & 1);
            tc_fence_after();
            const uint32_t t_hi_a = smem_u32(ring2 + s * 2 * T_BYTES), t_lo_a = t_hi_a + T_BYTES;
            const uint32_t e_hi = tmem + COL_EHI + b * BN, e_lo = tmem + COL_ELO + b * BN;
            // the two correction products go to their own accumulator: the tensor core rounds its fp32
            // accumulations toward zero, so the long hi*hi sum must not also carry them
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const uint32_t ea = (part == 0) ? e_lo : e_hi;
                const uint32_t tb = (part == 1) ? t_lo_a : t_hi_a;
                const uint32_t od = tmem + ((part == 2) ? COL_O : COL_OC);
#pragma unroll
                for (int c = 0; c < JCH; ++c) {
                    const bool first = (part == 2) ? (c == 0) : (part == 0 && c == 0);
                    if (elect_one()) mma_ts_x4(od, ea + c * 32, smem_desc(tb + c * T_CHUNK, 16, 1024), idesc2, (j > 0 || !first) ? 1u : 0u);
                }
            }
            if (elect_one()) {
                tc_commit(&empty2[s]);
                tc_commit(&e_free[b]);
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(o_full);
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue: two groups of 8 warps ping-pong over the tiles (group g owns the tiles and TMEM buffers of
        // ===== parity g); inside a group, thread = (TMEM lane = row, one 32-column half of the 64-column tile)
        const int e = warp - 4;
        const int g = e >> 3, half = (e >> 2) & 1, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        float rowsum = 0.f;
        for (int i = g; i < n_tiles; i += 2) {
            const int b = g;
            const uint32_t par = (i >> 1) & 1;
            mbar_wait(&s_full[b], par);
            tc_fence_after();
            uint32_t v[32], lo[32];
            tmem_ld32(lane_base + COL_S + b * BN + half * 32, v);
            tc_fence_before();
            mbar_arrive(&s_free[b]);                             // GEMM1 of tile i+2 may overwrite S now
            const int64_t col = (int64_t)(t0 + i) * BN + half * 32;
            if (col + 32 <= n_c) exp_chunk<false>(v, lo, offset, colscale, col, n_c, rowsum);
            else exp_chunk<true>(v, lo, offset, colscale, col, n_c, rowsum);
            mbar_wait(&e_free[b], par ^ 1);                      // GEMM2 of tile i-2 has consumed the previous E
            tc_fence_after();
            tmem_st32(lane_base + COL_EHI + b * BN + half * 32, v);
            tmem_st32(lane_base + COL_ELO + b * BN + half * 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&e_ready[b]);
        }
        // ---- combine the four sub-groups' row sums; group 0 reads O out of TMEM once per CTA ----
        const int sub = g * 2 + half;
        if (sub > 0) rowsum_x[(sub - 1) * 128 + row] = rowsum;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (g == 0) {
            mbar_wait(o_full, 0);
            tc_fence_after();
            const int64_t grow = (int64_t)row0 + row;
            if (half < D / 32) {
                if (n_tiles > 0) {
                    uint32_t v[32], c[32];
                    tmem_ld32(lane_base + COL_O + half * 32, v);
                    tmem_ld32(lane_base + COL_OC + half * 32, c);
                    if (grow < n_r) {
                        float4 *dst = reinterpret_cast<float4 *>(o_part + ((size_t)sp * n_r + grow) * D + half * 32);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            dst[k] = make_float4(__uint_as_float(v[4 * k]) + __uint_as_float(c[4 * k]),
                                                 __uint_as_float(v[4 * k + 1]) + __uint_as_float(c[4 * k + 1]),
                                                 __uint_as_float(v[4 * k + 2]) + __uint_as_float(c[4 * k + 2]),
                                                 __uint_as_float(v[4 * k + 3]) + __uint_as_float(c[4 * k + 3]));
                    }
                } else if (grow < n_r) {
                    for (int k = 0; k < 32; ++k) o_part[((size_t)sp * n_r + grow) * D + half * 32 + k] = 0.f;
                }
            }
            if (half == 0 && grow < n_r && rowsum_part != nullptr)
                rowsum_part[(size_t)sp * n_r + grow] = rowsum + rowsum_x[row] + rowsum_x[128 + row] + rowsum_x[256 + row];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, cols] fp32, row pitch ``pitch`` floats -> boxes of 32 floats (128 B, SWIZZLE_128B) x box_rows rows;
// out-of-range rows / columns read as 0
int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int64_t pitch, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        ssl::set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SSL_E_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ssl::set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return SSL_E_CUDA;
    }
    return SSL_OK;
}

template <int D>
int launch_tc(const float *R_hi, const float *R_lo, int64_t n_r, const float *C_hi, const float *C_lo, const float *CT_hi,
              const float *CT_lo, int64_t ct_pitch, int64_t n_c, const float *colscale, float offset, int n_split,
              float *rowsum_part, float *o_part, cudaStream_t st) {
    CUtensorMap mr_hi, mr_lo, mc_hi, mc_lo, mt_hi, mt_lo;
    int rc;
    if ((rc = make_map(&mr_hi, R_hi, n_r, D, D, BM)) != SSL_OK) return rc;
    if ((rc = make_map(&mr_lo, R_lo, n_r, D, D, BM)) != SSL_OK) return rc;
    if ((rc = make_map(&mc_hi, C_hi, n_c, D, D, BN)) != SSL_OK) return rc;
    if ((rc = make_map(&mc_lo, C_lo, n_c, D, D, BN)) != SSL_OK) return rc;
    if ((rc = make_map(&mt_hi, CT_hi, D, n_c, ct_pitch, D)) != SSL_OK) return rc;
    if ((rc = make_map(&mt_lo, CT_lo, D, n_c, ct_pitch, D)) != SSL_OK) return rc;
    constexpr int KCH = D / 32;
    const size_t smem = 1024 + 2 * (size_t)KCH * BM * 128 + (size_t)ST1 * 2 * KCH * BN * 128 + (size_t)ST2 * 2 * (BN / 32) * D * 128 +
                        32 * sizeof(uint64_t) + 16 + 3 * 128 * sizeof(float);
    // cudaFuncSetAttribute is per DEVICE: remember which devices of this process are configured
    static bool configured[64] = {};
    int dev = 0;
    SSL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        SSL_CUDA(cudaFuncSetAttribute(softmax_gemm_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int64_t grid = ((n_r + BM - 1) / BM) * n_split;
    softmax_gemm_tc_kernel<D><<<(unsigned)grid, kNumThreads, smem, st>>>(mr_hi, mr_lo, mc_hi, mc_lo, mt_hi, mt_lo, n_r, n_c, colscale, offset,
                                                                         n_split, rowsum_part, o_part);
    SSL_LAUNCH_CHECK("softmax_gemm_tc_kernel");
    return SSL_OK;
}

}  // namespace

extern "C" int ssl_softmax_gemm_tf32x3(const float *R_hi, const float *R_lo, int64_t n_r, const float *C_hi, const float *C_lo,
                                       const float *CT_hi, const float *CT_lo, int64_t ct_pitch, int64_t n_c, int32_t dim,
                                       const float *colscale, float offset, int32_t n_split, float *rowsum_part, float *o_part,
                                       void *stream) {
    SSL_CHECK_ARG(R_hi && R_lo && C_hi && C_lo && CT_hi && CT_lo && o_part, "ssl_softmax_gemm_tf32x3: null argument");
    SSL_CHECK_ARG(ct_pitch >= n_c && ct_pitch % 4 == 0, "ssl_softmax_gemm_tf32x3: ct_pitch must be >= n_c and a multiple of 4");
    SSL_CHECK_ARG(dim == 32 || dim == 64, "ssl_softmax_gemm_tf32x3: dim %d not supported (32 or 64; other sizes use ssl_softmax_gemm)", dim);
    SSL_CHECK_ARG((n_split >= 1 && n_split <= (n_c + BN - 1) / BN) || n_c == 0, "ssl_softmax_gemm_tf32x3: n_split %d exceeds the number of C tiles", n_split);
    SSL_CHECK_ARG(((reinterpret_cast<uintptr_t>(R_hi) | reinterpret_cast<uintptr_t>(R_lo) | reinterpret_cast<uintptr_t>(C_hi) |
                    reinterpret_cast<uintptr_t>(C_lo) | reinterpret_cast<uintptr_t>(CT_hi) | reinterpret_cast<uintptr_t>(CT_lo) |
                    reinterpret_cast<uintptr_t>(o_part)) & 15) == 0,
                  "ssl_softmax_gemm_tf32x3: operands must be 16-byte aligned");
    if (n_r == 0 || n_c == 0) return SSL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dim == 32) return launch_tc<32>(R_hi, R_lo, n_r, C_hi, C_lo, CT_hi, CT_lo, ct_pitch, n_c, colscale, offset, n_split, rowsum_part, o_part, st);
    return launch_tc<64>(R_hi, R_lo, n_r, C_hi, C_lo, CT_hi, CT_lo, ct_pitch, n_c, colscale, offset, n_split, rowsum_part, o_part, st);
}
