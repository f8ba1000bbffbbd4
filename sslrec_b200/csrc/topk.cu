// ssl_topk: the k largest scores of every prediction row (metrics.py:108 torch.topk(k = max(k))).
// One block per row: 4-pass MSB radix select on an order-preserving integer key finds the k-th
// value exactly, one ordered pass collects everything above it plus the lowest-index ties, and a
// bitonic sort of the <= 256 survivors orders them (value descending, index ascending).  Integer
// selection -> the result is exact and deterministic; ties resolve to the lower item id.
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kMaxK = 256;

__device__ __forceinline__ uint32_t okey(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(kThreads) topk_kernel(const float *__restrict__ preds, int64_t n_item, int k,
                                                        int64_t *__restrict__ out_idx, float *__restrict__ out_val) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_need, s_ngt, s_eq_base;
    __shared__ uint32_t warp_cnt[kThreads / 32];
    __shared__ unsigned long long cand[kMaxK];
    const float *row = preds + (size_t)blockIdx.x * n_item;
    const int tid = threadIdx.x;

    if (tid == 0) { s_prefix = 0; s_need = (uint32_t)k; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kThreads) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t hi_mask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
        for (int64_t i = tid; i < n_item; i += kThreads) {
            const uint32_t key = okey(row[i]);
            if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t need = s_need, acc = 0;
            int d = 255;
            for (; d > 0; --d) {
                if (acc + hist[d] >= need) break;
                acc += hist[d];
            }
            s_need = need - acc;
            s_prefix = prefix | ((uint32_t)d << shift);
        }
        __syncthreads();
    }
    const uint32_t thr = s_prefix;          // key of the k-th largest element
    const uint32_t need_eq = s_need;        // how many elements equal to it belong to the top-k
    if (tid == 0) { s_ngt = 0; s_eq_base = 0; }
    for (int i = tid; i < kMaxK; i += kThreads) cand[i] = 0ull;
    __syncthreads();
    const uint32_t n_gt = (uint32_t)k - need_eq;
    for (int64_t base = 0; base < n_item; base += kThreads) {
        const int64_t i = base + tid;
        uint32_t key = 0;
        bool gt = false, eq = false;
        if (i < n_item) {
            key = okey(row[i]);
            gt = key > thr;
            eq = key == thr;
        }
        if (gt) {
            const uint32_t slot = atomicAdd(&s_ngt, 1u);
            cand[slot] = ((unsigned long long)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)i);
        }
        // ordered rank among the ties
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        if ((tid & 31) == 0) warp_cnt[tid >> 5] = __popc(bal);
        __syncthreads();
        uint32_t before = s_eq_base;
        for (int w = 0; w < (tid >> 5); ++w) before += warp_cnt[w];
        const uint32_t rank = before + __popc(bal & ((1u << (tid & 31)) - 1u));
        if (eq && rank < need_eq) cand[n_gt + rank] = ((unsigned long long)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)i);
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kThreads / 32; ++w) t += warp_cnt[w];
            s_eq_base += t;
        }
        __syncthreads();
    }
    __syncthreads();
    // bitonic sort, descending, on the next power of two >= k (unused slots are 0 = smallest)
    int P = 1;
    while (P < k) P <<= 1;
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < P; i += kThreads) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const unsigned long long a = cand[i], b = cand[j];
                    if ((a < b) == desc) { cand[i] = b; cand[j] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += kThreads) {
        const unsigned long long c = cand[i];
        out_idx[(size_t)blockIdx.x * k + i] = (int64_t)(0xffffffffu - (uint32_t)(c & 0xffffffffull));
        if (out_val) out_val[(size_t)blockIdx.x * k + i] = okey_inv((uint32_t)(c >> 32));
    }
}

}  // namespace

extern "C" int ssl_topk(const float *preds, int64_t n_b, int64_t n_item, int32_t k, int64_t *out_idx, float *out_val, void *stream) {
    SSL_CHECK_ARG(preds && out_idx, "ssl_topk: null argument");
    SSL_CHECK_ARG(k >= 1 && k <= kMaxK && k <= n_item, "ssl_topk: k = %d must be in [1, min(%d, n_item)]", k, kMaxK);
    SSL_CHECK_ARG(n_item < (int64_t)0xffffffff, "ssl_topk: too many items");
    if (n_b == 0) return SSL_OK;
    topk_kernel<<<(unsigned)n_b, kThreads, 0, (cudaStream_t)stream>>>(preds, n_item, k, out_idx, out_val);
    SSL_LAUNCH_CHECK("topk_kernel");
    return SSL_OK;
}
