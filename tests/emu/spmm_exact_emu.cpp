// Host execution of sslrec_b200/csrc/spmm_exact.cuh against the sequential FMA chain over each CSR row (= what torch's CPU t.spmm computes on
// the reference's adjacency: tests/test_host_emulation.py).  usage: spmm_exact_emu n_rows n_cols dim x_stride y_stride max_deg seed
#include <stdio.h>
#include <stdlib.h>

#include "cuda_emu.h"
#include "spmm_exact.cuh"

static uint64_t rng_state;
static inline uint32_t rnd() {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(rng_state >> 33);
}
static inline float rndf() { return ((float)(rnd() & 0xffffff) / 16777216.0f - 0.5f) * 0.4f; }

int main(int argc, char **argv) {
    if (argc < 8) return 2;
    const int64_t n_rows = atoll(argv[1]), n_cols = atoll(argv[2]);
    const int dim = atoi(argv[3]);
    const int64_t xs = atoll(argv[4]), ys = atoll(argv[5]);
    const int max_deg = atoi(argv[6]);
    rng_state = (uint64_t)atoll(argv[7]) * 2654435761u + 7u;
    std::vector<int32_t> rowptr(1, 0), col;
    std::vector<float> val;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int deg = (r % 7 == 3) ? 0 : (int)(rnd() % (max_deg + 1));       // isolated rows included
        int32_t c = -1;
        for (int e = 0; e < deg; ++e) {
            c += 1 + (int32_t)(rnd() % 3);                                      // ascending, distinct
            if (c >= n_cols) break;
            col.push_back(c);
            val.push_back(rndf());
        }
        rowptr.push_back((int32_t)col.size());
    }
    if (col.empty()) { col.push_back(0); val.push_back(0.f); }
    std::vector<float> x((size_t)((n_cols - 1) * xs + dim)), y((size_t)((n_rows - 1) * ys + dim), -7.f);
    for (auto &v : x) v = rndf();
    const int dim_pad = (dim + 31) / 32 * 32;
    const int64_t blocks = (n_rows * dim_pad + ssl_exact::NT - 1) / ssl_exact::NT;
    const int32_t *rp = rowptr.data(), *cp = col.data();
    const float *vp = val.data(), *xp = x.data();
    float *yp = y.data();
    emu_launch(dim3((unsigned)blocks), dim3(ssl_exact::NT), [&]() { ssl_exact::spmm_exact_kernel(rp, cp, vp, n_rows, xp, xs, dim, dim_pad, yp, ys); });
    int64_t bad = 0;
    for (int64_t r = 0; r < n_rows; ++r)
        for (int j = 0; j < dim; ++j) {
            float acc = 0.f;
            for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) acc = fmaf(val[e], x[(int64_t)col[e] * xs + j], acc);
            if (y[r * ys + j] != acc) ++bad;
        }
    for (int64_t r = 0; r + 1 < n_rows; ++r)
        for (int64_t j = dim; j < ys; ++j)
            if (y[r * ys + j] != -7.f) ++bad;                                   // the padding between rows is untouched
    printf("rows=%lld dim=%d nnz=%zu bad=%lld\n", (long long)n_rows, dim, col.size(), (long long)bad);
    return bad == 0 ? 0 : 1;
}
