// Y = A X with the accumulation order of the reference's CPU operator (opt-in evaluation mode, test.exact_order).
//
// t.spmm(adj, embeds) (lightgcn.py:29) on the reference's column-sorted COO adjacency evaluates every output element as ONE
// sequential fp32 FMA chain over the row's stored entries in ascending column order (pinned on the host:
// tests/test_host_emulation.py::test_reference_spmm_is_a_sequential_fma_chain_in_column_order).  prop_kernel keeps that order
// inside a row segment but splits rows of more than 128 entries into segments whose partial sums are added afterwards, and
// fuses the layer sum in another association order: equal to rounding, not to the bit.  This kernel trades the load balance
// for the order: one thread per output element (row r, column j), the whole CSR row walked sequentially, acc = fma(w, x, acc).
// A warp covers 32 consecutive columns of one row (coalesced gathers; the (col, val) loads are warp-uniform broadcasts); a hub
// row is as slow as its length -- acceptable for an evaluation pass, not for training.  With the layer sum formed in the
// reference's order, ((E0 + X1) + X2) + ..., and predict_tile_kernel's scores, full_predict then reproduces the reference's
// CPU full_predict bit for bit on the same parameters.
//
// Kernel source only (no runtime API): tests/emu compiles this file for the host.
#pragma once
#include <stdint.h>

namespace ssl_exact {

constexpr int NT = 256;

// rows [0, n_rows) of a CSR (rowptr [n_rows + 1], ascending columns inside a row); x [*, dim] with row stride x_stride;
// y [n_rows, dim] with row stride y_stride.  dim_pad = dim rounded up to a multiple of 32 (a warp never straddles two rows).
static __global__ void __launch_bounds__(NT)
spmm_exact_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const float *__restrict__ vals, int64_t n_rows,
                  const float *__restrict__ x, int64_t x_stride, int dim, int dim_pad, float *__restrict__ y, int64_t y_stride) {
    const int64_t t = (int64_t)blockIdx.x * NT + (int64_t)threadIdx.x;
    const int64_t r = t / dim_pad;
    const int j = (int)(t - r * dim_pad);
    if (r >= n_rows || j >= dim) return;
    float acc = 0.f;
    const int e1 = rowptr[r + 1];
    for (int e = rowptr[r]; e < e1; ++e) acc = fmaf(vals[e], x[(int64_t)colidx[e] * x_stride + j], acc);
    y[r * y_stride + j] = acc;
}

}  // namespace ssl_exact
