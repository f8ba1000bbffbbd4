// ssl_spmm_exact: Y = A X in the accumulation order of the reference's CPU t.spmm (spmm_exact.cuh) -- the opt-in evaluation mode
// test.exact_order.  Not on the training path.
#include "common.cuh"
#include "spmm_exact.cuh"

extern "C" int ssl_spmm_exact(const int32_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows, const float *x, int64_t x_stride,
                              int32_t dim, float *y, int64_t y_stride, void *stream) {
    SSL_CHECK_ARG(rowptr && x && y, "ssl_spmm_exact: null argument");
    SSL_CHECK_ARG(dim >= 1 && dim <= SSL_MAX_DIM && x_stride >= dim && y_stride >= dim, "ssl_spmm_exact: dim %d / strides out of range", dim);
    if (n_rows == 0) return SSL_OK;
    SSL_CHECK_ARG(colidx && vals, "ssl_spmm_exact: null CSR arrays");
    const int dim_pad = (dim + 31) / 32 * 32;
    const int64_t threads = n_rows * dim_pad;
    const int64_t blocks = (threads + ssl_exact::NT - 1) / ssl_exact::NT;
    SSL_CHECK_ARG(blocks <= 0x7fffffffll, "ssl_spmm_exact: too many rows for one launch");
    ssl_exact::spmm_exact_kernel<<<(unsigned)blocks, ssl_exact::NT, 0, (cudaStream_t)stream>>>(rowptr, colidx, vals, n_rows, x, x_stride, dim, dim_pad, y,
                                                                                             y_stride);
    SSL_LAUNCH_CHECK("spmm_exact_kernel");
    return SSL_OK;
}
