// legacy.cu -- the reference's two C entry points, same symbols and signatures, executed on the B200.
//
//   void FloatCSRMulDenseOMP(...)  reference sgl/operators/csrc/matmul.c:23-40 (decl matmul.h:5); loaded through ctypes
//                                  by sgl/operators/utils.py:14-38.  `answer` is accumulated into.
//   int  FloatCSRMulDense(...)     reference sgl/operators/csrc/cudamatmul.c:28-146 (cuSPARSE wrapper, dormant);
//                                  overwrites `answer` (beta = 0, cudamatmul.c:48) and returns EXIT_SUCCESS/FAILURE.
// Swapping libmatmul.so / libcudamatmul.so for libsglb200.so lets the UNMODIFIED reference wrapper run on the GPU.
// Both use the EXACT schedule, so the float32 results equal the shipped CPU library bit for bit.  Per call they pay
// an upload of the CSR and of X and a download of Y -- the handle API (sglb200_graph_create + sglb200_propagate*) is
// the fast path; these exist for drop-in compatibility only.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace sglb200 {
int spmm_launch(sglb200_graph *g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode, int accumulate,
                cudaStream_t stream);

static int legacy_hop(float *answer, const float *data, const int *indices, const int *indptr, const float *mat,
                      int n, int d, int accumulate)
{
    if (n < 0 || d < 0) {
        set_error("legacy hop: negative size");
        return SGLB200_ERR_INVALID;
    }
    if (n == 0 || d == 0) return SGLB200_OK;
    const int64_t nnz = indptr[n];
    sglb200_graph_t g = nullptr;
    int st = sglb200_graph_create(&g, n, n, nnz, indptr, 0, indices, data, SGLB200_HOST, 0, 0, nullptr);
    if (st != SGLB200_OK) return st;
    const size_t bytes = (size_t)n * (size_t)d * sizeof(float);
    float *dx = nullptr, *dy = nullptr;
    cudaError_t e = cudaMalloc(&dx, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dy, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(dx, mat, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && accumulate) e = cudaMemcpy(dy, answer, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = spmm_launch(g, dx, d, dy, d, d, SGLB200_MODE_EXACT, accumulate, nullptr);
        if (st == SGLB200_OK) e = cudaMemcpy(answer, dy, bytes, cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) {
        set_error("legacy hop: %s", cudaGetErrorString(e));
        st = SGLB200_ERR_CUDA;
    }
    cudaFree(dx);
    cudaFree(dy);
    sglb200_graph_destroy(g);
    return st;
}
}  // namespace sglb200

extern "C" {

void FloatCSRMulDenseOMP(float answer[], float data[], int indices[], int indptr[], float mat[], int mat_row,
                         int mat_col)
{
    const int st = sglb200::legacy_hop(answer, data, indices, indptr, mat, mat_row, mat_col, 1);
    if (st != SGLB200_OK) {
        // the reference signature has no error channel.  Never return a silently wrong buffer: the message goes to stderr
        // and stays readable through sglb200_last_error(), and `answer` is filled with NaN so that every consumer sees the
        // failure; SGLB200_LEGACY_ABORT=1 restores the hard stop (abort()) for batch jobs that prefer to die at the hop.
        fprintf(stderr, "libsglb200: FloatCSRMulDenseOMP failed (status %d): %s\n", st, sglb200_last_error());
        const char *hard = getenv("SGLB200_LEGACY_ABORT");
        if (hard && hard[0] == '1') abort();
        if (answer && mat_row > 0 && mat_col > 0) {
            const size_t count = (size_t)mat_row * (size_t)mat_col;
            const float nan_value = __builtin_nanf("");
            for (size_t i = 0; i < count; ++i) answer[i] = nan_value;
        }
    }
}

int FloatCSRMulDense(float answer[], int data_nnz, float data[], int indices[], int indptr[], float mat[], int mat_row,
                     int mat_col)
{
    if (mat_row > 0 && indptr[mat_row] != data_nnz) {
        fprintf(stderr, "libsglb200: FloatCSRMulDense: data_nnz=%d does not match indptr[mat_row]=%d\n", data_nnz,
                indptr[mat_row]);
        return EXIT_FAILURE;
    }
    const int st = sglb200::legacy_hop(answer, data, indices, indptr, mat, mat_row, mat_col, 0);
    if (st != SGLB200_OK) {
        fprintf(stderr, "libsglb200: FloatCSRMulDense failed (status %d): %s\n", st, sglb200_last_error());
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

}  // extern "C"
