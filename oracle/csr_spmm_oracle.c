/*
 * oracle/csr_spmm_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's one-hop CSR x dense kernel
 *   FloatCSRMulDenseOMP   (reference: sgl/operators/csrc/matmul.c:23-40, decl matmul.h:5)
 * used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs as the checker.  Nothing under sgl_b200/ may link or call this file.
 *
 * Semantics restated:
 *   for every row i:  y[i, :] = y_in[i, :] + sum_{j in CSR order} a[j] * x[col[j], :]
 *   evaluated per output element as ONE sequential chain in CSR column order, starting from the
 *   caller supplied value of y (the reference requires a zeroed answer buffer and accumulates into it).
 *   Two roundings flavours exist in the wild and both are provided:
 *     *_fma  : one fused multiply-add per term  == the shipped libmatmul.so (built -mfma, contraction on)
 *     *_muladd: separate multiply then add       == scipy csr(float32).dot(X)  (csr_matvecs)
 *   plus a float64-accumulating variant that mirrors the non-Linux branch adj.dot(X)
 *   (reference: sgl/operators/base_op.py:34) for one hop.
 *
 * Differences from the reference that are deliberate (they are the defects listed in SURVEY.md 5):
 *   - row offsets are 64-bit (the reference multiplies two ints, matmul.c:33, and overflows at N*d >= 2^31);
 *   - indptr may be int64 (second entry point) so nnz >= 2^31 graphs can be checked;
 *   - accumulation is carried in a local row buffer instead of load/modify/store of y per term: the value
 *     chain per element is identical, so the bits are identical.
 *
 * Parity pinning: oracle/Makefile builds oracle/_ref/libmatmul_ref.so straight from the reference's own
 * matmul.c (when /root/reference is present) and tests/test_oracle.py asserts bit equality of *_fma with it
 * and with the committed golden vectors under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* one output row, fused multiply-add chain */
static inline void row_chain_fma(float *restrict yrow, const float *restrict a, const int32_t *restrict col,
                                 int64_t lo, int64_t hi, const float *restrict x, int64_t d)
{
    for (int64_t j = lo; j < hi; ++j) {
        const float w = a[j];
        const float *restrict xr = x + (int64_t)col[j] * d;
        for (int64_t k = 0; k < d; ++k) yrow[k] = __builtin_fmaf(w, xr[k], yrow[k]);
    }
}

/* one output row, rounded product then rounded sum (no contraction: the Makefile passes -ffp-contract=off) */
static inline void row_chain_muladd(float *restrict yrow, const float *restrict a, const int32_t *restrict col,
                                    int64_t lo, int64_t hi, const float *restrict x, int64_t d)
{
    for (int64_t j = lo; j < hi; ++j) {
        const float w = a[j];
        const float *restrict xr = x + (int64_t)col[j] * d;
        for (int64_t k = 0; k < d; ++k) {
            const float p = w * xr[k]; /* file is compiled with -ffp-contract=off: never fused */
            yrow[k] = yrow[k] + p;
        }
    }
}

#define DEFINE_SPMM(NAME, IPTR_T, ROWFN)                                                                        \
    ORACLE_API void NAME(float *y, const float *a, const int32_t *col, const IPTR_T *indptr, const float *x,    \
                         int64_t n_rows, int64_t d)                                                             \
    {                                                                                                           \
        _Pragma("omp parallel for schedule(dynamic, 256)") for (int64_t i = 0; i < n_rows; ++i)                 \
        {                                                                                                       \
            ROWFN(y + i * d, a, col, (int64_t)indptr[i], (int64_t)indptr[i + 1], x, d);                         \
        }                                                                                                       \
    }

DEFINE_SPMM(oracle_spmm_f32_fma_i32, int32_t, row_chain_fma)
DEFINE_SPMM(oracle_spmm_f32_fma_i64, int64_t, row_chain_fma)
DEFINE_SPMM(oracle_spmm_f32_muladd_i32, int32_t, row_chain_muladd)
DEFINE_SPMM(oracle_spmm_f32_muladd_i64, int64_t, row_chain_muladd)

/* float64 weights and accumulation, float64 input/output: one hop of the scipy fp64 path
 * (reference: sgl/operators/base_op.py:34 with the fp64 self._adj built at graph_op/laplacian_graph_op.py:18-19).
 * scipy's csr_matvecs does y += a*x with separate multiply and add in double. */
ORACLE_API void oracle_spmm_f64_i64(double *y, const double *a, const int32_t *col, const int64_t *indptr,
                                    const double *x, int64_t n_rows, int64_t d)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_rows; ++i) {
        double *restrict yrow = y + i * d;
        for (int64_t j = indptr[i]; j < indptr[i + 1]; ++j) {
            const double w = a[j];
            const double *restrict xr = x + (int64_t)col[j] * d;
            for (int64_t k = 0; k < d; ++k) {
                const double p = w * xr[k];
                yrow[k] = yrow[k] + p;
            }
        }
    }
}

ORACLE_API int oracle_num_threads(void)
{
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
