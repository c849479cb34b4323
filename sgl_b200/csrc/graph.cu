// graph.cu -- CSR operator handle: residency in HBM, merge-path tile schedules, degree normalisation.
//
// Replaces the per-hop host work of the reference wrapper (sgl/operators/utils.py:10-40: dlopen, three N*d
// temporaries and a float64->float32 cast of the values EVERY hop): the CSR is uploaded and cast once, the
// nnz-balanced warp schedule is computed once, and K hops reuse both.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "trace.cuh"

namespace sglb200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void clear_error() { g_err[0] = '\0'; }

int check_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s): libsglb200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return SGLB200_ERR_NO_DEVICE;
    }
    return SGLB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// schedule construction
// ---------------------------------------------------------------------------------------------------------------
// Boundary t sits at merge-path diagonal k = min(t*tile_items, n_rows+nnz) of the two lists
//   A = row-end markers indptr[1..n_rows]   and   B = non-zero positions 0..nnz-1,
// where marker r precedes position j iff indptr[r+1] <= j.  i = #markers among the first k items is found by
// bisection of the monotone predicate indptr[r+1] + r <= k-1.  A boundary that falls inside a row with at most
// split_threshold non-zeros (or any row when split_threshold < 0: EXACT schedule) is moved back to the row start so
// the row is owned by one warp and its fp32 chain stays sequential.
__global__ void build_tiles_kernel(const int64_t *__restrict__ indptr, int64_t n_rows, int64_t nnz, int tile_items,
                                   int64_t n_tiles, int64_t split_threshold, int32_t *__restrict__ tile_row,
                                   int64_t *__restrict__ tile_nnz)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    const int64_t total = n_rows + nnz;
    int64_t k = t * (int64_t)tile_items;
    if (k > total) k = total;
    int64_t lo = k - nnz > 0 ? k - nnz : 0;
    int64_t hi = k < n_rows ? k : n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (indptr[mid + 1] + mid <= k - 1) lo = mid + 1;
        else hi = mid;
    }
    const int64_t i = lo;
    int64_t j = k - i;
    if (i < n_rows) {
        const int64_t start = indptr[i];
        const int64_t deg = indptr[i + 1] - start;
        if (j > start && (split_threshold < 0 || deg <= split_threshold)) j = start;
    }
    tile_row[t] = (int32_t)i;
    tile_nnz[t] = j;
}

// row cut by the END of tile t (its partial sum must be carried), or -1
__device__ __forceinline__ int32_t carried_row(const int64_t *indptr, const int32_t *tile_row, const int64_t *tile_nnz,
                                               int64_t n_rows, int64_t t)
{
    const int64_t i_end = tile_row[t + 1];
    if (i_end >= n_rows) return -1;
    const int64_t j0 = tile_nnz[t], j_end = tile_nnz[t + 1];
    const int64_t row_start = indptr[i_end];
    const int64_t from = row_start > j0 ? row_start : j0;
    return j_end > from ? (int32_t)i_end : -1;
}

// pass 0: counts[0] += runs, counts[1] += carrier tiles.  pass 1: appends (row, carriers, first carrier tile) of every run.
__global__ void carry_runs_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ tile_row,
                                  const int64_t *__restrict__ tile_nnz, int64_t n_rows, int64_t n_tiles, int pass,
                                  unsigned long long *counts, int32_t *carry_slot, int32_t *run_row, int64_t *run_base,
                                  int32_t *run_len, int64_t *run_head)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int32_t row = carried_row(indptr, tile_row, tile_nnz, n_rows, t);
    if (row < 0) return;
    if (t > 0 && carried_row(indptr, tile_row, tile_nnz, n_rows, t - 1) == row) return;  // not the head of its run
    int64_t len = 1;
    while (t + len < n_tiles && carried_row(indptr, tile_row, tile_nnz, n_rows, t + len) == row) ++len;
    if (pass == 0) {
        atomicAdd(&counts[0], 1ULL);
        atomicAdd(&counts[1], (unsigned long long)len);
    } else {
        const unsigned long long r = atomicAdd(&counts[2], 1ULL);
        run_row[r] = row;
        run_len[r] = (int32_t)len;
        run_head[r] = t;  // slots and the per-tile maps are laid out on the host once the runs are sorted
    }
}

void free_schedule(Schedule *s)
{
    cudaFree(s->tile_row);
    cudaFree(s->tile_nnz);
    cudaFree(s->carry_slot);
    cudaFree(s->tail_run);
    cudaFree(s->head_run);
    cudaFree(s->run_count);
    cudaFree(s->run_row);
    cudaFree(s->run_base);
    cudaFree(s->run_len);
    *s = Schedule();
}

int build_schedule(sglb200_graph *g, Schedule *s, int64_t split_threshold, cudaStream_t stream)
{
    free_schedule(s);
    const int64_t total = g->n_rows + g->nnz;
    const int64_t n_tiles = total > 0 ? (total + g->tile_items - 1) / g->tile_items : 0;
    s->n_tiles = n_tiles;
    SGL_CUDA_CHECK(cudaMalloc(&s->tile_row, sizeof(int32_t) * (n_tiles + 1)));
    SGL_CUDA_CHECK(cudaMalloc(&s->tile_nnz, sizeof(int64_t) * (n_tiles + 1)));
    SGL_CUDA_CHECK(cudaMalloc(&s->carry_slot, sizeof(int32_t) * (n_tiles > 0 ? n_tiles : 1)));
    g->bytes_resident += (size_t)(n_tiles + 1) * 12 + (size_t)n_tiles * 4;
    const int threads = 256;
    build_tiles_kernel<<<(unsigned)((n_tiles + 1 + threads - 1) / threads), threads, 0, stream>>>(
        g->indptr, g->n_rows, g->nnz, g->tile_items, n_tiles, split_threshold, s->tile_row, s->tile_nnz);
    SGL_CUDA_CHECK(cudaGetLastError());
    SGL_CUDA_CHECK(cudaMemsetAsync(s->carry_slot, 0xFF, sizeof(int32_t) * (n_tiles > 0 ? n_tiles : 1), stream));
    if (n_tiles > 0 && split_threshold >= 0) {
        unsigned long long *counts = nullptr;
        SGL_CUDA_CHECK(cudaMalloc(&counts, 4 * sizeof(unsigned long long)));
        SGL_CUDA_CHECK(cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), stream));
        const unsigned blocks = (unsigned)((n_tiles + threads - 1) / threads);
        carry_runs_kernel<<<blocks, threads, 0, stream>>>(g->indptr, s->tile_row, s->tile_nnz, g->n_rows, n_tiles, 0,
                                                          counts, nullptr, nullptr, nullptr, nullptr, nullptr);
        SGL_CUDA_CHECK(cudaGetLastError());
        unsigned long long h[4];
        SGL_CUDA_CHECK(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, stream));
        SGL_CUDA_CHECK(cudaStreamSynchronize(stream));
        s->n_runs = (int64_t)h[0];
        s->n_slots = (int64_t)h[1];
        if (s->n_runs > 0) {
            SGL_CUDA_CHECK(cudaMalloc(&s->run_row, sizeof(int32_t) * s->n_runs));
            SGL_CUDA_CHECK(cudaMalloc(&s->run_base, sizeof(int64_t) * s->n_runs));
            SGL_CUDA_CHECK(cudaMalloc(&s->run_len, sizeof(int32_t) * s->n_runs));
            g->bytes_resident += (size_t)s->n_runs * 16;
            int64_t *run_head = nullptr;
            SGL_CUDA_CHECK(cudaMalloc(&run_head, sizeof(int64_t) * s->n_runs));
            carry_runs_kernel<<<blocks, threads, 0, stream>>>(g->indptr, s->tile_row, s->tile_nnz, g->n_rows, n_tiles,
                                                              1, counts, s->carry_slot, s->run_row, s->run_base,
                                                              s->run_len, run_head);
            cudaError_t e = cudaGetLastError();
            // order the runs by tile (== by row) on the host so that a tile range owns a contiguous run range
            const size_t nr = (size_t)s->n_runs;
            std::vector<int32_t> h_row(nr), h_len(nr);
            std::vector<int64_t> h_head(nr);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_row.data(), s->run_row, nr * 4, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_len.data(), s->run_len, nr * 4, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_head.data(), run_head, nr * 8, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            cudaFree(run_head);
            SGL_CUDA_CHECK(e);
            std::vector<size_t> order(nr);
            for (size_t i = 0; i < nr; ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return h_head[a] < h_head[b]; });
            std::vector<int32_t> s_row(nr), s_len(nr);
            std::vector<int64_t> s_base(nr);
            std::vector<int32_t> t_slot((size_t)n_tiles, -1), t_tail((size_t)n_tiles, -1), t_head((size_t)n_tiles, -1);
            s->run_last_tile.resize(nr);
            int64_t next_slot = 0;
            for (size_t i = 0; i < nr; ++i) {
                const size_t o = order[i];
                const int64_t head = h_head[o];
                const int32_t len = h_len[o];
                s_row[i] = h_row[o];
                s_len[i] = len;
                s_base[i] = next_slot;
                for (int32_t u = 0; u < len; ++u) {
                    t_slot[(size_t)(head + u)] = (int32_t)(next_slot + u);
                    t_tail[(size_t)(head + u)] = (int32_t)i;
                }
                t_head[(size_t)(head + len)] = (int32_t)i;   // the tile after the last carrier finishes the row
                s->run_last_tile[i] = head + len;
                next_slot += len + 1;                        // slot `len` parks the finisher's piece (fused row flush)
            }
            s->n_slots = next_slot;
            SGL_CUDA_CHECK(cudaMalloc(&s->tail_run, sizeof(int32_t) * n_tiles));
            SGL_CUDA_CHECK(cudaMalloc(&s->head_run, sizeof(int32_t) * n_tiles));
            SGL_CUDA_CHECK(cudaMalloc(&s->run_count, sizeof(uint32_t) * nr));
            g->bytes_resident += (size_t)n_tiles * 8 + nr * 4;
            SGL_CUDA_CHECK(cudaMemsetAsync(s->run_count, 0, sizeof(uint32_t) * nr, stream));
            SGL_CUDA_CHECK(cudaMemcpyAsync(s->run_row, s_row.data(), nr * 4, cudaMemcpyHostToDevice, stream));
            SGL_CUDA_CHECK(cudaMemcpyAsync(s->run_len, s_len.data(), nr * 4, cudaMemcpyHostToDevice, stream));
            SGL_CUDA_CHECK(cudaMemcpyAsync(s->run_base, s_base.data(), nr * 8, cudaMemcpyHostToDevice, stream));
            SGL_CUDA_CHECK(cudaMemcpyAsync(s->carry_slot, t_slot.data(), (size_t)n_tiles * 4, cudaMemcpyHostToDevice, stream));
            SGL_CUDA_CHECK(cudaMemcpyAsync(s->tail_run, t_tail.data(), (size_t)n_tiles * 4, cudaMemcpyHostToDevice, stream));
            SGL_CUDA_CHECK(cudaMemcpyAsync(s->head_run, t_head.data(), (size_t)n_tiles * 4, cudaMemcpyHostToDevice, stream));
            SGL_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        SGL_CUDA_CHECK(cudaStreamSynchronize(stream));
        cudaFree(counts);
    }
    s->built = true;
    return SGLB200_OK;
}

int ensure_carry_ws(sglb200_graph *g, size_t floats, cudaStream_t stream)
{
    if (floats <= g->carry_ws_floats) return SGLB200_OK;
    if (g->carry_ws) {
        // stream-ordered: hops already enqueued on this stream keep the old buffer until they are done (hops of one handle
        // are serialised on one stream, spmm_launch_ex); no device-wide synchronisation in the middle of a run
        SGL_CUDA_CHECK(cudaFreeAsync(g->carry_ws, stream));
        g->bytes_resident -= g->carry_ws_floats * sizeof(float);
        g->carry_ws = nullptr;
        g->carry_ws_floats = 0;
    }
    SGL_CUDA_CHECK(cudaMallocAsync(&g->carry_ws, floats * sizeof(float), stream));
    g->carry_ws_floats = floats;
    g->bytes_resident += floats * sizeof(float);
    return SGLB200_OK;
}

// idx_tag[j] = indices[j], bit 31 set on the last non-zero of every row; counts the rows without non-zeros
__global__ void tag_row_ends_kernel(const int64_t *__restrict__ indptr, int64_t n_rows, int32_t *__restrict__ idx_tag,
                                    unsigned long long *__restrict__ n_empty)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int64_t b = indptr[r], e = indptr[r + 1];
    if (e > b) idx_tag[e - 1] |= (int32_t)0x80000000;
    else atomicAdd(n_empty, 1ULL);
}

int build_stream_tags(sglb200_graph *g, cudaStream_t stream)
{
    if (g->idx_tag || g->nnz == 0 || g->n_rows == 0) return SGLB200_OK;
    unsigned long long *cnt = nullptr;
    SGL_CUDA_CHECK(cudaMalloc(&cnt, sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc(&g->idx_tag, sizeof(int32_t) * (g->nnz + kStreamPad));
    if (e == cudaSuccess) e = cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->idx_tag + g->nnz, 0, sizeof(int32_t) * kStreamPad, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->idx_tag, g->indices, sizeof(int32_t) * g->nnz, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess) {
        tag_row_ends_kernel<<<(unsigned)((g->n_rows + 255) / 256), 256, 0, stream>>>(g->indptr, g->n_rows, g->idx_tag, cnt);
        e = cudaGetLastError();
    }
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(cnt);
    if (e != cudaSuccess) {
        cudaFree(g->idx_tag);
    cudaFree(g->idx_cold);
        g->idx_tag = nullptr;
        SGL_CUDA_CHECK(e);
    }
    g->empty_rows = (int64_t)h;
    g->bytes_resident += sizeof(int32_t) * (size_t)(g->nnz + kStreamPad);
    return SGLB200_OK;
}

// ---- cold-column tags (experiment: SGLB200_COLD_HINT) ---------------------------------------------------------------
// idx_cold[j] = indices[j] | (1<<30 if column indices[j] is referenced fewer times than the hub threshold).  The hop
// kernel loads rows of cold columns with an L2 evict_first policy so that they do not push the few heavily re-read
// rows out of L2.  The threshold is the smallest reference count such that at most `hub_rows` columns reach it.
__global__ void count_cols_kernel(const int32_t *__restrict__ indices, int64_t nnz, int32_t *__restrict__ counts)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&counts[indices[j]], 1);
}

constexpr int kCountBuckets = 4096;

__global__ void count_hist_kernel(const int32_t *__restrict__ counts, int64_t n_cols, unsigned long long *__restrict__ hist)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += (int64_t)gridDim.x * blockDim.x) {
        const int k = counts[c] < kCountBuckets - 1 ? counts[c] : kCountBuckets - 1;
        atomicAdd(&hist[k], 1ULL);
    }
}

__global__ void tag_cold_kernel(const int32_t *__restrict__ indices, int64_t nnz, const int32_t *__restrict__ counts,
                                int threshold, int32_t *__restrict__ out)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = indices[j];
        out[j] = counts[c] < threshold ? (c | 0x40000000) : c;
    }
}

int build_cold_tags(sglb200_graph *g, int64_t hub_rows, cudaStream_t stream)
{
    if (g->nnz == 0 || g->n_cols >= (1LL << 30)) return SGLB200_OK;
    if (g->idx_cold && g->cold_hub_rows == hub_rows) return SGLB200_OK;
    int32_t *counts = nullptr;
    unsigned long long *hist = nullptr;
    SGL_CUDA_CHECK(cudaMalloc(&counts, sizeof(int32_t) * g->n_cols));
    cudaError_t e = cudaMalloc(&hist, sizeof(unsigned long long) * kCountBuckets);
    if (e == cudaSuccess && !g->idx_cold) {
        e = cudaMalloc(&g->idx_cold, sizeof(int32_t) * (g->nnz + kStreamPad));
        if (e == cudaSuccess) {
            g->bytes_resident += sizeof(int32_t) * (size_t)(g->nnz + kStreamPad);
            e = cudaMemsetAsync(g->idx_cold + g->nnz, 0, sizeof(int32_t) * kStreamPad, stream);
        }
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * g->n_cols, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * kCountBuckets, stream);
    std::vector<unsigned long long> h(kCountBuckets);
    if (e == cudaSuccess) {
        count_cols_kernel<<<148 * 16, 256, 0, stream>>>(g->indices, g->nnz, counts);
        count_hist_kernel<<<148 * 8, 256, 0, stream>>>(counts, g->n_cols, hist);
        e = cudaMemcpyAsync(h.data(), hist, sizeof(unsigned long long) * kCountBuckets, cudaMemcpyDeviceToHost, stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    int threshold = kCountBuckets;       // nothing is a hub
    if (e == cudaSuccess) {
        unsigned long long above = 0;
        for (int k = kCountBuckets - 1; k >= 1; --k) {
            if (above + h[k] > (unsigned long long)hub_rows) break;
            above += h[k];
            threshold = k;
        }
        tag_cold_kernel<<<148 * 16, 256, 0, stream>>>(g->indices, g->nnz, counts, threshold, g->idx_cold);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    }
    cudaFree(counts);
    cudaFree(hist);
    SGL_CUDA_CHECK(e);
    g->cold_hub_rows = hub_rows;
    g->cold_threshold = threshold;
    return SGLB200_OK;
}

__global__ void interleave_pairs_kernel(const int32_t *__restrict__ idx_tag, const float *__restrict__ vals, int64_t count,
                                        int2 *__restrict__ pairs)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (int64_t)gridDim.x * blockDim.x)
        pairs[j] = make_int2(idx_tag[j], __float_as_int(vals[j]));
}

int build_stream_pairs(sglb200_graph *g, cudaStream_t stream)
{
    if (!g->idx_tag || g->nnz == 0) return SGLB200_OK;
    if (!g->pairs) {
        SGL_CUDA_CHECK(cudaMalloc(&g->pairs, sizeof(int2) * (g->nnz + kStreamPad)));
        g->bytes_resident += sizeof(int2) * (size_t)(g->nnz + kStreamPad);
        g->pairs_valid = false;
    }
    if (!g->pairs_valid) {
        const int64_t count = g->nnz + kStreamPad;   // the padding of both arrays is zero
        int64_t blocks = (count + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        interleave_pairs_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g->idx_tag, g->vals, count, g->pairs);
        SGL_CUDA_CHECK(cudaGetLastError());
        g->pairs_valid = true;
    }
    return SGLB200_OK;
}

__global__ void widen_indptr_kernel(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)in[i];
}

// vals[i,j] = fl32( (1-alpha) * ((w * dL[i]) * dR[j]) + alpha*[i==j] ), float64 products in the reference's order
__global__ void normalize_values_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                        const double *__restrict__ raw_w, const double *__restrict__ d_left,
                                        const double *__restrict__ d_right, double one_minus_alpha, double alpha,
                                        int apply_ppr, int64_t n_rows, float *__restrict__ vals)
{
    // one warp per row: coalesced over the row's entries, no search
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const double dl = d_left[row];
    for (int64_t j = indptr[row] + lane; j < indptr[row + 1]; j += 32) {
        const int32_t c = indices[j];
        double v = __dmul_rn(__dmul_rn(raw_w[j], dl), d_right[c]);
        if (apply_ppr) {
            v = __dmul_rn(one_minus_alpha, v);
            if ((int64_t)c == row) v = __dadd_rn(v, alpha);
        }
        vals[j] = (float)v;
    }
}

// float32 copies for the fused normalisation: raw weights, (1-alpha)*dL, dR, alpha/dR; flags[0] counts non-unit weights
__global__ void scaling_vectors_kernel(const double *__restrict__ d_left, const double *__restrict__ d_right, double one_minus_alpha,
                                       double alpha, int apply_ppr, int64_t n, float *__restrict__ row_scale,
                                       float *__restrict__ col_scale, float *__restrict__ self_coef)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dl = d_left[i], dr = d_right[i];
    row_scale[i] = (float)(apply_ppr ? one_minus_alpha * dl : dl);
    col_scale[i] = (float)dr;
    if (self_coef) self_coef[i] = dr != 0.0 ? (float)(alpha / dr) : 0.0f;
}
__global__ void raw_weights_kernel(const double *__restrict__ raw_w, int64_t nnz, float *__restrict__ out,
                                   unsigned long long *__restrict__ non_unit)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    const double w = raw_w[j];
    out[j] = (float)w;
    if (w != 1.0) atomicAdd(non_unit, 1ULL);
}

}  // namespace sglb200

using namespace sglb200;

extern "C" {

int sglb200_version(void) { return SGLB200_VERSION; }
const char *sglb200_last_error(void) { return g_err; }

int sglb200_device_count(void)
{
    const int st = check_device();
    if (st != SGLB200_OK) return -st;
    int n = 0;
    cudaGetDeviceCount(&n);
    return n;
}

int sglb200_set_device(int device)
{
    const int st = check_device();
    if (st != SGLB200_OK) return st;
    SGL_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    SGL_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libsglb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return SGLB200_ERR_NO_DEVICE;
    }
    return SGLB200_OK;
}

int sglb200_graph_create(sglb200_graph_t *out, int64_t n_rows, int64_t n_cols, int64_t nnz, const void *indptr,
                         int indptr_is64, const int32_t *indices, const float *vals, int loc, int tile_items,
                         int split_threshold, void *stream_)
{
    clear_error();
    TraceRange range("sglb200_graph_create");
    SGL_REQUIRE(out != nullptr, "graph_create: out is NULL");
    *out = nullptr;
    SGL_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "graph_create: negative size");
    SGL_REQUIRE(n_rows < (1LL << 31) - 64 && n_cols < (1LL << 31), "graph_create: row/column ids must fit int32");
    SGL_REQUIRE(indptr != nullptr, "graph_create: indptr is NULL");
    SGL_REQUIRE(nnz == 0 || indices != nullptr, "graph_create: indices is NULL");
    SGL_REQUIRE(indptr_is64 || nnz < (1LL << 31), "graph_create: nnz >= 2^31 needs int64 indptr");
    SGL_REQUIRE(loc == SGLB200_HOST || loc == SGLB200_DEVICE, "graph_create: bad location");
    {
        const int st = check_device();
        if (st != SGLB200_OK) return st;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    sglb200_graph *g = new (std::nothrow) sglb200_graph();
    if (!g) {
        set_error("graph_create: out of host memory");
        return SGLB200_ERR_ALLOC;
    }
    int status = SGLB200_OK;
    auto fail = [&](int st) {
        sglb200_graph_destroy(g);
        return st;
    };
#define G_CHECK(expr)                                                                                            \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess) {                                                                                \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);              \
            return fail(SGLB200_ERR_CUDA);                                                                       \
        }                                                                                                        \
    } while (0)
    G_CHECK(cudaGetDevice(&g->device));
    {
        cudaDeviceProp prop;
        G_CHECK(cudaGetDeviceProperties(&prop, g->device));
        if (prop.major != 10) {
            set_error("device %d is sm_%d%d; libsglb200 is built for sm_100a (B200) only", g->device, prop.major,
                      prop.minor);
            return fail(SGLB200_ERR_NO_DEVICE);
        }
        g->sm_count = prop.multiProcessorCount;
    }
    g->n_rows = n_rows;
    g->n_cols = n_cols;
    g->nnz = nnz;
    if (tile_items > 0) {
        g->tile_items = tile_items;
    } else {
        // measured on B200 (profiles/): 256 items per warp once the graph fills the chip; smaller graphs get
        // proportionally smaller tiles so that every SM still holds ~48 warps' worth of tiles
        const int64_t total = n_rows + nnz;
        const int64_t per_warp = total / ((int64_t)g->sm_count * 48);
        int64_t ti = ((per_warp + 31) / 32) * 32;
        if (ti < 32) ti = 32;
        if (ti > 256) ti = 256;
        // graphs that fill the chip many times over: longer tiles amortise the tile start-up (descriptor, first stream
        // batch, first row-end window).  Measured per hop (profiles/r02_tile_items.txt): products-shape 5.32 / 5.04 / 4.87 ms
        // at 256 / 512 / 1024 items, rmat22 3.50 / 3.36 ms at 256 / 512, products d=16 1.18 / 1.04 / 0.99 ms at 128 / 256 / 512
        const char *cap = getenv("SGLB200_TILE_ITEMS_MAX");
        const int64_t max_items = cap ? atoll(cap) : 1024;
        while (ti < max_items && per_warp >= 8 * ti) ti *= 2;
        g->tile_items = (int)ti;
    }
    g->split_threshold = split_threshold > 0 ? split_threshold : 64;  // rows above 64 non-zeros may be cut
    const cudaMemcpyKind kind = loc == SGLB200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    G_CHECK(cudaMalloc(&g->indptr, sizeof(int64_t) * (n_rows + 1)));
    G_CHECK(cudaMalloc(&g->indices, sizeof(int32_t) * (nnz > 0 ? nnz : 1)));
    G_CHECK(cudaMalloc(&g->vals, sizeof(float) * (nnz + kStreamPad)));
    G_CHECK(cudaMemsetAsync(g->vals + nnz, 0, sizeof(float) * kStreamPad, stream));
    g->bytes_resident = sizeof(int64_t) * (n_rows + 1) + 8 * (size_t)nnz;
    if (indptr_is64) {
        G_CHECK(cudaMemcpyAsync(g->indptr, indptr, sizeof(int64_t) * (n_rows + 1), kind, stream));
    } else {
        int32_t *tmp = nullptr;
        G_CHECK(cudaMalloc(&tmp, sizeof(int32_t) * (n_rows + 1)));
        cudaError_t e = cudaMemcpyAsync(tmp, indptr, sizeof(int32_t) * (n_rows + 1), kind, stream);
        if (e == cudaSuccess) {
            widen_indptr_kernel<<<(unsigned)((n_rows + 1 + 255) / 256), 256, 0, stream>>>(tmp, g->indptr, n_rows + 1);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(tmp);
        G_CHECK(e);
    }
    if (nnz > 0) {
        G_CHECK(cudaMemcpyAsync(g->indices, indices, sizeof(int32_t) * nnz, kind, stream));
        if (vals) G_CHECK(cudaMemcpyAsync(g->vals, vals, sizeof(float) * nnz, kind, stream));
        else G_CHECK(cudaMemsetAsync(g->vals, 0, sizeof(float) * nnz, stream));
    }
    // consistency of the row pointer ends (cheap, catches truncated inputs)
    {
        int64_t ends[2] = {0, 0};
        G_CHECK(cudaMemcpyAsync(&ends[0], g->indptr, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
        G_CHECK(cudaMemcpyAsync(&ends[1], g->indptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
        G_CHECK(cudaStreamSynchronize(stream));
        if (ends[0] != 0 || ends[1] != nnz) {
            set_error("graph_create: indptr[0]=%lld indptr[n_rows]=%lld do not match nnz=%lld", (long long)ends[0],
                      (long long)ends[1], (long long)nnz);
            return fail(SGLB200_ERR_INVALID);
        }
    }
    status = build_schedule(g, &g->fast, g->split_threshold, stream);
    if (status != SGLB200_OK) return fail(status);
    G_CHECK(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 3; ++k) {
        G_CHECK(cudaEventCreateWithFlags(&g->ev_compute[k], cudaEventDisableTiming));
        G_CHECK(cudaEventCreateWithFlags(&g->ev_copy[k], cudaEventDisableTiming));
    }
#undef G_CHECK
    *out = g;
    return SGLB200_OK;
}

int sglb200_graph_destroy(sglb200_graph_t g)
{
    if (!g) return SGLB200_OK;
    cudaDeviceSynchronize();
    free_schedule(&g->fast);
    free_schedule(&g->exact);
    cudaFree(g->indptr);
    cudaFree(g->indices);
    cudaFree(g->vals);
    cudaFree(g->idx_tag);
    cudaFree(g->pairs);
    cudaFree(g->raw_w);
    cudaFree(g->row_scale);
    cudaFree(g->col_scale);
    cudaFree(g->self_coef);
    cudaFree(g->ping[0]);
    cudaFree(g->ping[1]);
    cudaFree(g->aux);
    cudaFree(g->carry_ws);
    for (int k = 0; k < 3; ++k) {
        cudaFree(g->stage[k]);
        if (g->ev_compute[k]) cudaEventDestroy(g->ev_compute[k]);
        if (g->ev_copy[k]) cudaEventDestroy(g->ev_copy[k]);
    }
    if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
    (void)cudaGetLastError();
    delete g;
    return SGLB200_OK;
}

int sglb200_graph_set_values(sglb200_graph_t g, const float *vals, int loc, void *stream)
{
    clear_error();
    SGL_REQUIRE(g && vals, "graph_set_values: NULL argument");
    if (g->nnz == 0) return SGLB200_OK;
    g->pairs_valid = false;
    SGL_CUDA_CHECK(cudaMemcpyAsync(g->vals, vals, sizeof(float) * g->nnz,
                                   loc == SGLB200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                   (cudaStream_t)stream));
    if (loc == SGLB200_HOST) SGL_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return SGLB200_OK;
}

int sglb200_graph_info(sglb200_graph_t g, int64_t info[9])
{
    clear_error();
    SGL_REQUIRE(g && info, "graph_info: NULL argument");
    info[0] = g->n_rows;
    info[1] = g->n_cols;
    info[2] = g->nnz;
    info[3] = g->fast.n_tiles;
    info[4] = g->fast.n_runs;
    info[5] = g->exact.built ? g->exact.n_tiles : -1;
    info[6] = g->tile_items;
    info[7] = g->split_threshold;
    info[8] = (int64_t)g->bytes_resident;
    return SGLB200_OK;
}

int sglb200_graph_chunks(sglb200_graph_t g, int mode, int n_chunks, int64_t *tile_bounds, int64_t *row_bounds)
{
    clear_error();
    SGL_REQUIRE(g && tile_bounds && row_bounds && n_chunks >= 1, "graph_chunks: bad argument");
    Schedule *s = &g->fast;
    if (mode == SGLB200_MODE_EXACT) {
        if (!g->exact.built) {
            const int st = build_schedule(g, &g->exact, -1, nullptr);
            if (st != SGLB200_OK) return st;
        }
        s = &g->exact;
    }
    for (int c = 0; c <= n_chunks; ++c) {
        tile_bounds[c] = s->n_tiles * (int64_t)c / n_chunks;
        int32_t row = (int32_t)g->n_rows;
        if (c < n_chunks && s->n_tiles > 0)
            SGL_CUDA_CHECK(cudaMemcpy(&row, s->tile_row + tile_bounds[c], sizeof(int32_t), cudaMemcpyDeviceToHost));
        row_bounds[c] = c == 0 ? 0 : row;
    }
    row_bounds[n_chunks] = g->n_rows;
    return SGLB200_OK;
}

int sglb200_normalize_values(sglb200_graph_t g, const double *raw_w, const double *d_left, const double *d_right,
                             double alpha, int apply_ppr, int loc, void *stream_)
{
    clear_error();
    TraceRange range("sglb200_normalize_values");
    SGL_REQUIRE(g && raw_w && d_left && d_right, "normalize_values: NULL argument");
    SGL_REQUIRE(g->n_rows == g->n_cols, "normalize_values: operator must be square");
    g->pairs_valid = false;
    cudaStream_t stream = (cudaStream_t)stream_;
    const double *w = raw_w, *dl = d_left, *dr = d_right;
    double *tmp = nullptr;
    if (loc == SGLB200_HOST) {
        const size_t n = (size_t)g->n_rows, m = (size_t)g->nnz;
        SGL_CUDA_CHECK(cudaMalloc(&tmp, sizeof(double) * (m + 2 * n + 1)));
        cudaError_t e = cudaMemcpyAsync(tmp, raw_w, sizeof(double) * m, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + m, d_left, sizeof(double) * n, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + m + n, d_right, sizeof(double) * n, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) {
            cudaFree(tmp);
            SGL_CUDA_CHECK(e);
        }
        w = tmp;
        dl = tmp + m;
        dr = tmp + m + n;
    }
    if (g->n_rows > 0) {
        const int threads = 256;
        const int64_t warps = g->n_rows;
        normalize_values_kernel<<<(unsigned)((warps * 32 + threads - 1) / threads), threads, 0, stream>>>(
            g->indptr, g->indices, w, dl, dr, 1.0 - alpha, alpha, apply_ppr, g->n_rows, g->vals);
    }
    cudaError_t e = cudaGetLastError();
    // float32 scaling vectors + raw weights for the fused normalisation of the fused driver (FAST mode)
    if (e == cudaSuccess && g->n_rows > 0 && g->nnz > 0) {
        const size_t n = (size_t)g->n_rows, m = (size_t)g->nnz;
        unsigned long long *cnt = nullptr;
        cudaFree(g->raw_w); cudaFree(g->row_scale); cudaFree(g->col_scale); cudaFree(g->self_coef);
        g->raw_w = g->row_scale = g->col_scale = g->self_coef = nullptr;
        g->has_scaling = false;
        e = cudaMalloc(&cnt, sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), stream);
        if (e == cudaSuccess) e = cudaMalloc(&g->raw_w, sizeof(float) * (m + kStreamPad));
        if (e == cudaSuccess) e = cudaMemsetAsync(g->raw_w + m, 0, sizeof(float) * kStreamPad, stream);
        if (e == cudaSuccess) e = cudaMalloc(&g->row_scale, sizeof(float) * n);
        if (e == cudaSuccess) e = cudaMalloc(&g->col_scale, sizeof(float) * n);
        if (e == cudaSuccess && apply_ppr) e = cudaMalloc(&g->self_coef, sizeof(float) * n);
        if (e == cudaSuccess) {
            scaling_vectors_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dl, dr, 1.0 - alpha, alpha, apply_ppr, (int64_t)n,
                                                                                   g->row_scale, g->col_scale, g->self_coef);
            raw_weights_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(w, (int64_t)m, g->raw_w, cnt);
            e = cudaGetLastError();
        }
        unsigned long long non_unit = 1;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&non_unit, cnt, sizeof(non_unit), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(cnt);
        if (e == cudaSuccess) {
            g->has_scaling = true;
            g->unit_weights = non_unit == 0;
            if (g->unit_weights) {
                cudaFree(g->raw_w);
                g->raw_w = nullptr;
            }
        }
    }
    if (tmp) {
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(tmp);
    }
    SGL_CUDA_CHECK(e);
    return SGLB200_OK;
}

}  // extern "C"
