// build_adj.cu -- A~ = A + I and its transpose in CSR order, built on the device from a COO edge list with our own kernels
// (SURVEY.md section 8f-2).
//
// What the reference does with scipy on one core (sgl/operators/utils.py:76-88, graph_op/laplacian_graph_op.py:19):
//     A   = csr_matrix((w, (row, col)))      duplicates merged in the weights' dtype (float32)
//     A~  = A + sp.eye(n)                    float64; the sparse add drops exact zeros
//     deg = A~.sum(1)                        float64, summed in column order
//     A^  = (A~ diag(dL))^T diag(dR)         i.e. entry (j, i) of A~ lands at (i, j); .tocsr() sorts the columns
// Here: one LSD radix sort of 64-bit keys (row | col | identity bit) brings duplicates and the identity entry of every
// (row, col) together in input order; one pass folds each run (float32 sum of A's duplicates, + 1.0 in float64 when the
// identity entry is present) and drops zeros; degrees are summed per row sequentially (column order, so they equal
// scipy's bit for bit); a second radix sort by column transposes.  The value pass (dL, dR, PPR mix) is
// sglb200_normalize_values.  The sort is ours (per-block digit histograms, a 3-kernel exclusive scan, a stable
// scatter that ranks equal digits with __match_any_sync); no library sort, no host round trip except the three scalars
// (entry counts, error flag).
#include <stdint.h>

#include <new>

#include "common.cuh"
#include "trace.cuh"

struct sglb200_adj_builder {
    int64_t n = 0;
    int64_t nnz = 0;
    int32_t *row_of = nullptr;     // nnz: row of A~ (= column of A^) of entry p, entries in (row, col) order
    double *w2 = nullptr;          // nnz: merged float64 weight of entry p
    uint32_t *col_sorted = nullptr;   // nnz: columns of A~ (= rows of A^) in transposed order
    uint32_t *perm = nullptr;      // nnz: entry p at transposed position q
    double *deg = nullptr;         // n
};

namespace sglb200 {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;
constexpr int kSortTile = kSortThreads * kSortRounds;

// ---- exclusive scan (int64 out) ----------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ int64_t block_exclusive_scan(int64_t mine, int64_t *total)
{
    __shared__ int64_t warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int64_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const int64_t s = warp_sums[w];
        if (w < warp) before += s;
        all += s;
    }
    __syncthreads();
    *total = all;
    return before + incl - mine;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const T *__restrict__ in, int64_t count, int64_t *__restrict__ partial)
{
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < count) s += (int64_t)in[base + i];
    int64_t total;
    block_exclusive_scan<T>(s, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const T *__restrict__ in, int64_t count, const int64_t *__restrict__ partial_scanned,
                                                                  int64_t *__restrict__ out)
{
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    int64_t v[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < count ? (int64_t)in[base + i] : 0;
        s += v[i];
    }
    int64_t total;
    int64_t run = block_exclusive_scan<T>(s, &total) + (partial_scanned ? partial_scanned[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < count) out[base + i] = run;
        run += v[i];
    }
}

// out[i] = sum of in[0..i); out may not alias in.  Temporary partials come from `scratch` (>= scan_scratch_elems(count)).
int64_t scan_scratch_elems(int64_t count)
{
    int64_t total = 0;
    while (count > kScanChunk) {
        count = (count + kScanChunk - 1) / kScanChunk;
        total += 2 * count;
    }
    return total + 2;
}

template <typename T>
cudaError_t exclusive_scan(const T *in, int64_t count, int64_t *out, int64_t *scratch, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    const int64_t blocks = (count + kScanChunk - 1) / kScanChunk;
    if (blocks == 1) {
        scan_apply_kernel<T><<<1, kScanThreads, 0, stream>>>(in, count, nullptr, out);
        return cudaGetLastError();
    }
    int64_t *partial = scratch, *partial_scanned = scratch + blocks;
    scan_reduce_kernel<T><<<(unsigned)blocks, kScanThreads, 0, stream>>>(in, count, partial);
    cudaError_t e = exclusive_scan<int64_t>(partial, blocks, partial_scanned, scratch + 2 * blocks, stream);
    if (e != cudaSuccess) return e;
    scan_apply_kernel<T><<<(unsigned)blocks, kScanThreads, 0, stream>>>(in, count, partial_scanned, out);
    return cudaGetLastError();
}

// ---- LSD radix sort, 8-bit digits, stable -----------------------------------------------------------------------------
template <typename Key>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const Key *__restrict__ keys, int64_t count, int shift, uint32_t *__restrict__ block_hist,
                                                                 int64_t n_blocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) {
        const int64_t j = base + i;
        if (j < count) atomicAdd(&h[(unsigned)(keys[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    block_hist[(int64_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];   // digit-major: the scan yields global offsets
}

// Every block walks its tile in rounds of 256 consecutive items.  Within a round equal digits are ranked inside a warp by
// __match_any_sync (lower lanes first), across warps by a prefix over the per-warp digit counts, across rounds by the
// running base of the digit: items keep their input order inside every digit bucket.
template <typename Key, typename Val, bool HAS_VAL>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const Key *__restrict__ keys_in, const Val *__restrict__ vals_in, Key *__restrict__ keys_out,
                                                                    Val *__restrict__ vals_out, int64_t count, int shift,
                                                                    const int64_t *__restrict__ offsets, int64_t n_blocks)
{
    __shared__ uint32_t warp_cnt[kSortThreads / 32][256];
    __shared__ int64_t base[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = offsets[(int64_t)threadIdx.x * n_blocks + blockIdx.x];
    const int64_t tile0 = (int64_t)blockIdx.x * kSortTile;
    for (int round = 0; round < kSortRounds; ++round) {
        const int64_t first = tile0 + (int64_t)round * kSortThreads;
        if (first >= count) break;   // uniform
        for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
        __syncthreads();
        const int64_t j = first + threadIdx.x;
        const bool ok = j < count;
        Key k = 0;
        unsigned dg = 256u;   // lanes past the end group among themselves and never write
        if (ok) {
            k = keys_in[j];
            dg = (unsigned)(k >> shift) & 255u;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, dg);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (ok && rank == 0) warp_cnt[warp][dg] = (uint32_t)__popc(peers);
        __syncthreads();
        uint32_t digit_total = 0;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {   // thread t owns digit t
            const uint32_t c = warp_cnt[w][threadIdx.x];
            warp_cnt[w][threadIdx.x] = digit_total;
            digit_total += c;
        }
        __syncthreads();
        if (ok) {
            const int64_t pos = base[dg] + warp_cnt[warp][dg] + rank;
            keys_out[pos] = k;
            if (HAS_VAL) vals_out[pos] = vals_in[j];
        }
        __syncthreads();
        base[threadIdx.x] += digit_total;
    }
}

struct SortWorkspace {
    uint32_t *block_hist = nullptr;   // 256 * n_blocks
    int64_t *offsets = nullptr;       // 256 * n_blocks
    int64_t *scan_scratch = nullptr;
    int64_t n_blocks = 0;
};

cudaError_t sort_workspace_alloc(SortWorkspace *ws, int64_t count)
{
    ws->n_blocks = (count + kSortTile - 1) / kSortTile;
    if (ws->n_blocks < 1) ws->n_blocks = 1;
    cudaError_t e = cudaMalloc(&ws->block_hist, sizeof(uint32_t) * 256 * ws->n_blocks);
    if (e == cudaSuccess) e = cudaMalloc(&ws->offsets, sizeof(int64_t) * 256 * ws->n_blocks);
    if (e == cudaSuccess) e = cudaMalloc(&ws->scan_scratch, sizeof(int64_t) * scan_scratch_elems(256 * ws->n_blocks));
    return e;
}

void sort_workspace_free(SortWorkspace *ws)
{
    cudaFree(ws->block_hist);
    cudaFree(ws->offsets);
    cudaFree(ws->scan_scratch);
    *ws = SortWorkspace();
}

// Sorts (keys, vals) by the low `bits` bits of the keys.  Ping-pongs between the a/b buffers; returns in *in_a whether the
// result ended in the a buffers.
template <typename Key, typename Val, bool HAS_VAL>
cudaError_t radix_sort(Key *keys_a, Key *keys_b, Val *vals_a, Val *vals_b, int64_t count, int bits, SortWorkspace *ws, bool *in_a,
                       cudaStream_t stream)
{
    bool a = true;
    for (int shift = 0; shift < bits && count > 1; shift += 8) {
        Key *ki = a ? keys_a : keys_b, *ko = a ? keys_b : keys_a;
        Val *vi = a ? vals_a : vals_b, *vo = a ? vals_b : vals_a;
        sort_hist_kernel<Key><<<(unsigned)ws->n_blocks, kSortThreads, 0, stream>>>(ki, count, shift, ws->block_hist, ws->n_blocks);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = exclusive_scan<uint32_t>(ws->block_hist, 256 * ws->n_blocks, ws->offsets, ws->scan_scratch, stream);
        if (e != cudaSuccess) return e;
        sort_scatter_kernel<Key, Val, HAS_VAL><<<(unsigned)ws->n_blocks, kSortThreads, 0, stream>>>(ki, vi, ko, vo, count, shift, ws->offsets,
                                                                                                   ws->n_blocks);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        a = !a;
    }
    *in_a = a;
    return cudaSuccess;
}

// ---- the adjacency passes ------------------------------------------------------------------------------------------------
// key = row << (col_bits + 1) | col << 1 | identity
__global__ void make_keys_kernel(const int64_t *__restrict__ rows, const int64_t *__restrict__ cols, const float *__restrict__ weights, int64_t n_edges,
                                 int64_t n, int add_identity, int col_bits, uint64_t *__restrict__ keys, float *__restrict__ vals,
                                 int *__restrict__ bad)
{
    const int64_t total = n_edges + (add_identity ? n : 0);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
        if (j < n_edges) {
            const int64_t r = rows[j], c = cols[j];
            if (r < 0 || r >= n || c < 0 || c >= n) {
                *bad = 1;
                keys[j] = 0;
            } else {
                keys[j] = ((uint64_t)r << (col_bits + 1)) | ((uint64_t)c << 1);
            }
            if (vals) vals[j] = weights ? weights[j] : 1.0f;
        } else {
            const uint64_t i = (uint64_t)(j - n_edges);
            keys[j] = (i << (col_bits + 1)) | (i << 1) | 1ULL;
            if (vals) vals[j] = 1.0f;
        }
    }
}

// One thread per sorted item; the first item of a (row, col) run folds the run: float32 sum of A's entries in input order,
// widened to float64, + 1.0 when the identity entry is there.  keep = run head with a non-zero result.
__global__ void fold_runs_kernel(const uint64_t *__restrict__ keys, const float *__restrict__ vals, int64_t count, uint8_t *__restrict__ keep,
                                 double *__restrict__ folded)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t id = keys[i] >> 1;
        uint8_t k = 0;
        if (i == 0 || (keys[i - 1] >> 1) != id) {
            float s = 0.0f;
            bool any = false, ident = false;
            for (int64_t t = i; t < count && (keys[t] >> 1) == id; ++t) {
                if (keys[t] & 1ULL) {
                    ident = true;
                } else {
                    const float w = vals ? vals[t] : 1.0f;
                    s = any ? s + w : w;
                    any = true;
                }
            }
            const double w2 = (any ? (double)s : 0.0) + (ident ? 1.0 : 0.0);
            folded[i] = w2;
            k = w2 != 0.0 ? 1 : 0;
        }
        keep[i] = k;
    }
}

// Stream compaction without a global position array: every block re-scans its chunk of keep flags and starts at the
// scanned count of the chunks before it.
__global__ void __launch_bounds__(kScanThreads) compact_kernel(const uint64_t *__restrict__ keys, const uint8_t *__restrict__ keep,
                                                               const int64_t *__restrict__ chunk_base, const double *__restrict__ folded, int64_t count,
                                                               int col_bits, int32_t *__restrict__ row_of, uint32_t *__restrict__ col_of,
                                                               uint32_t *__restrict__ ident_perm, double *__restrict__ w2)
{
    const uint64_t col_mask = (1ULL << col_bits) - 1ULL;
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    uint8_t k[kScanItems];
    int64_t mine = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        k[i] = base + i < count ? keep[base + i] : 0;
        mine += k[i];
    }
    int64_t total;
    int64_t p = block_exclusive_scan<uint8_t>(mine, &total) + chunk_base[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (!k[i]) continue;
        const uint64_t key = keys[base + i];
        row_of[p] = (int32_t)(key >> (col_bits + 1));
        col_of[p] = (uint32_t)((key >> 1) & col_mask);
        ident_perm[p] = (uint32_t)p;
        w2[p] = folded[base + i];
        ++p;
    }
}

// first position whose value is >= r in a sorted array (one thread per r in [0, n])
template <typename T>
__global__ void lower_bounds_kernel(const T *__restrict__ sorted, int64_t count, int64_t n, int64_t *__restrict__ out)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = count;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)sorted[mid] < r) lo = mid + 1;
            else hi = mid;
        }
        out[r] = lo;
    }
}

// deg[r] = sum of the row's merged weights in column order, one sequential float64 chain per row like scipy's A~.sum(1)
__global__ void row_degree_kernel(const int64_t *__restrict__ row_ptr, const double *__restrict__ w2, int64_t n, double *__restrict__ deg)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) s += w2[p];
        deg[r] = s;
    }
}

__global__ void export_kernel(const uint32_t *__restrict__ perm, const int32_t *__restrict__ row_of, const double *__restrict__ w2, int64_t nnz,
                              int32_t *__restrict__ indices, double *__restrict__ raw_w)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = perm[q];
        indices[q] = row_of[p];
        if (raw_w) raw_w[q] = w2[p];
    }
}

inline unsigned grid_for(int64_t count)
{
    int64_t b = (count + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

int bits_for(int64_t n)
{
    int b = 1;
    while ((1LL << b) < n) ++b;
    return b;
}

}  // namespace
}  // namespace sglb200

using namespace sglb200;

extern "C" void sglb200_adjacency_free(sglb200_adj_builder *b)
{
    if (!b) return;
    cudaFree(b->row_of);
    cudaFree(b->w2);
    cudaFree(b->col_sorted);
    cudaFree(b->perm);
    cudaFree(b->deg);
    delete b;
}

extern "C" int sglb200_adjacency_build(sglb200_adj_builder **out, int64_t n, int64_t n_edges, const int64_t *rows, const int64_t *cols,
                                       const float *weights, int add_identity, int64_t *nnz_out, void *stream_)
{
    clear_error();
    TraceRange range("sglb200_adjacency_build");
    SGL_REQUIRE(out != nullptr && nnz_out != nullptr, "adjacency_build: out / nnz_out is NULL");
    SGL_REQUIRE(n > 0 && n < (1LL << 31), "adjacency_build: n = %lld must be in [1, 2^31)", (long long)n);
    SGL_REQUIRE(n_edges >= 0 && (n_edges == 0 || (rows && cols)), "adjacency_build: bad edge list");
    const int64_t total = n_edges + (add_identity ? n : 0);
    SGL_REQUIRE(total > 0 && total < (1LL << 32), "adjacency_build: %lld entries, at most 2^32 - 1 are supported", (long long)total);
    const int st = check_device();
    if (st != SGLB200_OK) return st;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int col_bits = bits_for(n);
    const bool has_vals = weights != nullptr;

    sglb200_adj_builder *b = new (std::nothrow) sglb200_adj_builder();
    if (!b) {
        set_error("adjacency_build: out of host memory");
        return SGLB200_ERR_ALLOC;
    }
    b->n = n;
    uint64_t *keys_a = nullptr, *keys_b = nullptr;
    float *vals_a = nullptr, *vals_b = nullptr;
    uint8_t *keep = nullptr;
    uint32_t *col_a = nullptr, *perm_a = nullptr;
    int64_t *chunk_count = nullptr, *chunk_base = nullptr, *row_ptr = nullptr, *scratch = nullptr;
    int *bad = nullptr;
    SortWorkspace ws;
    int status = SGLB200_OK;
    cudaError_t e = cudaSuccess;
#define ADJ_TRY(expr)                                                                                                 \
    do {                                                                                                              \
        if (e == cudaSuccess) {                                                                                       \
            e = (expr);                                                                                               \
            if (e != cudaSuccess) set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e), __FILE__, __LINE__); \
        }                                                                                                             \
    } while (0)
    ADJ_TRY(cudaMalloc(&keys_a, sizeof(uint64_t) * total));
    ADJ_TRY(cudaMalloc(&keys_b, sizeof(uint64_t) * total));
    if (has_vals) {
        ADJ_TRY(cudaMalloc(&vals_a, sizeof(float) * total));
        ADJ_TRY(cudaMalloc(&vals_b, sizeof(float) * total));
    }
    ADJ_TRY(cudaMalloc(&bad, sizeof(int)));
    ADJ_TRY(sort_workspace_alloc(&ws, total));
    ADJ_TRY(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    if (e == cudaSuccess) {
        make_keys_kernel<<<grid_for(total), 256, 0, stream>>>(rows, cols, weights, n_edges, n, add_identity, col_bits, keys_a, vals_a, bad);
        ADJ_TRY(cudaGetLastError());
    }
    bool in_a = true;
    if (has_vals) ADJ_TRY((radix_sort<uint64_t, float, true>(keys_a, keys_b, vals_a, vals_b, total, 2 * col_bits + 1, &ws, &in_a, stream)));
    else ADJ_TRY((radix_sort<uint64_t, float, false>(keys_a, keys_b, nullptr, nullptr, total, 2 * col_bits + 1, &ws, &in_a, stream)));
    uint64_t *keys = in_a ? keys_a : keys_b, *spare = in_a ? keys_b : keys_a;   // the spare key buffer holds the folded weights
    const float *vals = has_vals ? (in_a ? vals_a : vals_b) : nullptr;
    const int64_t chunks = (total + kScanChunk - 1) / kScanChunk;
    ADJ_TRY(cudaMalloc(&keep, sizeof(uint8_t) * total));
    ADJ_TRY(cudaMalloc(&chunk_count, sizeof(int64_t) * chunks));
    ADJ_TRY(cudaMalloc(&chunk_base, sizeof(int64_t) * chunks));
    ADJ_TRY(cudaMalloc(&scratch, sizeof(int64_t) * scan_scratch_elems(chunks)));
    if (e == cudaSuccess) {
        fold_runs_kernel<<<grid_for(total), 256, 0, stream>>>(keys, vals, total, keep, reinterpret_cast<double *>(spare));
        scan_reduce_kernel<uint8_t><<<(unsigned)chunks, kScanThreads, 0, stream>>>(keep, total, chunk_count);
        ADJ_TRY(cudaGetLastError());
    }
    ADJ_TRY(exclusive_scan<int64_t>(chunk_count, chunks, chunk_base, scratch, stream));
    int64_t last_base = 0, last_count = 0;
    int h_bad = 0;
    ADJ_TRY(cudaMemcpyAsync(&last_base, chunk_base + chunks - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    ADJ_TRY(cudaMemcpyAsync(&last_count, chunk_count + chunks - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    ADJ_TRY(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ADJ_TRY(cudaStreamSynchronize(stream));
    if (e == cudaSuccess && h_bad) {
        set_error("adjacency_build: an edge endpoint lies outside [0, %lld)", (long long)n);
        status = SGLB200_ERR_INVALID;
    }
    const int64_t nnz = last_base + last_count;
    b->nnz = nnz;
    if (e == cudaSuccess && status == SGLB200_OK && nnz > 0) {
        ADJ_TRY(cudaMalloc(&b->row_of, sizeof(int32_t) * nnz));
        ADJ_TRY(cudaMalloc(&b->w2, sizeof(double) * nnz));
        ADJ_TRY(cudaMalloc(&col_a, sizeof(uint32_t) * nnz));
        ADJ_TRY(cudaMalloc(&perm_a, sizeof(uint32_t) * nnz));
        if (e == cudaSuccess) {
            compact_kernel<<<(unsigned)chunks, kScanThreads, 0, stream>>>(keys, keep, chunk_base, reinterpret_cast<const double *>(spare), total, col_bits,
                                                                        b->row_of, col_a, perm_a, b->w2);
            ADJ_TRY(cudaGetLastError());
        }
        // the sort buffers of the first phase are dead: release them before the transpose allocates its own (cudaFree waits
        // for the kernels above)
        cudaFree(keys_a);
        cudaFree(keys_b);
        cudaFree(vals_a);
        cudaFree(vals_b);
        cudaFree(keep);
        keys_a = keys_b = nullptr;
        vals_a = vals_b = nullptr;
        keep = nullptr;
        ADJ_TRY(cudaMalloc(&b->col_sorted, sizeof(uint32_t) * nnz));
        ADJ_TRY(cudaMalloc(&b->perm, sizeof(uint32_t) * nnz));
        ADJ_TRY(cudaMalloc(&b->deg, sizeof(double) * n));
        ADJ_TRY(cudaMalloc(&row_ptr, sizeof(int64_t) * (n + 1)));
        if (e == cudaSuccess) {
            lower_bounds_kernel<int32_t><<<grid_for(n + 1), 256, 0, stream>>>(b->row_of, nnz, n, row_ptr);
            row_degree_kernel<<<grid_for(n), 256, 0, stream>>>(row_ptr, b->w2, n, b->deg);
            ADJ_TRY(cudaGetLastError());
        }
        // transpose: stable sort of the (row, col)-ordered entries by column leaves every column's entries in row order
        bool t_in_a = true;
        ADJ_TRY((radix_sort<uint32_t, uint32_t, true>(col_a, b->col_sorted, perm_a, b->perm, nnz, col_bits, &ws, &t_in_a, stream)));
        if (e == cudaSuccess && t_in_a) {   // result sits in the temporaries: swap ownership
            uint32_t *t = b->col_sorted;
            b->col_sorted = col_a;
            col_a = t;
            t = b->perm;
            b->perm = perm_a;
            perm_a = t;
        }
        ADJ_TRY(cudaStreamSynchronize(stream));
    }
#undef ADJ_TRY
    cudaFree(keys_a);
    cudaFree(keys_b);
    cudaFree(vals_a);
    cudaFree(vals_b);
    cudaFree(keep);
    cudaFree(chunk_count);
    cudaFree(chunk_base);
    cudaFree(scratch);
    cudaFree(col_a);
    cudaFree(perm_a);
    cudaFree(row_ptr);
    cudaFree(bad);
    sort_workspace_free(&ws);
    if (e != cudaSuccess) status = SGLB200_ERR_CUDA;
    if (status != SGLB200_OK) {
        sglb200_adjacency_free(b);
        return status;
    }
    *out = b;
    *nnz_out = nnz;
    return SGLB200_OK;
}

extern "C" int sglb200_adjacency_export(sglb200_adj_builder *b, int64_t *indptr, int32_t *indices, double *raw_w, double *deg, void *stream_)
{
    clear_error();
    SGL_REQUIRE(b != nullptr && indptr != nullptr && indices != nullptr, "adjacency_export: NULL argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (b->nnz == 0) {
        SGL_CUDA_CHECK(cudaMemsetAsync(indptr, 0, sizeof(int64_t) * (b->n + 1), stream));
        if (deg) SGL_CUDA_CHECK(cudaMemsetAsync(deg, 0, sizeof(double) * b->n, stream));
        return SGLB200_OK;
    }
    lower_bounds_kernel<uint32_t><<<grid_for(b->n + 1), 256, 0, stream>>>(b->col_sorted, b->nnz, b->n, indptr);
    export_kernel<<<grid_for(b->nnz), 256, 0, stream>>>(b->perm, b->row_of, b->w2, b->nnz, indices, raw_w);
    SGL_CUDA_CHECK(cudaGetLastError());
    if (deg) SGL_CUDA_CHECK(cudaMemcpyAsync(deg, b->deg, sizeof(double) * b->n, cudaMemcpyDeviceToDevice, stream));
    return SGLB200_OK;
}
