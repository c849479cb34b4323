// common.cuh -- shared declarations of libsglb200 (B200 / sm_100a SGAP propagate + aggregate path).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "sglb200.h"

namespace sglb200 {

// thread-local error text behind sglb200_last_error()
void set_error(const char *fmt, ...);
void clear_error();

#define SGL_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            ::sglb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SGLB200_ERR_CUDA;                                                                      \
        }                                                                                                 \
    } while (0)

#define SGL_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::sglb200::set_error(__VA_ARGS__); \
            return SGLB200_ERR_INVALID;        \
        }                                      \
    } while (0)

// One merge-path tile schedule over the (row ends + non-zeros) item stream of a CSR operator.
// Boundary t is the coordinate (tile_row[t], tile_nnz[t]); tile t covers [boundary t, boundary t+1).
struct Schedule {
    bool built = false;
    int64_t n_tiles = 0;
    int32_t *tile_row = nullptr;    // n_tiles+1
    int64_t *tile_nnz = nullptr;    // n_tiles+1
    int32_t *carry_slot = nullptr;  // n_tiles : workspace slot for the partial sum of the row cut by the tile end, -1 none
    int32_t *tail_run = nullptr;    // n_tiles : run (cut row) that partial belongs to, -1 none
    int32_t *head_run = nullptr;    // n_tiles : run whose row this tile FINISHES (its first row is a continuation), -1 none
    uint32_t *run_count = nullptr;  // n_runs  : arrivals of the current hop (self-resetting), in-kernel fold
    int64_t n_runs = 0;             // rows cut across tiles
    int64_t n_slots = 0;            // carry partials in total
    int32_t *run_row = nullptr;     // n_runs
    int64_t *run_base = nullptr;    // n_runs : first slot (carriers in tile order)
    int32_t *run_len = nullptr;     // n_runs : number of carrier tiles
    // in-kernel fold bookkeeping (host): the arrival counters are only consistent if the tile ranges of one hop are issued
    // completely, in order, on one stream; `next_tile` is where the current hop is expected to continue (0 = no hop open)
    int64_t next_tile = 0;
    cudaStream_t last_stream = nullptr;   // a hop issued on another stream first waits for this one (host sync, rare)
    bool stream_known = false;
    std::vector<int64_t> run_last_tile;  // host copy, sorted: tile that finishes the cut row of run r (runs are
                                         // sorted by row, so a tile range maps to a contiguous run range)
};

}  // namespace sglb200

struct sglb200_graph {
    int device = 0;
    int sm_count = 148;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    int64_t *indptr = nullptr;  // n_rows+1, device
    int32_t *indices = nullptr; // nnz, device
    float *vals = nullptr;      // nnz (+ kStreamPad), device
    int32_t *idx_tag = nullptr; // nnz (+ kStreamPad): column id | bit 31 "last non-zero of its row" -- the stream the TMA hop
                                // kernel walks (spmm_tma.cu); valid when empty_rows == 0
    int32_t *idx_cold = nullptr;   // nnz (+ kStreamPad): column id | bit 30 "rarely referenced column" (build_cold_tags)
    int64_t cold_hub_rows = -1;    // the hub budget idx_cold was built for
    int cold_threshold = 0;        // reference count from which a column is a hub
    int64_t empty_rows = -1;    // rows without any non-zero (-1: not counted yet)
    int2 *pairs = nullptr;      // nnz (+ kStreamPad): (idx_tag[j], bits of vals[j]) interleaved -- ONE 8-byte load per non-zero
                                // for the lane-group kernel, whose 32/G groups each read their own stream (separate 4-byte
                                // arrays cost it as many L1 wavefronts as the feature gathers themselves)
    bool pairs_valid = false;   // false after the values were rewritten
    int tile_items = 0;
    int split_threshold = 0;
    sglb200::Schedule fast, exact;
    // fused degree normalisation (FAST mode of the fused driver): A^ = diag(row_scale) W diag(dr) (+ self_coef-weighted
    // teleport term), W = raw weights of (A+I)^T -- filled by sglb200_normalize_values next to the exact float32 `vals`
    bool has_scaling = false;
    bool unit_weights = false;     // every raw weight is 1: the hop streams 4 bytes per edge (column ids only)
    float *raw_w = nullptr;        // nnz (+ kStreamPad) raw weights, NULL when unit_weights
    float *row_scale = nullptr;    // n_rows: fl32((1-alpha) * deg^(r-1))
    float *col_scale = nullptr;    // n_cols: fl32(deg^-r)
    float *self_coef = nullptr;    // n_rows: fl32(alpha / deg^-r) (PPR only, else NULL)
    float *ping[2] = {nullptr, nullptr};  // internal hop slabs of the fused driver (n_rows * d each)
    size_t ping_floats = 0;
    float *aux = nullptr;          // 2 * n_rows floats (over-smoothing distance: |x| and the softmax denominator)
    float *carry_ws = nullptr;  // n_slots * ws_ld floats, grown on demand
    size_t carry_ws_floats = 0;
    float *stage[3] = {nullptr, nullptr, nullptr};  // propagate_host slabs
    size_t stage_floats = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_compute[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};
    size_t bytes_resident = 0;
};

namespace sglb200 {
constexpr int64_t kStreamPad = 128;  // zeroed elements behind indices / values: bulk copies of the last chunk stay in bounds
int build_stream_tags(sglb200_graph *g, cudaStream_t stream);
int build_cold_tags(sglb200_graph *g, int64_t hub_rows, cudaStream_t stream);
int build_stream_pairs(sglb200_graph *g, cudaStream_t stream);
int build_schedule(sglb200_graph *g, Schedule *s, int64_t split_threshold, cudaStream_t stream);
void free_schedule(Schedule *s);
int ensure_carry_ws(sglb200_graph *g, size_t floats, cudaStream_t stream);
struct Epilogue;
// one hop with every option of the fused driver: epi (NULL = plain store), raw_weights != 0 streams the raw weights
// (or nothing when they are all 1) instead of the normalised values
int spmm_launch_ex(sglb200_graph *g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode, int accumulate,
                   int64_t tile_begin, int64_t tile_end, const Epilogue *epi, int raw_weights, cudaStream_t stream);
int check_device();
}  // namespace sglb200
