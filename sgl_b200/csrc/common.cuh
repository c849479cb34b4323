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
    float *vals = nullptr;      // nnz, device
    int tile_items = 0;
    int split_threshold = 0;
    sglb200::Schedule fast, exact;
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
int build_schedule(sglb200_graph *g, Schedule *s, int64_t split_threshold, cudaStream_t stream);
void free_schedule(Schedule *s);
int ensure_carry_ws(sglb200_graph *g, size_t floats);
int check_device();
}  // namespace sglb200
