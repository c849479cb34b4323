// spmm.cu -- the hot path: one hop  Y = A^ X  (CSR x dense, fp32) and the K-hop drivers.
//
// Replaces the reference's CPU kernel FloatCSRMulDenseOMP (sgl/operators/csrc/matmul.c:23-40), its ctypes wrapper
// (sgl/operators/utils.py:10-40) and the hop loop of GraphOp.propagate (sgl/operators/base_op.py:29-36).
//
// Kernel shape (B200, sm_100a; the op is a random row gather, ~0.25 flop/byte, HBM/L2 bound):
//   * merge-path schedule (graph.cu): every warp owns `tile_items` consecutive items of the merged
//     (row ends + non-zeros) stream, so hubs and empty rows cost the same as anything else -- the reference's
//     static OpenMP row split (matmul.c:25) serialises on skewed graphs;
//   * a warp walks its non-zero range as ONE flat stream: 32 (col, val) pairs are fetched with one coalesced
//     streaming load each and published as 8-byte pairs in shared memory (one LDS.128 broadcast = two pairs for
//     every lane); all 32 lanes hold disjoint 128-bit column slices of the same output row, so one warp-wide
//     LDG.128 moves a whole 512 B feature row of X (d = 128);
//   * U gathered rows are in flight per warp before the first FMA consumes them (memory-level parallelism);
//   * accumulation in packed fp32 pairs (fma.rn.f32x2 -> FFMA2), bit-identical to scalar fmaf;
//   * a group of U non-zeros inside one row takes a check-free path; otherwise row boundaries are warp-uniform
//     compares against a shuffle-distributed prefetch of 32 row ends;
//   * per output element the additions happen in CSR order with one fused multiply-add per term: for rows that
//     are not cut (all rows in the EXACT schedule) this IS the reference's chain, bit for bit;
//   * rows cut across warps (FAST schedule, rows longer than split_threshold) leave partial sums in a small
//     workspace; the tile whose arrival completes a row's counter folds them in tile order in the kernel's
//     epilogue (deterministic; nothing waits on anything).  spmm_carry_fixup_kernel is the separate-launch form.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>

#include "spmm_common.cuh"
#include "trace.cuh"

namespace sglb200 {

// FLAG: the column stream is the tagged one (idx_tag: bit 31 = last non-zero of its row; graphs without empty rows), so a
// row ends where the stream says so: no row-pointer window, no countdown -- fewer live registers, more resident warps.
template <int VEC, int VPL, int U, bool ACCUM, int MINB, int PIPE, bool HINT = false, bool EPI = false, bool FLAG = false,
          bool RED = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB) spmm_flat_kernel(const __grid_constant__ SpmmParams p, const int32_t *__restrict__ idx_tag)
{
    static_assert(!(FLAG && ACCUM), "the flagged walk starts every chain from zero");
    static_assert(!(EPI && ACCUM), "the fused row flush starts every chain from zero");
    static_assert(32 % U == 0, "U must divide the batch of 32 non-zeros");
    static_assert(PIPE == 1 || PIPE == 2, "one or two groups of gathers in flight per warp");
    // (column id, value bits) of two batches of 32 non-zeros per warp: one LDS.128 broadcast hands every lane two
    // (row to gather, weight) pairs
    __shared__ int2 s_pairs[kWarpsPerBlock][64];
    const int lane = threadIdx.x & 31;
    const int64_t t = p.tile_begin + (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (t >= p.n_tiles) return;
    int2 *pairs = &s_pairs[threadIdx.x >> 5][0];

    int row = p.tile_row[t];
    const int row_end = p.tile_row[t + 1];
    const int64_t j0 = p.tile_nnz[t];
    const int n_nnz = (int)(p.tile_nnz[t + 1] - j0);
    const int n_rows = (int)p.n_rows;

    // this lane's column slices of the output row.  Slices beyond d are parked on a valid offset for the gathers
    // (same sector as an active lane: no extra traffic, no predicate in the hot loop) and masked at the stores.
    // Row addresses are ONE 32x32->64 multiply-add each: base pointer (kept in registers) + id * stride-in-bytes.
    const int col_block = blockIdx.y * (32 * VEC * VPL);
    bool act[VPL];
    int cofs_v[VPL];
    const char *xbase[VPL];
    char *ybase[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int cofs = col_block + (v * 32 + lane) * VEC;
        cofs_v[v] = cofs;
        act[v] = cofs < p.d;
        xbase[v] = reinterpret_cast<const char *>(p.X) + (size_t)(act[v] ? cofs : col_block) * sizeof(float);
        ybase[v] = reinterpret_cast<char *>(p.Y) + (size_t)cofs * sizeof(float);
        asm volatile("" : "+l"(xbase[v]));  // materialise: keeps the compiler from re-adding the parameter every load
        asm volatile("" : "+l"(ybase[v]));
    }
    // lean flush with a running aggregate: the first flushed row of a tile that FINISHES a cut row is only a piece of that
    // row -- the aggregate gets the folded row from whoever completes the fold, not this piece
    bool red_skip_first = false;
    if constexpr (RED) {
        if (p.fold) red_skip_first = p.head_run[t] >= 0;
    }
    // fused row flush: a tile whose first row continues a cut row parks that piece in the workspace (slot after the
    // row's carriers) instead of flushing it; the tile that completes the row's arrivals performs the one real flush
    int cont_slot = -1;
    if constexpr (EPI) {
        if (p.fold) {
            const int hr = p.head_run[t];
            if (hr >= 0) cont_slot = (int)(p.run_base[hr] + p.run_len[hr]);
        }
    }
    const uint32_t ldx_bytes = (uint32_t)p.ldx * (uint32_t)sizeof(float);  // strides < 2^30 elements (host check)
    const uint32_t ldy_bytes = (uint32_t)p.ldy * (uint32_t)sizeof(float);
    Slice<VEC> acc[VPL];

    RowPrefetch<VEC, EPI ? VPL : 1> pf;   // fused flush: what the flush of the CURRENT row will read (fetched at row start)
    auto init_acc = [&](int r) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            acc[v].zero();
            if (ACCUM) {
                if (act[v] && r < n_rows) acc[v].load(ybase[v] + (uint64_t)(uint32_t)r * ldy_bytes);
            }
        }
        if constexpr (EPI) {
            if (r < n_rows) prefetch_row<VEC, VPL>(p, (uint32_t)r, act, cofs_v, pf);
        }
    };

    // ends (relative to j0) of rows row_base .. row_base+31, one per lane, refilled every 32 finished rows
    auto load_row_ends = [&](int base) -> int {
        const int r = base + 1 + lane;
        if (r > n_rows) return INT_MAX;
        const int64_t rel = p.indptr[r] - j0;
        return rel > (int64_t)INT_MAX ? INT_MAX : (int)rel;
    };
    int row_base = row;
    int my_end = FLAG ? 0 : load_row_ends(row_base);
    // next_end: tile-relative position at which the current row ends; INT_MAX once the tile owns no further row end
    int next_end = FLAG ? 0 : __shfl_sync(kFull, my_end, 0);
    if (row >= row_end) next_end = INT_MAX;

    if (ACCUM) {
        // the chain of a row starts from the value already in Y only in the tile that holds the row start
        const bool starts_here = row < n_rows && p.indptr[row] == j0;
        init_acc(starts_here ? row : n_rows);
    } else {
        init_acc(n_rows);
        if constexpr (EPI) {
            if (row < n_rows) prefetch_row<VEC, VPL>(p, (uint32_t)row, act, cofs_v, pf);   // init_acc(n_rows) fetched nothing
        }
    }

    auto flush_row = [&]() {
        if constexpr (EPI) {
            if (cont_slot >= 0) {
                char *wrow = reinterpret_cast<char *>(p.carry_ws + (int64_t)cont_slot * p.ws_ld);
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (act[v]) acc[v].store(wrow + (size_t)cofs_v[v] * sizeof(float));
                cont_slot = -1;
            } else {
                AccPack<VEC, VPL> pack;
#pragma unroll
                for (int v = 0; v < VPL; ++v) pack.s[v] = acc[v];
                emit_row_call<VEC, VPL, 32>(&p, (uint32_t)row, pack, cofs_v[0], kFull, pf.row_scale, pf.z_scale, pf.self_coef);
            }
        } else {
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (act[v]) {
                    if (p.stream_y) acc[v].store_streaming(ybase[v] + (uint64_t)(uint32_t)row * ldy_bytes);
                    else acc[v].store(ybase[v] + (uint64_t)(uint32_t)row * ldy_bytes);
                    if constexpr (RED) {
                        if (!red_skip_first) red_row_slice<VEC>(p, (uint32_t)row, cofs_v[v], acc[v]);
                    }
                }
            if constexpr (RED) red_skip_first = false;
        }
        ++row;
        if constexpr (!FLAG) {
            if (row - row_base == 32) {
                row_base = row;
                my_end = load_row_ends(row_base);
            }
            const int e = __shfl_sync(kFull, my_end, row - row_base);
            next_end = row < row_end ? e : INT_MAX;
        }
        init_acc(row);  // every later row of the tile starts inside the tile
    };
    if constexpr (FLAG) {
        // the one row a flag cannot retire: a cut row whose non-zeros all lie in earlier tiles
        if (row < row_end && p.indptr[row + 1] == j0) flush_row();
    }

    const int32_t *cols = (FLAG || (HINT && idx_tag != nullptr) ? idx_tag : p.indices) + j0;
    const float *vals = p.vals + j0;
    uint64_t pol_hub = 0, pol_cold = 0;
    if constexpr (HINT) {
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_hub));
        if (p.cold_policy == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_cold));
        else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_cold));
    }
    const uint32_t hub_cols = p.hub_cols;
    const bool cold_tags = HINT && idx_tag != nullptr;

    // (col, val) of the next 32 non-zeros are fetched into registers one batch ahead of their publication
    int32_t col_next = 0;
    float val_next = 0.0f;
    const bool unit_w = p.vals == nullptr;   // fused normalisation on an unweighted graph: 4 bytes per edge
    if (lane < n_nnz) {
        col_next = load_stream_i32(cols + lane);
        val_next = unit_w ? 1.0f : load_stream_f32(vals + lane);
    }
    // lanes past the end of the tile publish (column 0, weight 0): a valid row to gather, never accumulated
    auto publish_batch = [&](int b) {
        __syncwarp();  // every lane is done with the batch that used this buffer two batches ago
        pairs[(b & 1) * 32 + lane] = make_int2(col_next, __float_as_int(val_next));
        __syncwarp();
        col_next = 0;
        val_next = 0.0f;
        const int nb = (b + 1) * 32 + lane;
        if (nb < n_nnz) {
            col_next = load_stream_i32(cols + nb);
            val_next = unit_w ? 1.0f : load_stream_f32(vals + nb);
        }
    };

    // software pipeline over groups of U non-zeros: the gathers of group g+1 are issued before group g is consumed,
    // so every warp keeps U..2U feature rows in flight without pausing for its own arithmetic
    auto issue = [&](Slice<VEC> (&buf)[U][VPL], int g) {
        const int pos = g * U;
        if ((pos & 31) == 0) publish_batch(pos >> 5);
        const int2 *pp = pairs + (pos & 63);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t raw = (uint32_t)pp[u].x;
            const uint32_t c = raw & (FLAG || HINT ? 0x3fffffffu : 0xffffffffu);
            if constexpr (HINT) {
                if (cold_tags) {   // bit 30 marks a rarely referenced column: evict_first, hubs load normally (warp-uniform)
                    if (raw & 0x40000000u) {
#pragma unroll
                        for (int v = 0; v < VPL; ++v) buf[u][v].load_hint(xbase[v] + (uint64_t)c * ldx_bytes, pol_cold);
                    } else {
#pragma unroll
                        for (int v = 0; v < VPL; ++v) buf[u][v].load_nc(xbase[v] + (uint64_t)c * ldx_bytes);
                    }
                } else {
                    const uint64_t pol = c < hub_cols ? pol_hub : pol_cold;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) buf[u][v].load_hint(xbase[v] + (uint64_t)c * ldx_bytes, pol);
                }
            } else {
#pragma unroll
                for (int v = 0; v < VPL; ++v) buf[u][v].load_nc(xbase[v] + (uint64_t)c * ldx_bytes);
            }
        }
    };
    auto consume = [&](Slice<VEC> (&buf)[U][VPL], int g) {
        const int pos = g * U;
        const int2 *pp = pairs + (pos & 63);
        if constexpr (FLAG) {
            const int valid = n_nnz - pos;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int2 cw = pp[u];
                if (u < valid) {   // warp-uniform
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v].fma(__int_as_float(cw.y), buf[u][v]);
                    if (cw.x < 0 && row < row_end) flush_row();
                }
            }
            return;
        }
        int left = next_end - pos;  // non-zeros of the current row still ahead, counted from the group start
        if (left >= U && n_nnz - pos >= U) {
            // the whole group belongs to the current row: no row-end checks, no predicates
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float w = __int_as_float(pp[u].y);
#pragma unroll
                for (int v = 0; v < VPL; ++v) acc[v].fma(w, buf[u][v]);
            }
        } else {
            const int valid = n_nnz - pos;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float w = __int_as_float(pp[u].y);
                while (u == left) {  // warp-uniform: the current row ends here (also retires empty rows)
                    flush_row();
                    left = next_end - pos;
                }
                if (u < valid) {     // false only in the padded tail of the last group
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v].fma(w, buf[u][v]);
                }
            }
        }
    };

    const int n_groups = (n_nnz + U - 1) / U;
    if constexpr (PIPE == 2) {
        Slice<VEC> buf_a[U][VPL], buf_b[U][VPL];
        if (n_groups > 0) issue(buf_a, 0);
#pragma unroll 1
        for (int g = 0; g < n_groups; g += 2) {
            const bool has_b = g + 1 < n_groups;
            if (has_b) issue(buf_b, g + 1);
            consume(buf_a, g);
            if (has_b) {
                if (g + 2 < n_groups) issue(buf_a, g + 2);
                consume(buf_b, g + 1);
            }
        }
    } else {
        Slice<VEC> buf[U][VPL];
#pragma unroll 1
        for (int g = 0; g < n_groups; ++g) {
            issue(buf, g);
            consume(buf, g);
        }
    }
    // rows (possibly empty ones) that end exactly at the end of the tile
    while (row < row_end) flush_row();
    // partial sum of the row cut by the tile end
    const int32_t slot = p.carry_slot[t];
    if (slot >= 0) {
        char *wrow = reinterpret_cast<char *>(p.carry_ws + (int64_t)slot * p.ws_ld);
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) acc[v].store(wrow + (size_t)(col_block + (v * 32 + lane) * VEC) * sizeof(float));
    }
    // in-kernel fold of cut rows (epilogue: nothing of the hot loop is live here).  A tile can be the FINISHER of one cut
    // row (its first row continued an earlier tile: the partial already sits in Y) and a CARRIER of another (the partial
    // just stored above).  Every participant arrives on the row's counter after a device-wide fence; the last arriver
    // adds the carriers in tile order to the finisher's partial -- the same sum, in the same order, whoever folds.
    if (p.fold) {
        const int finishes = p.head_run[t];
        const int carries = slot >= 0 ? p.tail_run[t] : -1;
        if (finishes >= 0 || carries >= 0) {
            __threadfence();
            __syncwarp();
#pragma unroll 1
            for (int role = 0; role < 2; ++role) {
                const int run = role == 0 ? finishes : carries;
                if (run < 0) continue;
                const int n_carriers = p.run_len[run];
                unsigned int seen = 0;
                if (lane == 0) seen = atomicAdd(p.run_count + run, 1u);
                seen = __shfl_sync(kFull, seen, 0);
                if (seen != (unsigned int)n_carriers) continue;  // someone else will arrive later and fold
                __threadfence();
                const char *ws0 = reinterpret_cast<const char *>(p.carry_ws + p.run_base[run] * p.ws_ld);
                const size_t ws_ld_bytes = (size_t)p.ws_ld * sizeof(float);
                const uint32_t out_row = (uint32_t)p.run_row[run];
                if constexpr (EPI) {
                    // finisher piece (slot n_carriers) + (c0 + c1 + ...): the SAME order as the plain kernel's fold, so a hop gives
                    // the same bits whichever flush it went through; then the one real flush of the row
                    Slice<VEC> sum[VPL];
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        sum[v].zero();
                        if (!act[v]) continue;
                        const size_t cb = (size_t)cofs_v[v] * sizeof(float);
                        Slice<VEC> part, carried;
                        carried.load_l2(ws0 + cb);
                        for (int u = 1; u < n_carriers; ++u) {
                            part.load_l2(ws0 + (size_t)u * ws_ld_bytes + cb);
                            carried.add(part);
                        }
                        sum[v].load_l2(ws0 + (size_t)n_carriers * ws_ld_bytes + cb);
                        sum[v].add(carried);
                    }
                    RowPrefetch<VEC, VPL> pf2;
                    prefetch_row<VEC, VPL>(p, out_row, act, cofs_v, pf2);
                    AccPack<VEC, VPL> pack;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) pack.s[v] = sum[v];
                    emit_row_call<VEC, VPL, 32>(&p, out_row, pack, cofs_v[0], kFull, pf2.row_scale, pf2.z_scale, pf2.self_coef);
                } else {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (!act[v]) continue;
                        const size_t cb = (size_t)(col_block + (v * 32 + lane) * VEC) * sizeof(float);
                        Slice<VEC> sum, part;
                        sum.load_l2(ws0 + cb);
                        for (int u = 1; u < n_carriers; ++u) {
                            part.load_l2(ws0 + (size_t)u * ws_ld_bytes + cb);
                            sum.add(part);
                        }
                        char *yp = ybase[v] + (uint64_t)out_row * ldy_bytes;
                        part.load_l2(yp);          // the finishing tile's partial
                        part.add(sum);             // Y = finisher + (c0 + c1 + ...): the order of the separate fold kernel
                        part.store(yp);
                        if constexpr (RED) red_row_slice<VEC>(p, out_row, cofs_v[v], part);
                    }
                }
                if (lane == 0) p.run_count[run] = 0u;  // ready for the next hop
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Ring variant of the hop kernel (float4 rows, d <= 256): gathered feature rows are staged with cp.async (LDGSTS.128)
// into a per-warp ring in shared memory instead of registers, so the bytes in flight per SM are bounded by shared
// memory (24 warps x 16 rows x 512 B = 192 KB) rather than by the register file.  Measured on B200 (profiles/): the
// L2->SM ingest of this gather scales with the rows in flight up to ~18.6 TB/s; the register-staged kernel holds 160
// rows/SM in flight (12 TB/s), this one up to 384.  Same schedule, same per-element fma order => same bits.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRingGroup = 8;  // rows committed per cp.async group

template <int VPL, int RING, bool ACCUM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3) spmm_ring_kernel(const SpmmParams p, int row_floats)
{
    constexpr int G = kRingGroup;
    constexpr int D = RING / G;  // groups in flight
    static_assert(RING % G == 0 && 32 % G == 0 && D >= 1, "ring geometry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t t = p.tile_begin + (int64_t)blockIdx.x * kWarpsPerBlock + wib;
    float *ring = reinterpret_cast<float *>(smem_raw) + (size_t)wib * RING * row_floats;
    int2 *pairs = reinterpret_cast<int2 *>(reinterpret_cast<float *>(smem_raw) + (size_t)kWarpsPerBlock * RING * row_floats) + wib * 64;
    if (t >= p.n_tiles) return;

    int row = p.tile_row[t];
    const int row_end = p.tile_row[t + 1];
    const int64_t j0 = p.tile_nnz[t];
    const int n_nnz = (int)(p.tile_nnz[t + 1] - j0);
    const int n_rows = (int)p.n_rows;

    bool act[VPL];
    const char *xbase[VPL];
    char *ybase[VPL];
    uint32_t soff[VPL];  // byte offset of this lane's slice inside a ring row
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int cofs = (v * 32 + lane) * 4;
        act[v] = cofs < p.d;
        soff[v] = (uint32_t)(act[v] ? cofs : 0) * 4u;
        xbase[v] = reinterpret_cast<const char *>(p.X) + soff[v];
        ybase[v] = reinterpret_cast<char *>(p.Y) + (size_t)cofs * sizeof(float);
        asm volatile("" : "+l"(xbase[v]));
        asm volatile("" : "+l"(ybase[v]));
    }
    const uint32_t ldx_bytes = (uint32_t)p.ldx * 4u, ldy_bytes = (uint32_t)p.ldy * 4u;
    const uint32_t row_bytes = (uint32_t)row_floats * 4u;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    float acc[VPL][4];

    auto init_acc = [&](int r) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[v][e] = 0.0f;
            if (ACCUM) {
                if (act[v] && r < n_rows)
                    load_plain<4>(acc[v], reinterpret_cast<const float *>(ybase[v] + (uint64_t)(uint32_t)r * ldy_bytes));
            }
        }
    };
    auto load_row_ends = [&](int base) -> int {
        const int r = base + 1 + lane;
        if (r > n_rows) return INT_MAX;
        const int64_t rel = p.indptr[r] - j0;
        return rel > (int64_t)INT_MAX ? INT_MAX : (int)rel;
    };
    int row_base = row;
    int my_end = load_row_ends(row_base);
    int next_end = __shfl_sync(kFull, my_end, 0);
    if (row >= row_end) next_end = INT_MAX;
    if (ACCUM) {
        const bool starts_here = row < n_rows && p.indptr[row] == j0;
        init_acc(starts_here ? row : n_rows);
    } else {
        init_acc(n_rows);
    }
    auto flush_row = [&]() {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) store_slice<4>(reinterpret_cast<float *>(ybase[v] + (uint64_t)(uint32_t)row * ldy_bytes), acc[v]);
        ++row;
        if (row - row_base == 32) {
            row_base = row;
            my_end = load_row_ends(row_base);
        }
        const int e = __shfl_sync(kFull, my_end, row - row_base);
        next_end = row < row_end ? e : INT_MAX;
        init_acc(row);
    };

    const int32_t *cols = p.indices + j0;
    const float *vals = p.vals + j0;
    int32_t col_next = 0;
    float val_next = 0.0f;
    if (lane < n_nnz) {
        col_next = load_stream_i32(cols + lane);
        val_next = load_stream_f32(vals + lane);
    }
    // publishes batch b (32 (col, val) pairs) to shared memory and prefetches batch b+1 into registers
    auto publish_batch = [&](int b) {
        pairs[(b & 1) * 32 + lane] = make_int2(col_next, __float_as_int(val_next));
        __syncwarp();
        col_next = 0;
        val_next = 0.0f;
        const int nb = (b + 1) * 32 + lane;
        if (nb < n_nnz) {
            col_next = load_stream_i32(cols + nb);
            val_next = load_stream_f32(vals + nb);
        }
    };
    // starts the asynchronous copies of the G rows of group g into their ring slots (padded positions copy row 0)
    auto issue_group = [&](int g) {
        const int pos = g * G;
        if ((pos & 31) == 0) publish_batch(pos >> 5);
        const int2 *pp = pairs + (pos & 63);
        const uint32_t slot0 = ring_s + (uint32_t)((g % D) * G) * row_bytes;
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const uint32_t c = (uint32_t)pp[u].x;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const char *src = xbase[v] + (uint64_t)c * ldx_bytes;
                const uint32_t dst = slot0 + (uint32_t)u * row_bytes + soff[v];
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
    };

    const int n_groups = (n_nnz + G - 1) / G;
#pragma unroll 1
    for (int g = 0; g < D && g < n_groups; ++g) {
        issue_group(g);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll 1
    for (int g = 0; g < n_groups; ++g) {
        // groups g+1 .. g+D-1 may stay in flight; group g must have landed
        if (g + D - 1 < n_groups) asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const int pos = g * G;
        const int2 *pp = pairs + (pos & 63);
        const float *slot = ring + (size_t)((g % D) * G) * row_floats;
        const int valid = n_nnz - pos;
        int left = next_end - pos;
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const float w = __int_as_float(pp[u].y);
            float x[VPL][4];
#pragma unroll
            for (int v = 0; v < VPL; ++v) load_plain<4>(x[v], slot + (size_t)u * row_floats + (soff[v] >> 2));
            while (u == left) {
                flush_row();
                left = next_end - pos;
            }
            const bool in_tile = u < valid;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[v][e] = in_tile ? fmaf(w, x[v][e], acc[v][e]) : acc[v][e];
        }
        __syncwarp();  // every lane is done with group g's slots and pairs before they are refilled
        if (g + D < n_groups) {
            issue_group(g + D);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    while (row < row_end) flush_row();
    const int32_t slot_id = p.carry_slot[t];
    if (slot_id >= 0) {
        float *wrow = p.carry_ws + (int64_t)slot_id * p.ws_ld;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) store_slice<4>(wrow + (v * 32 + lane) * 4, acc[v]);
    }
}

// folds the carried partial sums of every cut row into Y, in tile order (one warp per cut row)
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    spmm_carry_fixup_kernel(const int32_t *__restrict__ run_row, const int64_t *__restrict__ run_base,
                            const int32_t *__restrict__ run_len, int64_t n_runs, const float *__restrict__ ws,
                            int64_t ws_ld, float *__restrict__ Y, int64_t ldy, int d)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (r >= n_runs) return;
    const int64_t row = run_row[r];
    const int64_t base = run_base[r];
    const int len = run_len[r];
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float sum[VEC];
        load_plain<VEC>(sum, ws + base * ws_ld + c);
        int u = 1;
        for (; u + 4 <= len; u += 4) {  // four partials in flight, added strictly in tile order
            float part[4][VEC];
#pragma unroll
            for (int q = 0; q < 4; ++q) load_plain<VEC>(part[q], ws + (base + u + q) * ws_ld + c);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int e = 0; e < VEC; ++e) sum[e] += part[q][e];
        }
        for (; u < len; ++u) {
            float part[VEC];
            load_plain<VEC>(part, ws + (base + u) * ws_ld + c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) sum[e] += part[e];
        }
        float y[VEC];
        load_plain<VEC>(y, Y + row * ldy + c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) y[e] += sum[e];
        store_slice<VEC>(Y + row * ldy + c, y);
    }
}

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <typename Kern>
static cudaError_t launch_windowed(Kern kern, const SpmmParams &p, dim3 grid, cudaStream_t stream)
{
    // experiment: pin the first window_mb MB of X in L2 through a per-launch access policy window
    static int window_mb = env_int("SGLB200_L2_WINDOW_MB", 0);
    if (window_mb <= 0) {
        kern<<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr);
        return cudaGetLastError();
    }
    static bool limit_set = false;
    if (!limit_set) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)window_mb << 20);
        limit_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWarpsPerBlock * 32);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = const_cast<float *>(p.X);
    attr[0].val.accessPolicyWindow.num_bytes = (size_t)window_mb << 20;
    attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p, (const int32_t *)nullptr);
}

static thread_local const int32_t *g_cold_tags = nullptr;   // likewise: the cold-tagged column stream (SGLB200_COLD_HINT)
static thread_local const int32_t *g_flat_tags = nullptr;   // set by spmm_launch_ex for the duration of one launch (host, single thread per handle)

template <int VEC, int VPL, int U, int MINB, int PIPE = 1>
static cudaError_t launch_flat(const SpmmParams &p, bool accum, dim3 grid, cudaStream_t stream)
{
    const int32_t *tags = accum ? nullptr : g_flat_tags;
    if (!accum && p.epi.active) {
        // the fused flush keeps the next row's aggregate / scales in registers from the row start on: two CTAs per SM
        if (tags) spmm_flat_kernel<VEC, VPL, U, false, (MINB > 3 ? 3 : MINB), 1, false, true, true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, tags);
        else spmm_flat_kernel<VEC, VPL, U, false, (MINB > 3 ? 3 : MINB), 1, false, true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr);
        return cudaGetLastError();
    }
    if (!accum && p.red_agg) {   // lean flush + running aggregate (L2 reductions)
        spmm_flat_kernel<VEC, VPL, U, false, (MINB > 3 ? 3 : MINB), 1, false, false, false, true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr);
        return cudaGetLastError();
    }
    if (!accum && tags && p.hub_cols == 0) {
        static int minb4 = env_int("SGLB200_FLAG_MINB4", 0);
        if (VEC == 4 && VPL == 1 && minb4) spmm_flat_kernel<VEC, VPL, U, false, (VEC == 4 && VPL == 1 ? 4 : MINB), 1, false, false, true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, tags);
        else spmm_flat_kernel<VEC, VPL, U, false, MINB, 1, false, false, true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, tags);
        return cudaGetLastError();
    }
    if (!accum && (p.hub_cols > 0 || g_cold_tags)) {
        spmm_flat_kernel<VEC, VPL, U, false, MINB, PIPE, true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, g_cold_tags);
        return cudaGetLastError();
    }
    if (!accum) return launch_windowed(spmm_flat_kernel<VEC, VPL, U, false, MINB, PIPE>, p, grid, stream);
    if (accum) spmm_flat_kernel<VEC, VPL, U, true, (MINB > 3 ? 3 : MINB), 1><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr);
    else spmm_flat_kernel<VEC, VPL, U, false, MINB, PIPE><<<grid, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr);
    return cudaGetLastError();
}

// tuning knob for experiments: SGLB200_SPMM_VARIANT selects (rows in flight per warp, resident CTAs per SM) of the
// d = 128 kernel; unset = the measured best
static int spmm_variant()
{
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("SGLB200_SPMM_VARIANT");
        v = e ? atoi(e) : -1;
    }
    return v;
}

static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// choose the vector width and slices per lane: maximise the used fraction of the 32*VPL lane slots, prefer wide loads
static void pick_shape(int d, int max_vec, int *vec_out, int *vpl_out, int *col_blocks)
{
    double best_util = -1.0;
    int best_vec = 1, best_vpl = 1;
    for (int vec = max_vec; vec >= 1; vec >>= 1) {
        if (d % vec) continue;
        const int slots_needed = d / vec;
        int vpl = 1;
        while (vpl < 4 && 32 * vpl < slots_needed) vpl <<= 1;
        const int per_block = 32 * vpl;
        const int blocks = (slots_needed + per_block - 1) / per_block;
        const double util = (double)slots_needed / ((double)blocks * per_block);
        if (util > best_util + 1e-9) {
            best_util = util;
            best_vec = vec;
            best_vpl = vpl;
        }
    }
    *vec_out = best_vec;
    *vpl_out = best_vpl;
    const int per_block = 32 * best_vpl * best_vec;
    *col_blocks = (d + per_block - 1) / per_block;
}

int spmm_launch_tiles(sglb200_graph *g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode,
                      int accumulate, int64_t tile_begin, int64_t tile_end, cudaStream_t stream);
cudaError_t spmm_group_launch(const SpmmParams &p, bool accum, const int32_t *idx_tag, const int2 *pairs, cudaStream_t stream);  // spmm_group.cu
bool spmm_tma_eligible(const sglb200_graph *g, const float *X, int64_t ldx, int d);      // spmm_tma.cu
cudaError_t spmm_tma_launch(const sglb200_graph *g, const SpmmParams &p, cudaStream_t stream);

int spmm_launch(sglb200_graph *g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode, int accumulate,
                cudaStream_t stream)
{
    return spmm_launch_tiles(g, X, ldx, Y, ldy, d, mode, accumulate, 0, -1, stream);
}

// launches the hop on the tiles [tile_begin, tile_end) of the schedule (tile_end < 0: all) and folds the cut rows that
// FINISH inside that range; rows tile_row[tile_begin] .. tile_row[tile_end]-1 of Y are final afterwards
int spmm_launch_tiles(sglb200_graph *g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode,
                      int accumulate, int64_t tile_begin, int64_t tile_end, cudaStream_t stream)
{
    return spmm_launch_ex(g, X, ldx, Y, ldy, d, mode, accumulate, tile_begin, tile_end, nullptr, 0, stream);
}

int spmm_launch_ex(sglb200_graph *g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode, int accumulate,
                   int64_t tile_begin, int64_t tile_end, const Epilogue *epi, int raw_weights, cudaStream_t stream)
{
    SGL_REQUIRE(g != nullptr, "spmm: graph is NULL");
    SGL_REQUIRE(d >= 0, "spmm: negative feature width");
    if (g->n_rows == 0 || d == 0) return SGLB200_OK;
    SGL_REQUIRE(X != nullptr && (Y != nullptr || epi != nullptr), "spmm: X or Y is NULL");
    SGL_REQUIRE(ldx >= d && (Y == nullptr || ldy >= d), "spmm: row stride smaller than the feature width");
    SGL_REQUIRE(!(epi && accumulate), "spmm: the fused row flush cannot accumulate into Y");
    SGL_REQUIRE(!raw_weights || g->has_scaling, "spmm: the handle holds no raw weights (sglb200_normalize_values)");
    if (Y == nullptr) ldy = d;
    SGL_REQUIRE(ldx < (1LL << 30) && ldy < (1LL << 30), "spmm: row stride must be below 2^30 elements");
    SGL_REQUIRE(mode == SGLB200_MODE_FAST || mode == SGLB200_MODE_EXACT, "spmm: unknown mode %d", mode);
    Schedule *s = &g->fast;
    // accumulate starts every chain from the value already in Y: a cut row would race between the warp that reads
    // Y_in at the row start and the warp that stores the row end, so accumulation always runs on whole rows.
    if (mode == SGLB200_MODE_EXACT || accumulate) {
        if (!g->exact.built) {
            const int st = build_schedule(g, &g->exact, -1, stream);
            if (st != SGLB200_OK) return st;
        }
        s = &g->exact;
    }
    if (s->n_tiles == 0) return SGLB200_OK;
    {
        // experiment: column tiling -- the hop runs once per block of col_tile feature columns so that the hub rows
        // of one block (col_tile * 4 bytes each) fit L2 in larger numbers
        static int col_tile = env_int("SGLB200_COL_TILE", 0);
        if (col_tile > 0 && d > col_tile && d % col_tile == 0 && !accumulate && !epi && !raw_weights) {
            for (int c0 = 0; c0 < d; c0 += col_tile) {
                const int st = spmm_launch_tiles(g, X + c0, ldx, Y + c0, ldy, col_tile, mode, 0, tile_begin, tile_end, stream);
                if (st != SGLB200_OK) return st;
            }
            return SGLB200_OK;
        }
    }
    if (tile_end < 0 || tile_end > s->n_tiles) tile_end = s->n_tiles;
    if (tile_begin < 0) tile_begin = 0;
    if (tile_begin >= tile_end) return SGLB200_OK;

    int max_vec = 4;
    auto fits = [&](int64_t ld, const void *ptr, int v) { return ptr == nullptr || (ld % v == 0 && aligned(ptr, 4 * v)); };
    auto all_fit = [&](int v) {
        bool ok = d % v == 0 && fits(ldx, X, v) && fits(ldy, Y, v);
        if (epi) ok = ok && fits(epi->ldz, epi->Z, v) && fits(epi->ld_agg, epi->agg, v) && fits(epi->ld_self, epi->self_x, v) &&
                  fits(epi->ldx0, epi->x0, v) && fits(epi->ld_add, epi->add_rows, v);
        return ok;
    };
    if (!all_fit(4)) max_vec = 2;
    if (max_vec == 2 && !all_fit(2)) max_vec = 1;
    int vec, vpl, col_blocks;
    pick_shape(d, max_vec, &vec, &vpl, &col_blocks);
    // narrow float4 rows: 32/G tiles per warp, one lane group each (spmm_group.cu)
    const bool use_groups = max_vec == 4 && d <= 64 && spmm_variant() < 10 && env_int("SGLB200_GROUP", 1) != 0;
    if (use_groups) {
        vec = 4;
        vpl = 1;
        col_blocks = 1;
    }

    // TMA-staged kernel (spmm_tma.cu): float4 rows of 64..256 floats on graphs without empty rows
    bool use_tma = false;
    const bool lean_sum = epi && (epi->agg_op == EPI_AGG_SUM || epi->agg_op == EPI_AGG_WEIGHTED) && !epi->agg_init && !epi->row_scale &&
                          !epi->Z && epi->acc_scale == 0.0f;   // handled by the warp / lane-group kernels' lean flush
    if (!accumulate && !lean_sum && !use_groups && max_vec == 4 && d > 64 && d <= 128 && spmm_variant() < 10 && env_int("SGLB200_TMA", 0) != 0) {
        if (g->empty_rows < 0) {
            const int st = build_stream_tags(g, stream);
            if (st != SGLB200_OK) return st;
        }
        use_tma = spmm_tma_eligible(g, X, ldx, d);
        if (use_tma) {
            vec = 4;
            vpl = 1;
            col_blocks = 1;
        }
    }
    const int64_t ws_ld = (d + 3) & ~3;
    if (s->n_slots > 0) {
        const int st = ensure_carry_ws(g, (size_t)s->n_slots * (size_t)ws_ld, stream);
        if (st != SGLB200_OK) return st;
    }
    SpmmParams p = {};
    p.indptr = g->indptr;
    p.indices = g->indices;
    p.vals = raw_weights ? g->raw_w : g->vals;   // raw_w is NULL when every raw weight is 1
    p.tile_row = s->tile_row;
    p.tile_nnz = s->tile_nnz;
    p.carry_slot = s->carry_slot;
    p.tile_begin = tile_begin;
    p.n_tiles = tile_end;
    p.n_rows = g->n_rows;
    p.X = X;
    p.ldx = ldx;
    p.Y = Y;
    p.ldy = ldy;
    p.d = d;
    p.carry_ws = g->carry_ws;
    p.ws_ld = ws_ld;
    {
        // cut rows are folded inside the hop kernel by the last tile that arrives (default); SGLB200_FOLD=fixup keeps
        // the separate fold launch (also used by the cp.async ring variant, which does not carry the arrival logic)
        static int fold_mode = -1;
        if (fold_mode < 0) {
            const char *e = getenv("SGLB200_FOLD");
            fold_mode = (e && e[0] == 'f') ? 0 : 1;
        }
        // one counter per cut row: valid while a tile is one participant, i.e. a single column block (d <= 512)
        p.fold = (fold_mode == 1 && s->n_runs > 0 && spmm_variant() < 10 && col_blocks == 1) ? 1 : 0;
        if (p.fold) {
            // the arrival counters assume that one hop's tile ranges are issued completely, in order, on one stream
            // (include/sglb200.h, sglb200_spmm_tiles).  A hop that starts while an earlier one was left open (aborted
            // between ranges) gets fresh counters; a hop on another stream first waits for the previous fold to finish.
            if (s->next_tile != 0 && tile_begin != s->next_tile) {
                SGL_REQUIRE(tile_begin == 0, "spmm_tiles: tile ranges of one hop must be issued in order (expected tile %lld, got %lld)",
                            (long long)s->next_tile, (long long)tile_begin);
                if (s->stream_known && s->last_stream != stream) SGL_CUDA_CHECK(cudaStreamSynchronize(s->last_stream));
                SGL_CUDA_CHECK(cudaMemsetAsync(s->run_count, 0, sizeof(uint32_t) * (size_t)s->n_runs, stream));
            } else if (s->stream_known && s->last_stream != stream) {
                SGL_CUDA_CHECK(cudaStreamSynchronize(s->last_stream));   // rare: the previous hop's fold must have finished
            }
        }
        if (epi) {
            SGL_REQUIRE(s->n_runs == 0 || p.fold, "spmm: the fused row flush needs the in-kernel fold (d <= 512)");
            SGL_REQUIRE(col_blocks == 1, "spmm: the fused row flush supports feature widths up to 512");
            const bool only_running_sum = (epi->agg_op == EPI_AGG_SUM || epi->agg_op == EPI_AGG_WEIGHTED) && !epi->agg_init &&
                                          epi->agg_div == 0.0f && !epi->row_scale && !epi->self_coef && !epi->Z &&
                                          epi->acc_scale == 0.0f && Y != nullptr;
            if (only_running_sum) {
                // nothing but a running sum: the lean flush adds the row to the aggregate with one L2 reduction
                p.red_agg = epi->agg;
                p.ld_red = epi->ld_agg;
                p.red_w = epi->agg_w;
                p.red_weighted = epi->agg_op == EPI_AGG_WEIGHTED;
            } else {
                p.epi = *epi;
                p.epi.active = 1;
            }
        }
        p.tail_run = s->tail_run;
        p.head_run = s->head_run;
        p.run_row = s->run_row;
        p.run_base = s->run_base;
        p.run_len = s->run_len;
        p.run_count = s->run_count;
    }
    {
        // X (n_cols x d) should own L2 during a hop: stream Y past it unless Y is small enough to stay resident too
        static int force = -2;
        if (force == -2) {
            const char *e = getenv("SGLB200_STREAM_Y");
            force = e ? atoi(e) : -1;
        }
        p.stream_y = force >= 0 ? force : 1;
    }
    {
        static int hub = env_int("SGLB200_HUB_COLS", 0);
        static int cold = env_int("SGLB200_COLD_POLICY", 1);
        p.hub_cols = (uint32_t)hub;
        p.cold_policy = cold;
    }

    const dim3 grid((unsigned)((tile_end - tile_begin + kWarpsPerBlock - 1) / kWarpsPerBlock), (unsigned)col_blocks, 1);
    const bool acc = accumulate != 0;
    // the flagged column stream (row ends carried by bit 31) where the graph allows it: no empty rows, ids below 2^30
    const int32_t *tags = nullptr;
    if (!acc && env_int("SGLB200_FLAGS", 1) != 0) {
        if (g->empty_rows < 0) {
            const int st = build_stream_tags(g, stream);
            if (st != SGLB200_OK) return st;
        }
        if (g->empty_rows == 0 && g->n_cols < (1LL << 30)) tags = g->idx_tag;
    }
    // the warp kernel keeps its row-pointer windows by default: its check-free groups beat the per-non-zero flag test
    // (measured: arxiv 101 vs 111 us, products 5.3 vs 6.0 ms per hop); the lane-group and TMA kernels use the flags
    g_flat_tags = env_int("SGLB200_FLAT_FLAGS", 0) != 0 ? tags : nullptr;
    g_cold_tags = nullptr;
    if (!acc && !(epi && epi->active) && !p.red_agg && env_int("SGLB200_COLD_HINT", 0) != 0 && g->n_cols < (1LL << 30)) {
        // hub budget: the rows of X that should own L2 (default 48 MB of the 126 MB)
        const int64_t hub_rows = ((int64_t)env_int("SGLB200_HUB_MB", 48) << 20) / ((int64_t)ldx * 4);
        const int st = build_cold_tags(g, hub_rows, stream);
        if (st != SGLB200_OK) return st;
        g_cold_tags = g->idx_cold;
        p.cold_policy = env_int("SGLB200_COLD_POLICY", 1);
    }
    cudaError_t e = cudaSuccess;
#define SGL_SHAPE(V, L, UU, MB) \
    if (vec == V && vpl == L) e = launch_flat<V, L, UU, MB>(p, acc, grid, stream)
    const int variant = spmm_variant();
    const int row_floats = (d + 3) & ~3;
    if (vec == 4 && vpl <= 2 && col_blocks == 1 && (variant == -1 ? false : variant >= 10)) {
        // ring (cp.async) variant: variant 10 = 16-row ring, 11 = 24-row ring, 12 = 8-row ring
        const int ring_rows = variant == 11 ? 24 : (variant == 12 ? 8 : 16);
        const size_t smem = (size_t)kWarpsPerBlock * ring_rows * row_floats * sizeof(float) + kWarpsPerBlock * 64 * sizeof(int2);
#define SGL_RING(L, R, A)                                                                                          \
    do {                                                                                                           \
        static bool attr_done = false;                                                                             \
        if (!attr_done) {                                                                                          \
            e = cudaFuncSetAttribute(spmm_ring_kernel<L, R, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
            attr_done = (e == cudaSuccess);                                                                        \
        }                                                                                                          \
        if (e == cudaSuccess) {                                                                                    \
            spmm_ring_kernel<L, R, A><<<grid, kWarpsPerBlock * 32, smem, stream>>>(p, row_floats);                 \
            e = cudaGetLastError();                                                                                \
        }                                                                                                          \
    } while (0)
        if (vpl == 1) {
            if (ring_rows == 24) { if (acc) SGL_RING(1, 24, true); else SGL_RING(1, 24, false); }
            else if (ring_rows == 8) { if (acc) SGL_RING(1, 8, true); else SGL_RING(1, 8, false); }
            else { if (acc) SGL_RING(1, 16, true); else SGL_RING(1, 16, false); }
        } else {
            if (ring_rows == 24) { if (acc) SGL_RING(2, 24, true); else SGL_RING(2, 24, false); }
            else if (ring_rows == 8) { if (acc) SGL_RING(2, 8, true); else SGL_RING(2, 8, false); }
            else { if (acc) SGL_RING(2, 16, true); else SGL_RING(2, 16, false); }
        }
#undef SGL_RING
    }
    else if (use_tma) {
        e = spmm_tma_launch(g, p, stream);
    }
    else if (use_groups) {
        // the interleaved (flagged column, value) stream of the normalised values: one 8-byte load per non-zero
        const int2 *pairs = nullptr;
        if (tags && !acc && !p.epi.active && p.vals == g->vals && env_int("SGLB200_GROUP_PAIRS", 1) != 0) {
            const int st = build_stream_pairs(g, stream);
            if (st != SGLB200_OK) return st;
            pairs = g->pairs;
        }
        e = spmm_group_launch(p, acc, tags, pairs, stream);
    }
    else if (vec == 4 && vpl == 1) {
        switch (variant) {
        case 3: e = launch_flat<4, 1, 4, 3, 2>(p, acc, grid, stream); break;
        case 7: e = launch_flat<4, 1, 8, 2, 2>(p, acc, grid, stream); break;
        default: e = launch_flat<4, 1, 8, 3, 1>(p, acc, grid, stream); break;  // measured best on B200 (profiles/)
        }
    }
    else SGL_SHAPE(4, 2, 4, 3);
    else SGL_SHAPE(4, 4, 2, 2);
    else SGL_SHAPE(2, 1, 8, 4);
    else SGL_SHAPE(2, 2, 4, 3);
    else SGL_SHAPE(2, 4, 2, 2);
    else SGL_SHAPE(1, 1, 8, 4);
    else SGL_SHAPE(1, 2, 8, 3);
    else SGL_SHAPE(1, 4, 4, 2);
    else {
        set_error("spmm: no kernel shape for vec=%d vpl=%d", vec, vpl);
        return SGLB200_ERR_INVALID;
    }
#undef SGL_SHAPE
    SGL_CUDA_CHECK(e);
    if (p.fold) {
        s->next_tile = tile_end >= s->n_tiles ? 0 : tile_end;
        s->last_stream = stream;
        s->stream_known = true;
    }
    if (s->n_runs > 0 && !p.fold) {
        // runs are sorted by tile: those whose finishing tile lies in [tile_begin, tile_end) form one contiguous range
        const auto &last = s->run_last_tile;
        const int64_t r0 = std::lower_bound(last.begin(), last.end(), tile_begin) - last.begin();
        const int64_t r1 = std::lower_bound(last.begin(), last.end(), tile_end) - last.begin();
        const int64_t nr = r1 - r0;
        if (nr > 0) {
            const unsigned blocks = (unsigned)((nr + kWarpsPerBlock - 1) / kWarpsPerBlock);
            if (vec == 4)
                spmm_carry_fixup_kernel<4><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
                    s->run_row + r0, s->run_base + r0, s->run_len + r0, nr, g->carry_ws, ws_ld, Y, ldy, d);
            else if (vec == 2)
                spmm_carry_fixup_kernel<2><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
                    s->run_row + r0, s->run_base + r0, s->run_len + r0, nr, g->carry_ws, ws_ld, Y, ldy, d);
            else
                spmm_carry_fixup_kernel<1><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
                    s->run_row + r0, s->run_base + r0, s->run_len + r0, nr, g->carry_ws, ws_ld, Y, ldy, d);
            SGL_CUDA_CHECK(cudaGetLastError());
        }
    }
    return SGLB200_OK;
}

static int ensure_stage(sglb200_graph *g, size_t floats)
{
    if (floats <= g->stage_floats) return SGLB200_OK;
    // propagate_host is synchronous (it returns after both of its streams are idle): nothing of this handle is in flight
    for (int k = 0; k < 3; ++k) {
        cudaFree(g->stage[k]);
        g->stage[k] = nullptr;
    }
    g->bytes_resident -= 3 * g->stage_floats * sizeof(float);
    g->stage_floats = 0;
    for (int k = 0; k < 3; ++k) SGL_CUDA_CHECK(cudaMalloc(&g->stage[k], floats * sizeof(float)));
    g->stage_floats = floats;
    g->bytes_resident += 3 * floats * sizeof(float);
    return SGLB200_OK;
}

}  // namespace sglb200

using namespace sglb200;

extern "C" {

int sglb200_spmm(sglb200_graph_t g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode,
                 int accumulate, void *stream)
{
    clear_error();
    TraceRange range("sglb200_spmm");
    return spmm_launch(g, X, ldx, Y, ldy, d, mode, accumulate, (cudaStream_t)stream);
}

int sglb200_spmm_tiles(sglb200_graph_t g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode,
                       int64_t tile_begin, int64_t tile_end, void *stream)
{
    clear_error();
    return spmm_launch_tiles(g, X, ldx, Y, ldy, d, mode, 0, tile_begin, tile_end, (cudaStream_t)stream);
}

int sglb200_propagate(sglb200_graph_t g, float *const *hops, int64_t ld, int d, int K, int mode, void *stream)
{
    clear_error();
    SGL_REQUIRE(g && hops, "propagate: NULL argument");
    SGL_REQUIRE(K >= 0, "propagate: negative prop_steps");
    SGL_REQUIRE(g->n_rows == g->n_cols, "propagate: operator must be square (use sglb200_spmm for row partitions)");
    TraceRange range("sglb200_propagate");
    HopTimer timer("propagate");
    timer.mark((cudaStream_t)stream);
    for (int k = 1; k <= K; ++k) {
        SGL_REQUIRE(hops[k] != hops[k - 1], "propagate: hop %d aliases hop %d", k, k - 1);
        const int st = spmm_launch(g, hops[k - 1], ld, hops[k], ld, d, mode, 0, (cudaStream_t)stream);
        if (st != SGLB200_OK) return st;
        timer.mark((cudaStream_t)stream);
    }
    timer.report();
    return SGLB200_OK;
}

int sglb200_propagate_host(sglb200_graph_t g, const float *X, float *const *hops_out, int d, int K, int mode)
{
    clear_error();
    SGL_REQUIRE(g && X, "propagate_host: NULL argument");
    SGL_REQUIRE(K >= 0 && d >= 0, "propagate_host: negative size");
    SGL_REQUIRE(g->n_rows == g->n_cols, "propagate_host: operator must be square");
    SGL_REQUIRE(K == 0 || hops_out != nullptr, "propagate_host: hops_out is NULL");
    const size_t slab = (size_t)g->n_rows * (size_t)d;
    if (slab == 0 || K == 0) return SGLB200_OK;
    {
        const int st = ensure_stage(g, slab);
        if (st != SGLB200_OK) return st;
    }
    // three device slabs in a ring: hop k is computed into slab k%3 on the compute stream while hop k-1 drains to the
    // host on the copy stream; slab k%3 is reused only after the download of hop k-3 has finished.
    cudaStream_t cs = g->copy_stream;  // uploads + downloads
    cudaStream_t ks = nullptr;         // kernels on the legacy default stream of the caller's thread
    TraceRange range("sglb200_propagate_host");
    HopTimer timer("propagate_host");
    SGL_CUDA_CHECK(cudaMemcpyAsync(g->stage[0], X, slab * sizeof(float), cudaMemcpyHostToDevice, cs));
    SGL_CUDA_CHECK(cudaEventRecord(g->ev_copy[0], cs));
    SGL_CUDA_CHECK(cudaStreamWaitEvent(ks, g->ev_copy[0], 0));
    timer.mark(ks);
    for (int k = 1; k <= K; ++k) {
        const int dst = k % 3, src = (k - 1) % 3;
        if (k >= 3) SGL_CUDA_CHECK(cudaStreamWaitEvent(ks, g->ev_copy[dst], 0));  // download of hop k-3 done
        const int st = spmm_launch(g, g->stage[src], d, g->stage[dst], d, d, mode, 0, ks);
        if (st != SGLB200_OK) return st;
        timer.mark(ks);
        SGL_CUDA_CHECK(cudaEventRecord(g->ev_compute[dst], ks));
        if (hops_out[k - 1]) {
            SGL_CUDA_CHECK(cudaStreamWaitEvent(cs, g->ev_compute[dst], 0));
            SGL_CUDA_CHECK(cudaMemcpyAsync(hops_out[k - 1], g->stage[dst], slab * sizeof(float), cudaMemcpyDeviceToHost, cs));
        }
        SGL_CUDA_CHECK(cudaEventRecord(g->ev_copy[dst], cs));
    }
    SGL_CUDA_CHECK(cudaStreamSynchronize(ks));
    SGL_CUDA_CHECK(cudaStreamSynchronize(cs));
    timer.report();
    return SGLB200_OK;
}

}  // extern "C"
