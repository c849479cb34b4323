// spmm_group.cu -- the hop kernel for NARROW feature rows (d <= 64 floats): lane groups instead of whole warps.
//
// With d <= 64 a feature row is at most 16 float4 slices, so a 32-lane warp walking one tile (spmm.cu) idles half of its
// lanes or more -- and narrow rows are exactly what the feature-split partition of the multi-GPU path produces (d / 8
// columns per GPU) and what column-tiled hops use.  Here a warp is cut into 32 / G groups of G lanes
// (G = 4, 8, 16); every group is a "virtual warp" that owns its OWN merge-path tile and walks it exactly like the
// full-warp kernel does: flat (col, val) stream staged in shared memory, U gathered rows in flight, sequential fma
// chain per output element in CSR order (=> the same bits as the reference's matmul.c:30-37 chain for rows that are not
// cut), cut rows folded by the last arriver.  One warp-wide LDG.128 therefore fetches 32 / G different feature rows.
// Control flow that is warp-uniform in spmm.cu (row ends, loop bounds) is group-uniform here: shuffles and barriers
// name the group's lane mask only.
#include <stdlib.h>

#include "spmm_common.cuh"

namespace sglb200 {

// FLAG: the column stream is the tagged one (bit 31 = last non-zero of its row; graphs without empty rows): a row ends
// where the stream says so, which turns the group-divergent row flush (window refill + shuffles) into a few predicated
// instructions -- with 8 groups per warp the divergent flush was 70 % of the instruction stream.
// COOP (with FLAG): the (col, val) stream of ALL groups of a warp is fetched by the whole warp -- one coalesced load per
// group and batch (1-2 cache lines) instead of every group reading 16 bytes of its own stream per instruction (8 lines):
// the L1 tag stage was the busiest unit of the narrow-row hop (61 %), and a quarter of its work was this stream.
// PAIRS (with FLAG): (flagged column, value) come interleaved from ONE array: one 8-byte load per non-zero.  ncu on the
// separate-array form (products-shape, d=16): L1 data pipe 70 % busy, and the two 4-byte streams -- 8 cache lines touched
// per instruction for 128 useful bytes -- cost as many wavefronts as the gathers.
template <int G, int U, bool ACCUM, bool EPI = false, bool FLAG = false, int MINB = 3, bool COOP = false, bool PAIRS = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB) spmm_group_kernel(const __grid_constant__ SpmmParams p, const int32_t *__restrict__ idx_tag,
                                                                            const int2 *__restrict__ pair_stream)
{
    static_assert(!COOP || FLAG, "the cooperative fetch walks the flagged stream");
    static_assert(!PAIRS || (FLAG && !COOP), "the pair stream is the flagged one");
    static_assert(!(FLAG && ACCUM), "the flagged walk starts every chain from zero");
    static_assert(!(EPI && ACCUM), "the fused row flush starts every chain from zero");
    constexpr int NG = 32 / G;             // tiles per warp
    constexpr int BATCH = G == 4 ? 16 : 32;  // (col, val) pairs published per step (two batches live per group)
    constexpr int R = BATCH / G;           // pairs each lane fetches per batch
    constexpr int PSTRIDE = 2 * BATCH + 2; // int2 per group: +2 staggers the groups over the banks (the 32/G groups read
                                           // the same position of their own buffers in one LDS.64)
    static_assert(G == 4 || G == 8 || G == 16, "group width");
    static_assert(BATCH % U == 0, "U must divide the batch");
    __shared__ int2 s_pairs[kWarpsPerBlock][NG][PSTRIDE];
    __shared__ int64_t s_j0[COOP ? kWarpsPerBlock : 1][NG];
    __shared__ int s_nnz[COOP ? kWarpsPerBlock : 1][NG];
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);   // lane inside the group
    const int gid = lane / G;        // group inside the warp
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (gid * G);
    const int64_t t_raw = p.tile_begin + ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * NG + gid;
    if (!COOP && t_raw >= p.n_tiles) return;   // whole groups leave; nothing below synchronises beyond the group
    // cooperative fetch: groups past the last tile stay (they help loading) and walk an empty tile
    const bool live = t_raw < p.n_tiles;
    const int64_t t = live ? t_raw : p.n_tiles - 1;
    int2 *pairs = &s_pairs[threadIdx.x >> 5][gid][0];

    int row = live ? p.tile_row[t] : 0;
    const int row_end = live ? p.tile_row[t + 1] : 0;
    const int64_t j0 = p.tile_nnz[t];
    const int n_nnz = live ? (int)(p.tile_nnz[t + 1] - j0) : 0;
    const int n_rows = (int)p.n_rows;

    const int cofs = gl * 4;
    const bool act = cofs < p.d;
    const char *xbase = reinterpret_cast<const char *>(p.X) + (size_t)(act ? cofs : 0) * sizeof(float);
    char *ybase = reinterpret_cast<char *>(p.Y) + (size_t)cofs * sizeof(float);
    const uint32_t ldx_bytes = (uint32_t)p.ldx * (uint32_t)sizeof(float);
    const uint32_t ldy_bytes = (uint32_t)p.ldy * (uint32_t)sizeof(float);
    Slice<4> accs[1];
    Slice<4> &acc = accs[0];
    const bool acts[1] = {act};
    const int cofss[1] = {cofs};
    bool red_skip_first = false;   // lean flush + running aggregate: see spmm_flat_kernel
    if constexpr (!EPI) {
        if (p.red_agg && p.fold && live) red_skip_first = p.head_run[t] >= 0;
    }
    int cont_slot = -1;   // fused row flush: see spmm_flat_kernel
    if constexpr (EPI) {
        if (p.fold && live) {
            const int hr = p.head_run[t];
            if (hr >= 0) cont_slot = (int)(p.run_base[hr] + p.run_len[hr]);
        }
    }

    RowPrefetch<4, 1> pf;
    auto init_acc = [&](int r) {
        acc.zero();
        if (ACCUM) {
            if (act && r < n_rows) acc.load(ybase + (uint64_t)(uint32_t)r * ldy_bytes);
        }
        if constexpr (EPI) {
            if (r < n_rows) prefetch_row<4, 1>(p, (uint32_t)r, acts, cofss, pf);
        }
    };
    // ends (relative to j0) of rows row_base .. row_base+G-1, one per lane of the group
    auto load_row_ends = [&](int base) -> int {
        const int r = base + 1 + gl;
        if (r > n_rows) return INT_MAX;
        const int64_t rel = p.indptr[r] - j0;
        return rel > (int64_t)INT_MAX ? INT_MAX : (int)rel;
    };
    int row_base = row;
    int my_end = FLAG ? 0 : load_row_ends(row_base);
    int next_end = FLAG ? 0 : __shfl_sync(gmask, my_end, 0, G);
    if (row >= row_end) next_end = INT_MAX;
    if (ACCUM) {
        const bool starts_here = row < n_rows && p.indptr[row] == j0;
        init_acc(starts_here ? row : n_rows);
    } else {
        init_acc(n_rows);
        if constexpr (EPI) {
            if (row < n_rows) prefetch_row<4, 1>(p, (uint32_t)row, acts, cofss, pf);   // init_acc(n_rows) fetched nothing
        }
    }
    auto flush_row = [&]() {
        if constexpr (EPI) {
            if (cont_slot >= 0) {
                if (act) acc.store(reinterpret_cast<char *>(p.carry_ws + (int64_t)cont_slot * p.ws_ld) + (size_t)cofs * sizeof(float));
                cont_slot = -1;
            } else {
                AccPack<4, 1> pack;
                pack.s[0] = acc;
                emit_row_call<4, 1, G>(&p, (uint32_t)row, pack, cofs, gmask, pf.row_scale, pf.z_scale, pf.self_coef);
            }
        } else {
            if (act) {
                if (p.stream_y) acc.store_streaming(ybase + (uint64_t)(uint32_t)row * ldy_bytes);
                else acc.store(ybase + (uint64_t)(uint32_t)row * ldy_bytes);
                if (p.red_agg && !red_skip_first) red_row_slice<4>(p, (uint32_t)row, cofs, acc);
            }
            red_skip_first = false;
        }
        ++row;
        if constexpr (!FLAG) {
            if (row - row_base == G) {
                row_base = row;
                my_end = load_row_ends(row_base);
            }
            const int e = __shfl_sync(gmask, my_end, row - row_base, G);
            next_end = row < row_end ? e : INT_MAX;
        }
        init_acc(row);
    };

    if constexpr (FLAG) {
        // the one row a flag cannot retire: a cut row whose non-zeros all lie in earlier tiles (the boundary fell between
        // its last non-zero and its end marker) -- this tile finishes it with an empty piece
        if (row < row_end && p.indptr[row + 1] == j0) flush_row();
    }
    const int32_t *cols = (FLAG ? idx_tag : p.indices) + j0;
    const float *vals = p.vals + j0;
    // every lane of the group fetches R of the next 32 (col, val) pairs (interleaved: the group reads G consecutive
    // elements per instruction), one batch ahead of their publication
    int32_t col_next[R];
    float val_next[R];
    // cooperative form: slot q of this lane belongs to group q; BATCH == 16: lanes 0-15 hold columns, 16-31 values;
    // BATCH == 32: every lane holds one column (coop_c) and one value (coop_v) per group
    constexpr int CW = COOP ? NG : 1;
    int32_t coop_c[CW];
    int32_t coop_v[(COOP && BATCH == 32) ? NG : 1];
    int n_groups = (n_nnz + U - 1) / U;
    if constexpr (COOP) {
        if (gl == 0) {
            s_j0[threadIdx.x >> 5][gid] = j0;
            s_nnz[threadIdx.x >> 5][gid] = n_nnz;
        }
        // warp-uniform trip count: every lane takes part in every batch's fetch
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_groups = max(n_groups, __shfl_xor_sync(kFull, n_groups, o));
        __syncwarp();
    }
    auto fetch_batch = [&](int b) {
        if constexpr (COOP) {
            const int e = lane & (BATCH - 1);
            const int idx = b * BATCH + e;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                const int64_t base = s_j0[threadIdx.x >> 5][q];
                const bool ok = idx < s_nnz[threadIdx.x >> 5][q];
                if constexpr (BATCH == 16) {
                    int32_t v = 0;
                    if (ok) {
                        if (lane < 16) v = __ldg(idx_tag + base + idx);
                        else v = p.vals ? __float_as_int(__ldg(p.vals + base + idx)) : __float_as_int(1.0f);
                    }
                    coop_c[q] = v;
                } else {
                    coop_c[q] = ok ? __ldg(idx_tag + base + idx) : 0;
                    coop_v[q] = ok ? (p.vals ? __float_as_int(__ldg(p.vals + base + idx)) : __float_as_int(1.0f)) : 0;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int nb = b * BATCH + i * G + gl;
                col_next[i] = 0;
                val_next[i] = 0.0f;
                if (nb < n_nnz) {
                    if constexpr (PAIRS) {
                        const int2 pr = __ldg(pair_stream + j0 + nb);
                        col_next[i] = pr.x;
                        val_next[i] = __int_as_float(pr.y);
                    } else {
                        col_next[i] = __ldg(cols + nb);
                        val_next[i] = p.vals ? __ldg(vals + nb) : 1.0f;
                    }
                }
            }
        }
    };
    fetch_batch(0);
    auto publish_batch = [&](int b) {
        if constexpr (COOP) {
            __syncwarp();   // every group is done with the batch that used this buffer two batches ago
            const int e = lane & (BATCH - 1);
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                int32_t *dst = reinterpret_cast<int32_t *>(&s_pairs[threadIdx.x >> 5][q][(b & 1) * BATCH + e]);
                if constexpr (BATCH == 16) {
                    dst[lane >> 4] = coop_c[q];          // .x for the column lanes, .y for the value lanes
                } else {
                    dst[0] = coop_c[q];
                    dst[1] = coop_v[q];
                }
            }
            __syncwarp();
            fetch_batch(b + 1);
        } else {
            __syncwarp(gmask);  // the group is done with the batch that used this buffer two batches ago
#pragma unroll
            for (int i = 0; i < R; ++i) pairs[(b & 1) * BATCH + i * G + gl] = make_int2(col_next[i], __float_as_int(val_next[i]));
            __syncwarp(gmask);
            fetch_batch(b + 1);
        }
    };

#pragma unroll 1
    for (int g = 0; g < n_groups; ++g) {
        const int pos = g * U;
        if (pos % BATCH == 0) publish_batch(pos / BATCH);
        const int2 *pp = pairs + (pos % (2 * BATCH));
        Slice<4> buf[U];
#pragma unroll
        for (int u = 0; u < U; ++u) buf[u].load_nc(xbase + (uint64_t)((uint32_t)pp[u].x & (FLAG ? 0x3fffffffu : 0xffffffffu)) * ldx_bytes);
        const int valid = n_nnz - pos;
        if constexpr (FLAG) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int2 cw = pp[u];
                if (u < valid) {
                    acc.fma(__int_as_float(cw.y), buf[u]);
                    if (cw.x < 0 && row < row_end) flush_row();
                }
            }
        } else {
            int left = next_end - pos;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float w = __int_as_float(pp[u].y);
                while (u == left) {  // group-uniform: the current row ends here (also retires empty rows)
                    flush_row();
                    left = next_end - pos;
                }
                if (u < valid) acc.fma(w, buf[u]);
            }
        }
    }
    while (row < row_end) flush_row();
    if (COOP && !live) return;
    const int32_t slot = p.carry_slot[t];
    if (slot >= 0 && act) {
        char *wrow = reinterpret_cast<char *>(p.carry_ws + (int64_t)slot * p.ws_ld);
        acc.store(wrow + (size_t)cofs * sizeof(float));
    }
    // in-kernel fold of cut rows, as in spmm_flat_kernel, one group per participant
    if (p.fold) {
        const int finishes = p.head_run[t];
        const int carries = slot >= 0 ? p.tail_run[t] : -1;
        if (finishes >= 0 || carries >= 0) {
            __threadfence();
            __syncwarp(gmask);
#pragma unroll 1
            for (int role = 0; role < 2; ++role) {
                const int run = role == 0 ? finishes : carries;
                if (run < 0) continue;
                const int n_carriers = p.run_len[run];
                unsigned int seen = 0;
                if (gl == 0) seen = atomicAdd(p.run_count + run, 1u);
                seen = __shfl_sync(gmask, seen, 0, G);
                if (seen != (unsigned int)n_carriers) continue;
                __threadfence();
                if constexpr (EPI) {
                    const char *ws0 = reinterpret_cast<const char *>(p.carry_ws + p.run_base[run] * p.ws_ld);
                    const size_t ws_ld_bytes = (size_t)p.ws_ld * sizeof(float);
                    const size_t cb = (size_t)cofs * sizeof(float);
                    Slice<4> sums[1], part;
                    sums[0].zero();
                    if (act) {   // finisher piece + (c0 + c1 + ...): the plain fold's order
                        Slice<4> carried;
                        carried.load_l2(ws0 + cb);
                        for (int u = 1; u < n_carriers; ++u) {
                            part.load_l2(ws0 + (size_t)u * ws_ld_bytes + cb);
                            carried.add(part);
                        }
                        sums[0].load_l2(ws0 + (size_t)n_carriers * ws_ld_bytes + cb);
                        sums[0].add(carried);
                    }
                    RowPrefetch<4, 1> pf2;
                    prefetch_row<4, 1>(p, (uint32_t)p.run_row[run], acts, cofss, pf2);
                    AccPack<4, 1> pack;
                    pack.s[0] = sums[0];
                    emit_row_call<4, 1, G>(&p, (uint32_t)p.run_row[run], pack, cofs, gmask, pf2.row_scale, pf2.z_scale, pf2.self_coef);
                } else if (act) {
                    const char *ws0 = reinterpret_cast<const char *>(p.carry_ws + p.run_base[run] * p.ws_ld);
                    const size_t ws_ld_bytes = (size_t)p.ws_ld * sizeof(float);
                    const size_t cb = (size_t)cofs * sizeof(float);
                    Slice<4> sum, part;
                    sum.load_l2(ws0 + cb);
                    for (int u = 1; u < n_carriers; ++u) {
                        part.load_l2(ws0 + (size_t)u * ws_ld_bytes + cb);
                        sum.add(part);
                    }
                    char *yp = ybase + (uint64_t)(uint32_t)p.run_row[run] * ldy_bytes;
                    part.load_l2(yp);
                    part.add(sum);
                    part.store(yp);
                    if (p.red_agg) red_row_slice<4>(p, (uint32_t)p.run_row[run], cofs, part);
                }
                if (gl == 0) p.run_count[run] = 0u;
            }
        }
    }
}

template <int G, int U>
static cudaError_t launch_group(const SpmmParams &p, bool accum, const int32_t *idx_tag, const int2 *pairs, cudaStream_t stream)
{
    constexpr int NG = 32 / G;
    const int64_t tiles = p.n_tiles - p.tile_begin;
    const unsigned blocks = (unsigned)((tiles + (int64_t)kWarpsPerBlock * NG - 1) / ((int64_t)kWarpsPerBlock * NG));
    static const char *coop_env = getenv("SGLB200_GROUP_COOP");
    const bool coop = coop_env && atoi(coop_env) == 1;   // measured and rejected as default: d=16 1139 vs 1044 us (products), 799 vs 728 (rmat22)
    if (accum) spmm_group_kernel<G, U, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr, nullptr);
    else if (p.epi.active && idx_tag) spmm_group_kernel<G, U, false, true, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, idx_tag, nullptr);
    else if (idx_tag && U == 8 && coop) spmm_group_kernel<G, U, false, false, true, 3, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, idx_tag, nullptr);
    else if (p.epi.active) spmm_group_kernel<G, U, false, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr, nullptr);
    else if (idx_tag && U == 4) spmm_group_kernel<G, U, false, false, true, 4><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, idx_tag, nullptr);
    else if (idx_tag && pairs) spmm_group_kernel<G, U, false, false, true, 3, false, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, idx_tag, pairs);
    else if (idx_tag) spmm_group_kernel<G, U, false, false, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, idx_tag, nullptr);
    else spmm_group_kernel<G, U, false><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p, nullptr, nullptr);
    return cudaGetLastError();
}

// d <= 64, float4 rows: picks the narrowest group that holds a row
// idx_tag: the flagged column stream (NULL: walk with row-pointer windows -- graphs with empty rows, accumulate)
cudaError_t spmm_group_launch(const SpmmParams &p, bool accum, const int32_t *idx_tag, const int2 *pairs, cudaStream_t stream)
{
    const int slices = (p.d + 3) / 4;
    static const char *u_env = getenv("SGLB200_GROUP_U");
    const bool u4 = u_env && atoi(u_env) == 4 && !accum && !p.epi.active && idx_tag;   // experiment: 4 rows in flight, 32 warps / SM
    if (slices <= 4) return u4 ? launch_group<4, 4>(p, accum, idx_tag, pairs, stream) : launch_group<4, 8>(p, accum, idx_tag, pairs, stream);
    if (slices <= 8) return u4 ? launch_group<8, 4>(p, accum, idx_tag, pairs, stream) : launch_group<8, 8>(p, accum, idx_tag, pairs, stream);
    return u4 ? launch_group<16, 4>(p, accum, idx_tag, pairs, stream) : launch_group<16, 8>(p, accum, idx_tag, pairs, stream);
}

}  // namespace sglb200
