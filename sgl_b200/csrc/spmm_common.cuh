// spmm_common.cuh -- pieces shared by the hop kernels (spmm.cu: one warp per tile; spmm_group.cu: lane groups per tile).
#pragma once

#include <limits.h>

#include "common.cuh"

// L1 qualifier of the hinted gathers (experiment paths only): "" = default allocation, ".L1::no_allocate" to bypass.
// Measured the same either way on products-shape (profiles/r02_cold_tags.txt).
#ifndef SGL_HINT_L1
#define SGL_HINT_L1 ""
#endif

namespace sglb200 {

constexpr int kWarpsPerBlock = 8;
constexpr unsigned kFull = 0xffffffffu;

// What happens to a finished output row beyond the plain store into Y (all optional; `active` gates the whole block so
// that the plain hop keeps its lean row flush).  Used by the fused K-hop driver (sglb200_propagate_fused):
//   y      = row_scale[i] * acc                                 degree normalisation fused in-kernel: dL_i (utils.py:79-87)
//   y     += self_coef[i] * self_x[i, :]                        PPR teleport term alpha * x_i (ppr_graph_op.py:19)
//   Y[i]   = y                                                  per-hop store (skipped when Y == NULL)
//   Z[i]   = z_scale[i] * y                                     the NEXT hop's input, pre-scaled by dR_i
//   agg[i] = combine(agg[i], y)                                 cross-hop aggregation (message_op/*.py) as a running update
enum { EPI_AGG_NONE = 0, EPI_AGG_SUM = 1, EPI_AGG_WEIGHTED = 2, EPI_AGG_MAX = 3, EPI_AGG_MIN = 4, EPI_AGG_OSD = 5 };
struct Epilogue {
    int active;
    const float *row_scale;
    const float *self_coef;
    const float *self_x;
    int64_t ld_self;
    // label propagation (tricks/utils.py:54-56):  y = clamp(acc_scale * y + add_rows[i, :], lo, hi); acc_scale == 0: unused
    float acc_scale;
    const float *add_rows;
    int64_t ld_add;
    int clamp;
    float clamp_lo, clamp_hi;
    float *Z;
    int64_t ldz;
    const float *z_scale;
    int agg_op;
    int agg_init;     // 1: first included hop -- agg is written, not read
    float agg_w;      // WEIGHTED: the hop's weight
    float agg_div;    // != 0: divide after the update (mean: hop count; applied at the last included hop)
    float *agg;
    int64_t ld_agg;
    // over-smoothing distance (NAFS): agg holds the running numerator sum_k e^{c_k} y_k, den the denominator
    const float *x0;
    int64_t ldx0;
    const float *x0_norm;  // |x_i| + 1e-10
    float *den;
    int osd_final;         // 1: last hop -- agg = num / den
};

struct SpmmParams {
    const int64_t *indptr;
    const int32_t *indices;
    const float *vals;
    const int32_t *tile_row;
    const int64_t *tile_nnz;
    const int32_t *carry_slot;
    int64_t tile_begin;  // first tile of this launch
    int64_t n_tiles;     // one past the last tile of this launch
    int64_t n_rows;
    const float *X;
    int64_t ldx;
    float *Y;
    int64_t ldy;
    int d;
    float *carry_ws;
    int64_t ws_ld;
    int stream_y;  // 1: output rows are stored with the streaming (evict-first) policy
    // L2 residency control for gathered rows: columns below hub_cols are loaded with an evict_last policy, the
    // rest with cold_policy (0 = no hint, 1 = evict_first); hub_cols == 0 disables the hints
    uint32_t hub_cols;
    int cold_policy;
    // in-kernel fold of cut rows (fold != 0): every tile that holds a piece of a cut row stores its partial in the
    // workspace and arrives on the row's counter; the LAST arriver adds the partials in tile order and writes Y
    int fold;
    const int32_t *tail_run;
    const int32_t *head_run;
    const int32_t *run_row;
    const int64_t *run_base;
    const int32_t *run_len;
    unsigned int *run_count;
    // running sum / mean / weighted aggregate on the plain (lean) row flush: every finished row is also added to
    // red_agg[row, :] (scaled by red_w when red_weighted) with one fire-and-forget L2 reduction; NULL: off
    float *red_agg;
    int64_t ld_red;
    float red_w;
    int red_weighted;
    Epilogue epi;
};

// gathered feature rows: read-only path, L1-allocating (hub rows of skewed graphs are re-read by neighbouring warps)
template <int VEC> __device__ __forceinline__ void load_row_slice(float (&r)[VEC], const float *p)
{
    if constexpr (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else if constexpr (VEC == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        r[0] = v.x; r[1] = v.y;
    } else {
        r[0] = __ldg(p);
    }
}
template <int VEC> __device__ __forceinline__ void load_plain(float (&r)[VEC], const float *p)
{
    if constexpr (VEC == 4) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else if constexpr (VEC == 2) {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        r[0] = v.x; r[1] = v.y;
    } else {
        r[0] = *p;
    }
}
template <int VEC> __device__ __forceinline__ void store_slice(float *p, const float (&r)[VEC])
{
    if constexpr (VEC == 4) *reinterpret_cast<float4 *>(p) = make_float4(r[0], r[1], r[2], r[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2 *>(p) = make_float2(r[0], r[1]);
    else *p = r[0];
}
// the CSR stream is touched once per hop: keep it out of L1
__device__ __forceinline__ int32_t load_stream_i32(const int32_t *p)
{
    int32_t v;
    asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float load_stream_f32(const float *p)
{
    float v;
    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// A lane's slice of a feature row: VEC floats kept as packed fp32 pairs so that the accumulation is issued as
// FFMA2 (fma.rn.f32x2, sm_100): two IEEE fused multiply-adds per instruction, bit-identical to two fmaf.
template <int VEC> struct Slice;
template <> struct Slice<4> {
    ulonglong2 v;
    __device__ __forceinline__ void unpack(float (&f)[4]) const
    {
        f[0] = __uint_as_float((unsigned)v.x); f[1] = __uint_as_float((unsigned)(v.x >> 32));
        f[2] = __uint_as_float((unsigned)v.y); f[3] = __uint_as_float((unsigned)(v.y >> 32));
    }
    __device__ __forceinline__ void pack(const float (&f)[4])
    {
        v.x = (unsigned long long)__float_as_uint(f[0]) | ((unsigned long long)__float_as_uint(f[1]) << 32);
        v.y = (unsigned long long)__float_as_uint(f[2]) | ((unsigned long long)__float_as_uint(f[3]) << 32);
    }
    __device__ __forceinline__ void zero() { v.x = 0ULL; v.y = 0ULL; }
    __device__ __forceinline__ void load_nc(const char *p) { v = __ldg(reinterpret_cast<const ulonglong2 *>(p)); }
    __device__ __forceinline__ void load(const char *p) { v = *reinterpret_cast<const ulonglong2 *>(p); }
    __device__ __forceinline__ void load_l2(const char *p) { v = __ldcg(reinterpret_cast<const ulonglong2 *>(p)); }
    __device__ __forceinline__ void load_hint(const char *p, uint64_t pol)
    {
        asm volatile("ld.global.nc" SGL_HINT_L1 ".L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;"
                     : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol));
    }
    __device__ __forceinline__ void add(const Slice &x)
    {
        asm("add.rn.f32x2 %0, %0, %2; add.rn.f32x2 %1, %1, %3;" : "+l"(v.x), "+l"(v.y) : "l"(x.v.x), "l"(x.v.y));
    }
    __device__ __forceinline__ void store(char *p) const { *reinterpret_cast<ulonglong2 *>(p) = v; }
    // output rows are written once and not re-read by this hop: streaming store, so they do not evict X from L2
    __device__ __forceinline__ void store_streaming(char *p) const
    {
        asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
    }
    __device__ __forceinline__ void fma(float w, const Slice &x)
    {
        asm("{ .reg .b64 ww; mov.b64 ww, {%2, %2}; fma.rn.f32x2 %0, ww, %3, %0; fma.rn.f32x2 %1, ww, %4, %1; }"
            : "+l"(v.x), "+l"(v.y) : "f"(w), "l"(x.v.x), "l"(x.v.y));
    }
};
template <> struct Slice<2> {
    unsigned long long v;
    __device__ __forceinline__ void unpack(float (&f)[2]) const
    {
        f[0] = __uint_as_float((unsigned)v); f[1] = __uint_as_float((unsigned)(v >> 32));
    }
    __device__ __forceinline__ void pack(const float (&f)[2])
    {
        v = (unsigned long long)__float_as_uint(f[0]) | ((unsigned long long)__float_as_uint(f[1]) << 32);
    }
    __device__ __forceinline__ void zero() { v = 0ULL; }
    __device__ __forceinline__ void load_nc(const char *p) { v = __ldg(reinterpret_cast<const unsigned long long *>(p)); }
    __device__ __forceinline__ void load(const char *p) { v = *reinterpret_cast<const unsigned long long *>(p); }
    __device__ __forceinline__ void load_l2(const char *p) { v = __ldcg(reinterpret_cast<const unsigned long long *>(p)); }
    __device__ __forceinline__ void load_hint(const char *p, uint64_t pol)
    {
        asm volatile("ld.global.nc" SGL_HINT_L1 ".L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    }
    __device__ __forceinline__ void add(const Slice &x) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(x.v)); }
    __device__ __forceinline__ void store(char *p) const { *reinterpret_cast<unsigned long long *>(p) = v; }
    __device__ __forceinline__ void store_streaming(char *p) const
    {
        asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
    __device__ __forceinline__ void fma(float w, const Slice &x)
    {
        asm("{ .reg .b64 ww; mov.b64 ww, {%1, %1}; fma.rn.f32x2 %0, ww, %2, %0; }" : "+l"(v) : "f"(w), "l"(x.v));
    }
};
template <> struct Slice<1> {
    float v;
    __device__ __forceinline__ void unpack(float (&f)[1]) const { f[0] = v; }
    __device__ __forceinline__ void pack(const float (&f)[1]) { v = f[0]; }
    __device__ __forceinline__ void zero() { v = 0.0f; }
    __device__ __forceinline__ void load_nc(const char *p) { v = __ldg(reinterpret_cast<const float *>(p)); }
    __device__ __forceinline__ void load(const char *p) { v = *reinterpret_cast<const float *>(p); }
    __device__ __forceinline__ void load_l2(const char *p) { v = __ldcg(reinterpret_cast<const float *>(p)); }
    __device__ __forceinline__ void load_hint(const char *p, uint64_t pol)
    {
        asm volatile("ld.global.nc" SGL_HINT_L1 ".L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    }
    __device__ __forceinline__ void add(const Slice &x) { v = __fadd_rn(v, x.v); }
    __device__ __forceinline__ void store(char *p) const { *reinterpret_cast<float *>(p) = v; }
    __device__ __forceinline__ void store_streaming(char *p) const
    {
        asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    }
    __device__ __forceinline__ void fma(float w, const Slice &x) { v = fmaf(w, x.v, v); }
};


__device__ __forceinline__ float epi_nan_max(float a, float b) { return (a > b || a != a) ? a : b; }  // torch.max propagates NaN
__device__ __forceinline__ float epi_nan_min(float a, float b) { return (a < b || a != a) ? a : b; }

// The full row flush of the fused driver.  `acc` holds the finished sum of output row `row`: slice v of this lane covers
// columns cofs[v] .. cofs[v]+VEC-1 (act[v] false: beyond d).  GW lanes (mask gmask) cooperate on the row; every one of
// them calls this function together (the over-smoothing distance needs row-wide dot products).
// Arithmetic orders follow the reference's torch expressions so that sum / mean / weighted / max / min stay bit-exact:
//   sum, mean  python sum(): ((0 + f_s) + f_s+1) + ...   then ONE true division   (sum_message_op.py:9-10, mean_...:9-10)
//   weighted   acc = f_s*w_s;  acc = acc + f_k*w_k  with separately rounded products  (utils.py:91-102)
//   max / min  NaN-propagating running extremum                                       (max_message_op.py:11-12)
//   osd        c_k = <x, y_k> / (|y_k| + 1e-10) / (|x| + 1e-10), softmax over hops, weighted sum (over_smooth_distance_op.py:11-33);
//              |c_k| <= 1, so the softmax needs no max shift and folds into one pass: num += e^{c_k} y_k, den += e^{c_k}
// The per-row SCALARS of the fused row flush (degree scales, teleport coefficient), fetched when the row STARTS so that
// their latency hides behind the row's gathers: a load issued inside the flush stalls the warp once per row (arxiv-shape
// tiles hold ~128 one-entry rows).  Row-sized reads (running max / min, over-smoothing distance, teleport row) stay in the
// flush: they are the slow path; sum / mean / weighted aggregates need no read at all (red.global.add below).
template <int VEC, int VPL> struct RowPrefetch {
    float row_scale, z_scale, self_coef;
};

template <int VEC, int VPL>
__device__ __forceinline__ void prefetch_row(const SpmmParams &p, uint32_t row, const bool (&act)[VPL], const int (&cofs)[VPL],
                                             RowPrefetch<VEC, VPL> &pf)
{
    const Epilogue &e = p.epi;
    pf.row_scale = e.row_scale ? __ldg(e.row_scale + row) : 1.0f;
    pf.z_scale = (e.Z && e.z_scale) ? __ldg(e.z_scale + row) : 1.0f;
    pf.self_coef = e.self_coef ? __ldg(e.self_coef + row) : 0.0f;
    (void)act;
    (void)cofs;
}

// agg[0..VEC) += v[0..VEC) as ONE fire-and-forget reduction at the L2 (REDG.E.ADD.F32x4): no load, no stall.  Each element
// receives exactly one IEEE add per hop and the hops are ordered by kernel boundaries, so the running sum equals the
// reference's left-to-right sum bit for bit (the hardware flushes SUBNORMAL sums to zero; normal-range values are exact).
template <int VEC> __device__ __forceinline__ void red_add(float *p, const float (&v)[VEC])
{
    if constexpr (VEC == 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    else if constexpr (VEC == 2)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    else
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
}

template <int VEC, int VPL, int GW>
__device__ __forceinline__ void emit_row(const SpmmParams &p, uint32_t row, const Slice<VEC> (&acc)[VPL],
                                         const bool (&act)[VPL], const int (&cofs)[VPL], unsigned gmask,
                                         const RowPrefetch<VEC, VPL> &pf)
{
    const Epilogue &e = p.epi;
    float y[VPL][VEC];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v].unpack(y[v]);
    if (e.row_scale) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int q = 0; q < VEC; ++q) y[v][q] = __fmul_rn(y[v][q], pf.row_scale);
    }
    if (e.self_coef) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) {
                float x[VEC];
                load_plain<VEC>(x, e.self_x + (size_t)row * e.ld_self + cofs[v]);
#pragma unroll
                for (int q = 0; q < VEC; ++q) y[v][q] = fmaf(pf.self_coef, x[q], y[v][q]);
            }
    }
    if (e.acc_scale != 0.0f) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) {
                float r[VEC];
#pragma unroll
                for (int q = 0; q < VEC; ++q) r[q] = 0.0f;
                if (e.add_rows) load_plain<VEC>(r, e.add_rows + (size_t)row * e.ld_add + cofs[v]);
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    float t = __fadd_rn(__fmul_rn(e.acc_scale, y[v][q]), r[q]);   // alpha * (A^ out) + res, separately rounded
                    if (e.clamp) t = fminf(fmaxf(t, e.clamp_lo), e.clamp_hi);
                    y[v][q] = t;
                }
            }
    }
    if (p.Y) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) {
                Slice<VEC> o;
                o.pack(y[v]);
                char *yp = reinterpret_cast<char *>(p.Y + (size_t)row * p.ldy + cofs[v]);
                if (p.stream_y) o.store_streaming(yp);
                else o.store(yp);
            }
    }
    if (e.Z) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) {
                float z[VEC];
#pragma unroll
                for (int q = 0; q < VEC; ++q) z[q] = __fmul_rn(y[v][q], pf.z_scale);
                store_slice<VEC>(e.Z + (size_t)row * e.ldz + cofs[v], z);
            }
    }
    if (e.agg_op == EPI_AGG_NONE) return;
    if ((e.agg_op == EPI_AGG_SUM || e.agg_op == EPI_AGG_WEIGHTED) && !e.agg_init && e.agg_div == 0.0f) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) {
                float t[VEC];
#pragma unroll
                for (int q = 0; q < VEC; ++q) t[q] = e.agg_op == EPI_AGG_SUM ? y[v][q] : __fmul_rn(y[v][q], e.agg_w);
                red_add<VEC>(e.agg + (size_t)row * e.ld_agg + cofs[v], t);
            }
        return;
    }
    float wk = 1.0f;   // osd: e^{c_k}
    if (e.agg_op == EPI_AGG_OSD) {
        float dot = 0.0f, ny = 0.0f;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) {
                float x[VEC];
                load_plain<VEC>(x, e.x0 + (size_t)row * e.ldx0 + cofs[v]);
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    dot = fmaf(x[q], y[v][q], dot);
                    ny = fmaf(y[v][q], y[v][q], ny);
                }
            }
#pragma unroll
        for (int o = GW / 2; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(gmask, dot, o, GW);
            ny += __shfl_xor_sync(gmask, ny, o, GW);
        }
        const float c = __fdiv_rn(__fdiv_rn(dot, sqrtf(ny) + 1e-10f), __ldg(e.x0_norm + row));
        wk = expf(c);
        const float den = wk + (e.agg_init ? 0.0f : e.den[row]);
        if (e.osd_final) {
            // out = (num + e^{c} y) / den, evaluated as num/den + (e^{c}/den) y
            wk = __fdiv_rn(wk, den);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (act[v]) {
                    float a[VEC];
#pragma unroll
                    for (int q = 0; q < VEC; ++q) a[q] = 0.0f;
                    if (!e.agg_init) load_plain<VEC>(a, e.agg + (size_t)row * e.ld_agg + cofs[v]);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) a[q] = fmaf(wk, y[v][q], e.agg_init ? 0.0f : __fdiv_rn(a[q], den));
                    store_slice<VEC>(e.agg + (size_t)row * e.ld_agg + cofs[v], a);
                }
            return;
        }
        if ((threadIdx.x & (GW - 1)) == 0) e.den[row] = den;
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v)
        if (act[v]) {
            float a[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) a[q] = 0.0f;
            if (!e.agg_init) load_plain<VEC>(a, e.agg + (size_t)row * e.ld_agg + cofs[v]);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const float yv = y[v][q];
                float r;
                switch (e.agg_op) {
                case EPI_AGG_SUM: r = __fadd_rn(e.agg_init ? 0.0f : a[q], yv); break;
                case EPI_AGG_WEIGHTED: r = e.agg_init ? __fmul_rn(yv, e.agg_w) : __fadd_rn(a[q], __fmul_rn(yv, e.agg_w)); break;
                case EPI_AGG_MAX: r = e.agg_init ? yv : epi_nan_max(yv, a[q]); break;
                case EPI_AGG_MIN: r = e.agg_init ? yv : epi_nan_min(yv, a[q]); break;
                default: r = e.agg_init ? __fmul_rn(wk, yv) : fmaf(wk, yv, a[q]); break;  // osd numerator
                }
                if (e.agg_div != 0.0f) r = __fdiv_rn(r, e.agg_div);
                a[q] = r;
            }
            store_slice<VEC>(e.agg + (size_t)row * e.ld_agg + cofs[v], a);
        }
}

// the lean flush's aggregate update: red_agg[row, cofs ..] += (weighted ? acc * w : acc)
template <int VEC> __device__ __forceinline__ void red_row_slice(const SpmmParams &p, uint32_t row, int cofs, const Slice<VEC> &acc)
{
    float t[VEC];
    acc.unpack(t);
    if (p.red_weighted) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) t[q] = __fmul_rn(t[q], p.red_w);
    }
    red_add<VEC>(p.red_agg + (size_t)row * p.ld_red + cofs, t);
}

// Out-of-line entry of the fused row flush.  emit_row inlined at every row-end site of the unrolled walk (8-16 sites per
// kernel) blows the hop kernels up to ~140 KB of SASS and the instruction cache thrashes (measured: +55 % per hop on
// products-shape whatever the epilogue does); ONE copy per kernel, reached by a call, costs ~20 instructions per row.
// cofs0 = first column of this lane's first slice; slice v starts GW * VEC columns further.
template <int VEC, int VPL> struct AccPack {
    Slice<VEC> s[VPL];
};
template <int VEC, int VPL, int GW>
__device__ __noinline__ void emit_row_call(const SpmmParams *pp, uint32_t row, AccPack<VEC, VPL> acc, int cofs0, unsigned gmask,
                                           float row_scale, float z_scale, float self_coef)
{
    bool act[VPL];
    int cofs[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        cofs[v] = cofs0 + v * GW * VEC;
        act[v] = cofs[v] < pp->d;
    }
    RowPrefetch<VEC, VPL> pf;
    pf.row_scale = row_scale;
    pf.z_scale = z_scale;
    pf.self_coef = self_coef;
    emit_row<VEC, VPL, GW>(*pp, row, acc.s, act, cofs, gmask, pf);
}

}  // namespace sglb200
