// aggregate.cu -- cross-hop message aggregation and the device-resident feature gather.
//
// Replaces the torch-CPU combiners of the reference (sgl/operators/message_op/*.py, sgl/operators/utils.py:91-116):
//   sum / mean / max / min / simple-weighted : one streaming pass, (K'+1) * N*d*4 bytes, no stacked temporaries
//     (the reference materialises a [K', N, d] stack for max/min and a [K', N*d] vstack for the weighted sum);
//   concat : strided copy into the [N, K'*d] slab (propagate can also write hops straight into such a slab);
//   over-smoothing distance (NAFS) : one warp per node, replaces the reference's python loop over N x (K+1)
//     (over_smooth_distance_op.py:27-31).
// All arithmetic orders follow the reference so that results are bit-identical where the reference's own order
// is defined (sequential left-to-right fp32 adds, separately rounded products).
#include <math.h>

#include "common.cuh"
#include "trace.cuh"

namespace sglb200 {

constexpr int kMaxFeats = 64;

struct FeatPtrs {
    const float *p[kMaxFeats];
    float w[kMaxFeats];
};
struct OutPtrs {
    float *p[kMaxFeats];
};

__device__ __forceinline__ float nan_max(float a, float b) { return (a > b || a != a) ? a : b; }  // torch.max propagates NaN
__device__ __forceinline__ float nan_min(float a, float b) { return (a < b || a != a) ? a : b; }

template <int OP, int VEC>
__global__ void __launch_bounds__(256) agg_stream_kernel(const FeatPtrs f, int n_feats, int64_t n, int d, int64_t ld_in,
                                                         float *__restrict__ out, int64_t ld_out)
{
    const int dv = d / VEC;
    const int64_t total = n * dv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / dv;
        const int col = (int)(idx - row * dv) * VEC;
        const int64_t in_off = row * ld_in + col;
        float acc[VEC], v[VEC];
        auto load = [&](int k, float (&r)[VEC]) {
            if constexpr (VEC == 4) {
                const float4 t = __ldcs(reinterpret_cast<const float4 *>(f.p[k] + in_off));
                r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
            } else {
                r[0] = __ldcs(f.p[k] + in_off);
            }
        };
        load(0, acc);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (OP == SGLB200_AGG_SUM || OP == SGLB200_AGG_MEAN) acc[e] = __fadd_rn(acc[e], 0.0f);  // python sum(): 0 + f0
            if (OP == SGLB200_AGG_WEIGHTED) acc[e] = __fmul_rn(acc[e], f.w[0]);
        }
        for (int k = 1; k < n_feats; ++k) {
            load(k, v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (OP == SGLB200_AGG_SUM || OP == SGLB200_AGG_MEAN) acc[e] = __fadd_rn(acc[e], v[e]);
                else if (OP == SGLB200_AGG_MAX) acc[e] = nan_max(v[e], acc[e]);
                else if (OP == SGLB200_AGG_MIN) acc[e] = nan_min(v[e], acc[e]);
                else if (OP == SGLB200_AGG_WEIGHTED) acc[e] = __fadd_rn(acc[e], __fmul_rn(v[e], f.w[k]));
            }
        }
        if (OP == SGLB200_AGG_MEAN) {
            const float cnt = (float)n_feats;
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = __fdiv_rn(acc[e], cnt);
        }
        float *o = out + row * ld_out + col;
        if constexpr (VEC == 4) *reinterpret_cast<float4 *>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else *o = acc[0];
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) agg_concat_kernel(const FeatPtrs f, int n_feats, int64_t n, int d, int64_t ld_in,
                                                         float *__restrict__ out, int64_t ld_out)
{
    const int dv = d / VEC;
    const int64_t per_feat = n * dv;
    const int64_t total = per_feat * n_feats;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / ((int64_t)dv * n_feats);
        const int64_t rem = idx - row * ((int64_t)dv * n_feats);
        const int k = (int)(rem / dv);
        const int col = (int)(rem - (int64_t)k * dv) * VEC;
        const float *src = f.p[k] + row * ld_in + col;
        float *dst = out + row * ld_out + (int64_t)k * d + col;
        if constexpr (VEC == 4) *reinterpret_cast<float4 *>(dst) = __ldcs(reinterpret_cast<const float4 *>(src));
        else *dst = __ldcs(src);
    }
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// NAFS over-smoothing-distance weights (over_smooth_distance_op.py:11-33); feats[0] is the reference feature X.
//   c_k = <x, y_k> / (|y_k| + 1e-10) / (|x| + 1e-10);  w = softmax_k(c);  out = sum_k w_k * y_k   (hop order)
// One warp per node.  NV float4 slices per lane cover the row (d <= 128 NV); the hop rows are fetched FOUR AT A TIME
// (4 NV independent 128-bit loads in flight per lane) and their dot products / norms are reduced together, so the
// kernel streams instead of waiting out one load-reduce chain per hop; the second pass (weighted sum) re-reads rows that
// are still in L1/L2.  Bytes: (K'+1) N d 4 from HBM.
template <int NV>
__global__ void __launch_bounds__(256) agg_osd_vec_kernel(const FeatPtrs f, int n_feats, int64_t n, int d, int64_t ld_in,
                                                          float *__restrict__ out, int64_t ld_out)
{
    __shared__ float s_c[8][kMaxFeats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= n) return;
    bool act[NV];
    float4 x[NV];
    float nx = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c = (v * 32 + lane) * 4;
        act[v] = c < d;
        x[v] = act[v] ? __ldg(reinterpret_cast<const float4 *>(f.p[0] + row * ld_in + c)) : make_float4(0, 0, 0, 0);
        nx = fmaf(x[v].x, x[v].x, fmaf(x[v].y, x[v].y, fmaf(x[v].z, x[v].z, fmaf(x[v].w, x[v].w, nx))));
    }
    nx = sqrtf(warp_sum(nx)) + 1e-10f;
    float cmax = -INFINITY;
    for (int k0 = 0; k0 < n_feats; k0 += 4) {
        float4 y[4][NV];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int v = 0; v < NV; ++v)
                y[q][v] = (k0 + q < n_feats && act[v])
                              ? __ldcs(reinterpret_cast<const float4 *>(f.p[k0 + q] + row * ld_in + (v * 32 + lane) * 4))
                              : make_float4(0, 0, 0, 0);
        float dot[4], ny[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            dot[q] = 0.0f;
            ny[q] = 0.0f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                dot[q] = fmaf(x[v].x, y[q][v].x, fmaf(x[v].y, y[q][v].y, fmaf(x[v].z, y[q][v].z, fmaf(x[v].w, y[q][v].w, dot[q]))));
                ny[q] = fmaf(y[q][v].x, y[q][v].x, fmaf(y[q][v].y, y[q][v].y, fmaf(y[q][v].z, y[q][v].z, fmaf(y[q][v].w, y[q][v].w, ny[q]))));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
                ny[q] += __shfl_xor_sync(0xffffffffu, ny[q], o);
            }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (k0 + q < n_feats) {
                const float ck = __fdiv_rn(__fdiv_rn(dot[q], sqrtf(ny[q]) + 1e-10f), nx);
                if (lane == 0) s_c[warp][k0 + q] = ck;
                cmax = fmaxf(cmax, ck);
            }
    }
    __syncwarp();
    float denom = 0.0f;
    for (int k = 0; k < n_feats; ++k) denom += expf(s_c[warp][k] - cmax);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0, 0, 0, 0);
    for (int k0 = 0; k0 < n_feats; k0 += 4) {
        float4 y[4][NV];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int v = 0; v < NV; ++v)
                y[q][v] = (k0 + q < n_feats && act[v])
                              ? __ldg(reinterpret_cast<const float4 *>(f.p[k0 + q] + row * ld_in + (v * 32 + lane) * 4))
                              : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (k0 + q < n_feats) {
                const float wk = __fdiv_rn(expf(s_c[warp][k0 + q] - cmax), denom);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc[v].x = __fadd_rn(acc[v].x, __fmul_rn(wk, y[q][v].x));
                    acc[v].y = __fadd_rn(acc[v].y, __fmul_rn(wk, y[q][v].y));
                    acc[v].z = __fadd_rn(acc[v].z, __fmul_rn(wk, y[q][v].z));
                    acc[v].w = __fadd_rn(acc[v].w, __fmul_rn(wk, y[q][v].w));
                }
            }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
        if (act[v]) *reinterpret_cast<float4 *>(out + row * ld_out + (v * 32 + lane) * 4) = acc[v];
}

// generic widths / alignments: scalar lanes
__global__ void __launch_bounds__(256) agg_osd_kernel(const FeatPtrs f, int n_feats, int64_t n, int d, int64_t ld_in,
                                                      float *__restrict__ out, int64_t ld_out)
{
    __shared__ float s_c[8][kMaxFeats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= n) return;
    const float *x = f.p[0] + row * ld_in;
    float nx = 0.0f;
    for (int c = lane; c < d; c += 32) nx = fmaf(x[c], x[c], nx);
    nx = sqrtf(warp_sum(nx)) + 1e-10f;
    float cmax = -INFINITY;
    for (int k = 0; k < n_feats; ++k) {
        const float *y = f.p[k] + row * ld_in;
        float dot = 0.0f, ny = 0.0f;
        for (int c = lane; c < d; c += 32) {
            const float yv = y[c];
            dot = fmaf(x[c], yv, dot);
            ny = fmaf(yv, yv, ny);
        }
        dot = warp_sum(dot);
        ny = sqrtf(warp_sum(ny)) + 1e-10f;
        const float ck = __fdiv_rn(__fdiv_rn(dot, ny), nx);
        if (lane == 0) s_c[warp][k] = ck;
        cmax = fmaxf(cmax, ck);
    }
    __syncwarp();
    float denom = 0.0f;
    for (int k = 0; k < n_feats; ++k) denom += expf(s_c[warp][k] - cmax);
    for (int c = lane; c < d; c += 32) {
        float acc = 0.0f;
        for (int k = 0; k < n_feats; ++k) {
            const float wk = __fdiv_rn(expf(s_c[warp][k] - cmax), denom);
            acc = __fadd_rn(acc, __fmul_rn(wk, f.p[k][row * ld_in + c]));
        }
        out[row * ld_out + c] = acc;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const FeatPtrs f, const OutPtrs o, int n_feats, int64_t ld_in,
                                                          const int64_t *__restrict__ idx, int64_t B, int d,
                                                          int64_t ld_out)
{
    const int dv = d / VEC;
    const int64_t total = B * dv;
    const int k = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / dv;
        const int col = (int)(i - b * dv) * VEC;
        const float *src = f.p[k] + idx[b] * ld_in + col;
        float *dst = o.p[k] + b * ld_out + col;
        if constexpr (VEC == 4) *reinterpret_cast<float4 *>(dst) = __ldg(reinterpret_cast<const float4 *>(src));
        else *dst = __ldg(src);
    }
}

static bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace sglb200

using namespace sglb200;

extern "C" {

int sglb200_aggregate(int op, const float *const *feats, int n_feats, int64_t n, int d, int64_t ld_in,
                      const float *weights, float *out, int64_t ld_out, void *stream_)
{
    clear_error();
    TraceRange range("sglb200_aggregate");
    SGL_REQUIRE(feats && out, "aggregate: NULL argument");
    SGL_REQUIRE(n_feats >= 1 && n_feats <= kMaxFeats, "aggregate: n_feats=%d outside [1,%d]", n_feats, kMaxFeats);
    SGL_REQUIRE(n >= 0 && d >= 0 && ld_in >= d, "aggregate: bad sizes");
    SGL_REQUIRE(op >= SGLB200_AGG_SUM && op <= SGLB200_AGG_OSD, "aggregate: unknown op %d", op);
    SGL_REQUIRE(op != SGLB200_AGG_WEIGHTED || weights, "aggregate: weighted op needs weights");
    SGL_REQUIRE(ld_out >= (op == SGLB200_AGG_CONCAT ? (int64_t)n_feats * d : (int64_t)d), "aggregate: ld_out too small");
    {
        const int st = check_device();
        if (st != SGLB200_OK) return st;
    }
    if (n == 0 || d == 0) return SGLB200_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    FeatPtrs f;
    bool vec4 = (d % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) && al16(out);
    for (int k = 0; k < n_feats; ++k) {
        SGL_REQUIRE(feats[k] != nullptr, "aggregate: feats[%d] is NULL", k);
        f.p[k] = feats[k];
        f.w[k] = weights ? weights[k] : 0.0f;
        vec4 = vec4 && al16(feats[k]);
    }
    const int threads = 256;
    const int64_t work = n * (int64_t)(vec4 ? d / 4 : d) * (op == SGLB200_AGG_CONCAT ? n_feats : 1);
    int64_t blocks64 = (work + threads - 1) / threads;
    const unsigned blocks = (unsigned)(blocks64 > 148 * 32 ? 148 * 32 : (blocks64 < 1 ? 1 : blocks64));
#define AGG_LAUNCH(OPC)                                                                                            \
    do {                                                                                                           \
        if (vec4) agg_stream_kernel<OPC, 4><<<blocks, threads, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out); \
        else agg_stream_kernel<OPC, 1><<<blocks, threads, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out);      \
    } while (0)
    switch (op) {
    case SGLB200_AGG_SUM: AGG_LAUNCH(SGLB200_AGG_SUM); break;
    case SGLB200_AGG_MEAN: AGG_LAUNCH(SGLB200_AGG_MEAN); break;
    case SGLB200_AGG_MAX: AGG_LAUNCH(SGLB200_AGG_MAX); break;
    case SGLB200_AGG_MIN: AGG_LAUNCH(SGLB200_AGG_MIN); break;
    case SGLB200_AGG_WEIGHTED: AGG_LAUNCH(SGLB200_AGG_WEIGHTED); break;
    case SGLB200_AGG_CONCAT:
        if (vec4) agg_concat_kernel<4><<<blocks, threads, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out);
        else agg_concat_kernel<1><<<blocks, threads, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out);
        break;
    case SGLB200_AGG_OSD:
        if (vec4 && d <= 128) agg_osd_vec_kernel<1><<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out);
        else if (vec4 && d <= 256) agg_osd_vec_kernel<2><<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out);
        else if (vec4 && d <= 512) agg_osd_vec_kernel<4><<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out);
        else agg_osd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(f, n_feats, n, d, ld_in, out, ld_out);
        break;
    }
#undef AGG_LAUNCH
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

int sglb200_gather_rows(const float *const *feats, int n_feats, int64_t ld_in, const int64_t *idx, int64_t B, int d,
                        float *const *outs, int64_t ld_out, void *stream_)
{
    clear_error();
    SGL_REQUIRE(feats && outs && (idx || B == 0), "gather_rows: NULL argument");
    SGL_REQUIRE(n_feats >= 1 && n_feats <= kMaxFeats, "gather_rows: n_feats=%d outside [1,%d]", n_feats, kMaxFeats);
    SGL_REQUIRE(B >= 0 && d >= 0 && ld_in >= d && ld_out >= d, "gather_rows: bad sizes");
    {
        const int st = check_device();
        if (st != SGLB200_OK) return st;
    }
    if (B == 0 || d == 0) return SGLB200_OK;
    FeatPtrs f;
    OutPtrs o;
    bool vec4 = (d % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0);
    for (int k = 0; k < n_feats; ++k) {
        SGL_REQUIRE(feats[k] && outs[k], "gather_rows: pointer %d is NULL", k);
        f.p[k] = feats[k];
        f.w[k] = 0.0f;
        o.p[k] = outs[k];
        vec4 = vec4 && al16(feats[k]) && al16(outs[k]);
    }
    const int threads = 256;
    const int64_t work = B * (int64_t)(vec4 ? d / 4 : d);
    int64_t blocks64 = (work + threads - 1) / threads;
    const dim3 grid((unsigned)(blocks64 > 148 * 16 ? 148 * 16 : blocks64), (unsigned)n_feats, 1);
    if (vec4) gather_rows_kernel<4><<<grid, threads, 0, (cudaStream_t)stream_>>>(f, o, n_feats, ld_in, idx, B, d, ld_out);
    else gather_rows_kernel<1><<<grid, threads, 0, (cudaStream_t)stream_>>>(f, o, n_feats, ld_in, idx, B, d, ld_out);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

}  // extern "C"
