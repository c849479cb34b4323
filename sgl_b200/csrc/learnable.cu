// learnable.cu -- fused forward / backward of LearnableWeightedMessageOp for the per-node kinds gate / ori_ref / jk.
//
// Reference: sgl/operators/message_op/learnable_weighted_messahe_op.py:68-101 + sgl/operators/utils.py:105-116.
// The reference materialises, per mini-batch, a [(K'*B), (K+2)*d] matrix for 'jk' (hstack of all hops repeated K'
// times), runs a Linear over it, and a [B, d, K'] stack + bmm for the weighted sum.  Here:
//   lw_scores_kernel  one warp per node: the reference-row dot product is computed ONCE and shared by the K' hops;
//                     scores[h*B + n] = <ref_n, w_ref> + <y_{start+h,n}, w_hop> + b            (hop-major, like vstack)
//   lw_combine_kernel one warp per node i: gathers its K' scores through the reference's view
//                     (gate: scores[j*B+i]; ori_ref/jk AS WRITTEN: flat[i*K'+j]), sigmoid, softmax,
//                     out_i = sum_j W_ij y_{start+j,i}
//   lw_back_node_kernel / lw_back_score_kernel  the exact adjoints (softmax, sigmoid, both dot products, the sum),
//                     parameter gradients reduced in shared memory per block, then one atomicAdd per element.
#include <math.h>

#include "common.cuh"

namespace sglb200 {

constexpr int kLwMaxFeats = 64;
constexpr int kLwWarps = 8;

struct LwPtrs {
    const float *f[kLwMaxFeats];
};
struct LwGradPtrs {
    float *g[kLwMaxFeats];
};

__device__ __forceinline__ float lw_warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float lw_sigmoid(float z) { return 1.0f / (1.0f + expf(-z)); }

// flat position (in the hop-major score vector) of weight (node i, hop j) under the reference's view
__device__ __forceinline__ int64_t lw_flat_index(int kind, int64_t i, int j, int64_t B, int kp)
{
    return kind == 2 ? (int64_t)j * B + i : i * kp + j;
}

__global__ void __launch_bounds__(kLwWarps * 32)
    lw_scores_kernel(int kind, LwPtrs feats, int n_all, int start, int kp, int64_t B, int d,
                     const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ scores)
{
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kLwWarps + (threadIdx.x >> 5);
    if (n >= B) return;
    float ref = 0.0f;
    const float *w_hop = w;
    if (kind == 3) {
        const float *x = feats.f[0] + n * d;
        for (int c = lane; c < d; c += 32) ref = fmaf(x[c], w[c], ref);
        w_hop = w + d;
    } else if (kind == 4) {
        for (int k = 0; k < n_all; ++k) {
            const float *x = feats.f[k] + n * d;
            const float *wk = w + (size_t)k * d;
            for (int c = lane; c < d; c += 32) ref = fmaf(x[c], wk[c], ref);
        }
        w_hop = w + (size_t)n_all * d;
    }
    ref = lw_warp_sum(ref);
    const float b = bias[0];
    for (int h = 0; h < kp; ++h) {
        const float *y = feats.f[start + h] + n * d;
        float s = 0.0f;
        for (int c = lane; c < d; c += 32) s = fmaf(y[c], w_hop[c], s);
        s = lw_warp_sum(s);
        if (lane == 0) scores[(int64_t)h * B + n] = s + ref + b;
    }
}

__global__ void __launch_bounds__(kLwWarps * 32)
    lw_combine_kernel(int kind, LwPtrs feats, int start, int kp, int64_t B, int d, const float *__restrict__ scores,
                      float *__restrict__ hop_w, float *__restrict__ out)
{
    __shared__ float s_w[kLwWarps][kLwMaxFeats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * kLwWarps + warp;
    if (i >= B) return;
    // softmax over hops of sigmoid(score): sigmoid outputs lie in (0,1), the max-shift is kept for fidelity with torch
    float g = -INFINITY;
    if (lane < kp) g = lw_sigmoid(scores[lw_flat_index(kind, i, lane, B, kp)]);
    float gj[2] = {g, -INFINITY};
    if (lane + 32 < kp) gj[1] = lw_sigmoid(scores[lw_flat_index(kind, i, lane + 32, B, kp)]);
    float m = fmaxf(gj[0], gj[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float e0 = lane < kp ? expf(gj[0] - m) : 0.0f;
    float e1 = lane + 32 < kp ? expf(gj[1] - m) : 0.0f;
    const float denom = lw_warp_sum(e0 + e1);
    if (lane < kp) {
        const float wv = e0 / denom;
        s_w[warp][lane] = wv;
        hop_w[i * kp + lane] = wv;
    }
    if (lane + 32 < kp) {
        const float wv = e1 / denom;
        s_w[warp][lane + 32] = wv;
        hop_w[i * kp + lane + 32] = wv;
    }
    __syncwarp();
    for (int c = lane; c < d; c += 32) {
        float acc = 0.0f;
        for (int j = 0; j < kp; ++j) acc = fmaf(s_w[warp][j], feats.f[start + j][i * d + c], acc);
        out[i * d + c] = acc;
    }
}

// per node i: dW_ij = <gout_i, y_ij>; softmax and sigmoid adjoints -> dscore at the viewed position;
// grad_y_{start+j, i} += W_ij * gout_i
__global__ void __launch_bounds__(kLwWarps * 32)
    lw_back_node_kernel(int kind, LwPtrs feats, LwGradPtrs grads, int start, int kp, int64_t B, int d,
                        const float *__restrict__ scores, const float *__restrict__ hop_w,
                        const float *__restrict__ gout, float *__restrict__ dscore)
{
    __shared__ float s_dw[kLwWarps][kLwMaxFeats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * kLwWarps + warp;
    if (i >= B) return;
    const float *go = gout + i * d;
    for (int j = 0; j < kp; ++j) {
        const float *y = feats.f[start + j] + i * d;
        float s = 0.0f;
        for (int c = lane; c < d; c += 32) s = fmaf(go[c], y[c], s);
        s = lw_warp_sum(s);
        if (lane == 0) s_dw[warp][j] = s;
    }
    __syncwarp();
    float dot = 0.0f;
    for (int j = lane; j < kp; j += 32) dot = fmaf(hop_w[i * kp + j], s_dw[warp][j], dot);
    dot = lw_warp_sum(dot);
    for (int j = lane; j < kp; j += 32) {
        const float wv = hop_w[i * kp + j];
        const float dz = wv * (s_dw[warp][j] - dot);                       // softmax adjoint
        const int64_t pos = lw_flat_index(kind, i, j, B, kp);
        const float sg = lw_sigmoid(scores[pos]);
        dscore[pos] = dz * sg * (1.0f - sg);                               // sigmoid adjoint
    }
    for (int j = 0; j < kp; ++j) {
        float *gy = grads.g[start + j];
        if (!gy) continue;
        const float wv = hop_w[i * kp + j];
        for (int c = lane; c < d; c += 32) gy[i * d + c] += wv * go[c];
    }
}

// per node n (hop-major scores): adjoints of the two dot products; parameter gradients reduced per block in shared
// memory (len_w floats) and added to global memory once per block
__global__ void __launch_bounds__(kLwWarps * 32)
    lw_back_score_kernel(int kind, LwPtrs feats, LwGradPtrs grads, int n_all, int start, int kp, int64_t B, int d,
                         const float *__restrict__ w, const float *__restrict__ dscore, float *__restrict__ grad_w,
                         float *__restrict__ grad_bias, int len_w, int private_rows)
{
    // parameter-gradient partials: one PRIVATE accumulator row per warp when they fit in shared memory (`private_rows`):
    // lane c owns columns c, c+32, ... of its warp's row, so the node loop needs no atomics at all (the round-1 form did a
    // shared-memory atomicAdd per element, node and hop, 8 warps contending on the same addresses: 1.2 TB/s); the rows
    // are summed once at the end.  Otherwise all warps share one row through atomics.
    extern __shared__ float s_gw[];  // rows x (len_w + 1)
    const int rows = private_rows ? kLwWarps : 1;
    const int row_len = len_w + 1;
    for (int c = threadIdx.x; c < rows * row_len; c += blockDim.x) s_gw[c] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ref_hops = kind == 3 ? 1 : (kind == 4 ? n_all : 0);
    const float *w_hop = w + (size_t)ref_hops * d;
    float *my = s_gw + (size_t)(private_rows ? warp : 0) * row_len;
    float *s_hop = my + (size_t)ref_hops * d;
    float dr_total = 0.0f;
    for (int64_t n = (int64_t)blockIdx.x * kLwWarps + warp; n < B; n += (int64_t)gridDim.x * kLwWarps) {
        float dr = 0.0f;
        for (int h = 0; h < kp; ++h) {
            const float ds = dscore[(int64_t)h * B + n];
            dr += ds;
            const float *y = feats.f[start + h] + n * d;
            float *gy = grads.g[start + h];
            for (int c = lane; c < d; c += 32) {
                if (gy) gy[n * d + c] += ds * w_hop[c];
                if (private_rows) s_hop[c] = fmaf(ds, y[c], s_hop[c]);
                else atomicAdd(&s_hop[c], ds * y[c]);
            }
        }
        dr_total += dr;
        for (int k = 0; k < ref_hops; ++k) {
            const float *x = feats.f[k] + n * d;
            float *gx = grads.g[k];
            const float *wk = w + (size_t)k * d;
            for (int c = lane; c < d; c += 32) {
                if (gx) gx[n * d + c] += dr * wk[c];
                if (private_rows) my[(size_t)k * d + c] = fmaf(dr, x[c], my[(size_t)k * d + c]);
                else atomicAdd(&my[(size_t)k * d + c], dr * x[c]);
            }
        }
    }
    if (lane == 0) atomicAdd(&my[len_w], dr_total);
    __syncthreads();
    for (int c = threadIdx.x; c < row_len; c += blockDim.x) {
        float t = 0.0f;
        for (int r = 0; r < rows; ++r) t += s_gw[(size_t)r * row_len + c];
        if (c < len_w) atomicAdd(&grad_w[c], t);
        else atomicAdd(grad_bias, t);
    }
}

static int lw_check(const char *who, int kind, const float *const *feats, int n_all, int start, int end, int64_t B, int d)
{
    SGL_REQUIRE(kind >= 2 && kind <= 4, "%s: kind %d is not a per-node kind (2 gate, 3 ori_ref, 4 jk)", who, kind);
    SGL_REQUIRE(feats != nullptr, "%s: feats is NULL", who);
    SGL_REQUIRE(n_all >= 1 && n_all <= kLwMaxFeats, "%s: n_all=%d outside [1,%d]", who, n_all, kLwMaxFeats);
    SGL_REQUIRE(0 <= start && start < end && end <= n_all, "%s: bad hop range [%d,%d) of %d", who, start, end, n_all);
    SGL_REQUIRE(B >= 0 && d >= 1, "%s: bad sizes", who);
    for (int k = 0; k < n_all; ++k) SGL_REQUIRE(feats[k] != nullptr, "%s: feats[%d] is NULL", who, k);
    return check_device();
}

}  // namespace sglb200

using namespace sglb200;

extern "C" {

int sglb200_lw_forward(int kind, const float *const *feats, int n_all, int start, int end, int64_t B, int d,
                       const float *w, const float *bias, float *scores, float *hop_w, float *out, void *stream_)
{
    clear_error();
    {
        const int st = lw_check("lw_forward", kind, feats, n_all, start, end, B, d);
        if (st != SGLB200_OK) return st;
    }
    SGL_REQUIRE(w && bias && scores && hop_w && out, "lw_forward: NULL argument");
    if (B == 0) return SGLB200_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    LwPtrs f;
    for (int k = 0; k < n_all; ++k) f.f[k] = feats[k];
    const int kp = end - start;
    const unsigned blocks = (unsigned)((B + kLwWarps - 1) / kLwWarps);
    lw_scores_kernel<<<blocks, kLwWarps * 32, 0, stream>>>(kind, f, n_all, start, kp, B, d, w, bias, scores);
    SGL_CUDA_CHECK(cudaGetLastError());
    lw_combine_kernel<<<blocks, kLwWarps * 32, 0, stream>>>(kind, f, start, kp, B, d, scores, hop_w, out);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

int sglb200_lw_backward(int kind, const float *const *feats, int n_all, int start, int end, int64_t B, int d,
                        const float *w, const float *bias, const float *scores, const float *hop_w,
                        const float *grad_out, float *const *grad_feats, float *grad_w, float *grad_bias,
                        float *scratch, void *stream_)
{
    clear_error();
    (void)bias;
    {
        const int st = lw_check("lw_backward", kind, feats, n_all, start, end, B, d);
        if (st != SGLB200_OK) return st;
    }
    SGL_REQUIRE(w && scores && hop_w && grad_out && grad_feats && grad_w && grad_bias && scratch,
                "lw_backward: NULL argument");
    if (B == 0) return SGLB200_OK;
    const int ref_hops = kind == 3 ? 1 : (kind == 4 ? n_all : 0);
    const int len_w = (ref_hops + 1) * d;
    SGL_REQUIRE((size_t)(len_w + 1) * sizeof(float) <= 160 * 1024, "lw_backward: parameter vector too long (%d)", len_w);
    cudaStream_t stream = (cudaStream_t)stream_;
    LwPtrs f;
    LwGradPtrs g;
    for (int k = 0; k < n_all; ++k) {
        f.f[k] = feats[k];
        g.g[k] = grad_feats[k];
    }
    const int kp = end - start;
    const unsigned blocks = (unsigned)((B + kLwWarps - 1) / kLwWarps);
    lw_back_node_kernel<<<blocks, kLwWarps * 32, 0, stream>>>(kind, f, g, start, kp, B, d, scores, hop_w, grad_out, scratch);
    SGL_CUDA_CHECK(cudaGetLastError());
    const int private_rows = (size_t)kLwWarps * (len_w + 1) * sizeof(float) <= 96 * 1024 ? 1 : 0;
    const size_t smem = (size_t)(private_rows ? kLwWarps : 1) * (len_w + 1) * sizeof(float);
    if (smem > 48 * 1024)
        SGL_CUDA_CHECK(cudaFuncSetAttribute(lw_back_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks2 = blocks < 148u * 4u ? blocks : 148u * 4u;
    lw_back_score_kernel<<<blocks2, kLwWarps * 32, smem, stream>>>(kind, f, g, n_all, start, kp, B, d, w, scratch, grad_w,
                                                                   grad_bias, len_w, private_rows);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

}  // extern "C"
