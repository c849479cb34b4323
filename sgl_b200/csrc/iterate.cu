// iterate.cu -- a12: fused forward / backward of IterateLearnableWeightedMessageOp ("recursive") and the ReLU + concat
// epilogue of ProjectedConcatMessageOp.
//
// Reference (sgl/operators/message_op/iterate_learnable_weighted_message_op.py:28-51), per node, hops f_0 .. f_{K'-1}:
//     c = f_0;  S = []
//     for i in 0 .. K'-1:
//         g_i = sigmoid( Linear([f_i | c]) )                     Linear(2d -> 1): weight [w_a | w_b], bias b
//         S   = softmax( [S_0 .. S_{i-1}, g_i] )                 the running row is RE-normalised every step (:37)
//         c   = sum_{j<=i} S_j f_j
//     return c
// which costs O(K'^2) [B, d] tensor products and K' hstacks of [B, 2d] in torch.  Because  w_b . c = sum_j S_j (w_b . f_j),
// the recursion only needs the 2 K' scalars  a_j = w_a . f_j,  b_j = w_b . f_j  per node: one warp per node computes them
// with warp reductions while the K' rows stream through once, runs the scalar recursion, and forms c in a second pass
// over rows that are still in L1/L2.  The backward pass recomputes the scalar recursion from (a, b), back-propagates
// through the chained softmaxes in registers, and emits  df_j = S_j dc + da_j w_a + db_j w_b,  dw_a = sum da_j f_j,
// dw_b = sum db_j f_j (reduced in shared memory per block, one atomicAdd per element per block).
//
// ProjectedConcatMessageOp (projected_concat_message_op.py:19-28) is K' dense MLPs (cuBLAS, outside this library) whose
// outputs are ReLU'd (all but the first) and hstacked: relu_concat writes them straight into the column blocks of the
// result in one pass, relu_concat_backward masks the gradient on the way back.
#include <math.h>

#include "common.cuh"

namespace sglb200 {

constexpr int kItMaxHops = 16;
constexpr int kItWarps = 8;

struct ItPtrs {
    const float *f[kItMaxHops];
};
struct ItGradPtrs {
    float *g[kItMaxHops];
};

__device__ __forceinline__ float it_warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float it_sigmoid(float z) { return 1.0f / (1.0f + expf(-z)); }

// scalar recursion of one node.  a, b: the 2 K' dot products; S_all (optional): row i holds S^(i) (i+1 entries, stride
// kItMaxHops), g (optional): the gates.  Returns the final weights in S_out.
__device__ __forceinline__ void it_recursion(const float *a, const float *b, float bias, int kp, float *S_out, float *S_all,
                                             float *gates)
{
    float S[kItMaxHops];
    for (int i = 0; i < kp; ++i) {
        float cdot = 0.0f;
        if (i == 0) cdot = b[0];
        else
            for (int j = 0; j < i; ++j) cdot = fmaf(S[j], b[j], cdot);
        const float g = it_sigmoid(a[i] + cdot + bias);
        if (gates) gates[i] = g;
        S[i] = g;
        float mx = S[0];
        for (int j = 1; j <= i; ++j) mx = fmaxf(mx, S[j]);
        float den = 0.0f;
        for (int j = 0; j <= i; ++j) {
            S[j] = expf(S[j] - mx);
            den += S[j];
        }
        for (int j = 0; j <= i; ++j) {
            S[j] = S[j] / den;
            if (S_all) S_all[i * kItMaxHops + j] = S[j];
        }
    }
    for (int j = 0; j < kp; ++j) S_out[j] = S[j];
}

__global__ void __launch_bounds__(kItWarps * 32)
    it_forward_kernel(ItPtrs feats, int kp, int64_t B, int d, const float *__restrict__ w, const float *__restrict__ bias,
                      float *__restrict__ dots, float *__restrict__ hop_w, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kItWarps + (threadIdx.x >> 5);
    if (n >= B) return;
    float a[kItMaxHops], b[kItMaxHops];
    for (int j = 0; j < kp; ++j) {
        const float *f = feats.f[j] + n * d;
        float pa = 0.0f, pb = 0.0f;
        for (int c = lane; c < d; c += 32) {
            const float v = f[c];
            pa = fmaf(v, w[c], pa);
            pb = fmaf(v, w[d + c], pb);
        }
        a[j] = it_warp_sum(pa);
        b[j] = it_warp_sum(pb);
    }
    float S[kItMaxHops];
    it_recursion(a, b, bias[0], kp, S, nullptr, nullptr);
    if (lane < kp) {
        // every lane holds the same scalars: lane j writes entry j
        for (int j = 0; j < kp; ++j)
            if (lane == j) {
                dots[n * 2 * kp + j] = a[j];
                dots[n * 2 * kp + kp + j] = b[j];
                hop_w[n * kp + j] = S[j];
            }
    }
    for (int c = lane; c < d; c += 32) {
        float acc = feats.f[0][n * d + c] * S[0];
        for (int j = 1; j < kp; ++j) acc = acc + feats.f[j][n * d + c] * S[j];   // the reference's left-to-right sum (:42-46)
        out[n * d + c] = acc;
    }
}

__global__ void __launch_bounds__(kItWarps * 32)
    it_backward_kernel(ItPtrs feats, ItGradPtrs grads, int kp, int64_t B, int d, const float *__restrict__ w,
                       const float *__restrict__ bias, const float *__restrict__ dots, const float *__restrict__ grad_out,
                       float *__restrict__ grad_w, float *__restrict__ grad_bias)
{
    extern __shared__ float s_gw[];   // 2d + 1 partial parameter gradients of this block
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * d + 1; i += blockDim.x) s_gw[i] = 0.0f;
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * kItWarps + (threadIdx.x >> 5);
    if (n < B) {
        float a[kItMaxHops], b[kItMaxHops], dS[kItMaxHops];
        for (int j = 0; j < kp; ++j) {
            a[j] = dots[n * 2 * kp + j];
            b[j] = dots[n * 2 * kp + kp + j];
            const float *f = feats.f[j] + n * d;
            float p = 0.0f;
            for (int c = lane; c < d; c += 32) p = fmaf(grad_out[n * d + c], f[c], p);
            dS[j] = it_warp_sum(p);   // dL/dS_j of the final weights
        }
        float S_all[kItMaxHops * kItMaxHops], gates[kItMaxHops], S_fin[kItMaxHops];
        it_recursion(a, b, bias[0], kp, S_fin, S_all, gates);
        float da[kItMaxHops], db[kItMaxHops], G[kItMaxHops];
        float dbias = 0.0f;
        for (int j = 0; j < kp; ++j) {
            da[j] = 0.0f;
            db[j] = 0.0f;
            G[j] = dS[j];
        }
        for (int i = kp - 1; i >= 0; --i) {
            const float *S = S_all + i * kItMaxHops;      // S^(i): i+1 entries
            float dotgs = 0.0f;
            for (int k = 0; k <= i; ++k) dotgs = fmaf(G[k], S[k], dotgs);
            float dV[kItMaxHops];
            for (int k = 0; k <= i; ++k) dV[k] = S[k] * (G[k] - dotgs);
            const float g = gates[i];
            const float dz = dV[i] * g * (1.0f - g);
            da[i] += dz;
            dbias += dz;
            if (i == 0) {
                db[0] += dz;
            } else {
                const float *P = S_all + (i - 1) * kItMaxHops;   // S^(i-1): the weights that formed c before step i
                for (int j = 0; j < i; ++j) {
                    db[j] = fmaf(dz, P[j], db[j]);
                    G[j] = dV[j] + dz * b[j];
                }
            }
        }
        // feature gradients and the block's share of the parameter gradients
        for (int c = lane; c < d; c += 32) {
            const float go = grad_out[n * d + c];
            const float wa = w[c], wb = w[d + c];
            float gwa = 0.0f, gwb = 0.0f;
            for (int j = 0; j < kp; ++j) {
                const float fv = feats.f[j][n * d + c];
                grads.g[j][n * d + c] += S_fin[j] * go + da[j] * wa + db[j] * wb;
                gwa = fmaf(da[j], fv, gwa);
                gwb = fmaf(db[j], fv, gwb);
            }
            atomicAdd(&s_gw[c], gwa);
            atomicAdd(&s_gw[d + c], gwb);
        }
        if (lane == 0) atomicAdd(&s_gw[2 * d], dbias);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) atomicAdd(grad_w + i, s_gw[i]);
    if (threadIdx.x == 0) atomicAdd(grad_bias, s_gw[2 * d]);
}

// out[:, k*h .. (k+1)*h) = k == 0 ? y_0 : relu(y_k)          (forward);   dy_k = k == 0 ? dout_k : dout_k * [y_k > 0]
template <bool BACKWARD>
__global__ void __launch_bounds__(256) relu_concat_kernel(ItPtrs ys, ItGradPtrs dys, int kp, int64_t B, int h, float *__restrict__ out,
                                                          const float *__restrict__ grad_out)
{
    const int64_t width = (int64_t)kp * h;
    const int64_t total = B * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / width;
        const int col = (int)(i - row * width);
        const int k = col / h, c = col - k * h;
        const float y = ys.f[k][row * h + c];
        if (BACKWARD) dys.g[k][row * h + c] = (k == 0 || y > 0.0f) ? grad_out[i] : 0.0f;
        else out[i] = (k == 0) ? y : fmaxf(y, 0.0f);
    }
}

}  // namespace sglb200

using namespace sglb200;

extern "C" {

int sglb200_it_forward(const float *const *feats, int n_hops, int64_t B, int d, const float *w, const float *bias, float *dots,
                       float *hop_w, float *out, void *stream)
{
    clear_error();
    SGL_REQUIRE(feats && w && bias && dots && hop_w && out, "it_forward: NULL argument");
    SGL_REQUIRE(n_hops >= 1 && n_hops <= kItMaxHops, "it_forward: n_hops=%d outside [1,%d]", n_hops, kItMaxHops);
    SGL_REQUIRE(B >= 0 && d >= 1, "it_forward: bad sizes");
    if (B == 0) return SGLB200_OK;
    ItPtrs f;
    for (int k = 0; k < n_hops; ++k) {
        SGL_REQUIRE(feats[k] != nullptr, "it_forward: feats[%d] is NULL", k);
        f.f[k] = feats[k];
    }
    it_forward_kernel<<<(unsigned)((B + kItWarps - 1) / kItWarps), kItWarps * 32, 0, (cudaStream_t)stream>>>(f, n_hops, B, d, w, bias,
                                                                                                               dots, hop_w, out);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

int sglb200_it_backward(const float *const *feats, int n_hops, int64_t B, int d, const float *w, const float *bias,
                        const float *dots, const float *grad_out, float *const *grad_feats, float *grad_w, float *grad_bias,
                        void *stream)
{
    clear_error();
    SGL_REQUIRE(feats && w && bias && dots && grad_out && grad_feats && grad_w && grad_bias, "it_backward: NULL argument");
    SGL_REQUIRE(n_hops >= 1 && n_hops <= kItMaxHops, "it_backward: n_hops=%d outside [1,%d]", n_hops, kItMaxHops);
    SGL_REQUIRE(B >= 0 && d >= 1 && (size_t)(2 * d + 1) * sizeof(float) <= 48 * 1024, "it_backward: bad sizes");
    if (B == 0) return SGLB200_OK;
    ItPtrs f;
    ItGradPtrs g;
    for (int k = 0; k < n_hops; ++k) {
        SGL_REQUIRE(feats[k] && grad_feats[k], "it_backward: pointer %d is NULL", k);
        f.f[k] = feats[k];
        g.g[k] = grad_feats[k];
    }
    it_backward_kernel<<<(unsigned)((B + kItWarps - 1) / kItWarps), kItWarps * 32, (size_t)(2 * d + 1) * sizeof(float),
                         (cudaStream_t)stream>>>(f, g, n_hops, B, d, w, bias, dots, grad_out, grad_w, grad_bias);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

int sglb200_relu_concat(const float *const *ys, int n_hops, int64_t B, int h, float *out, void *stream)
{
    clear_error();
    SGL_REQUIRE(ys && out && n_hops >= 1 && n_hops <= kItMaxHops && B >= 0 && h >= 1, "relu_concat: bad argument");
    if (B == 0) return SGLB200_OK;
    ItPtrs f;
    ItGradPtrs g = {};
    for (int k = 0; k < n_hops; ++k) f.f[k] = ys[k];
    const int64_t total = B * n_hops * h;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    relu_concat_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(f, g, n_hops, B, h, out, nullptr);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

int sglb200_relu_concat_backward(const float *const *ys, int n_hops, int64_t B, int h, const float *grad_out, float *const *grad_ys,
                                 void *stream)
{
    clear_error();
    SGL_REQUIRE(ys && grad_out && grad_ys && n_hops >= 1 && n_hops <= kItMaxHops && B >= 0 && h >= 1, "relu_concat_backward: bad argument");
    if (B == 0) return SGLB200_OK;
    ItPtrs f;
    ItGradPtrs g;
    for (int k = 0; k < n_hops; ++k) {
        f.f[k] = ys[k];
        g.g[k] = grad_ys[k];
    }
    const int64_t total = B * n_hops * h;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    relu_concat_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(f, g, n_hops, B, h, nullptr, grad_out);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

}  // extern "C"
