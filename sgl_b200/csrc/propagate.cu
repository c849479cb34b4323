// propagate.cu -- the fused K-hop driver: propagate + degree normalisation + cross-hop aggregation in ONE pass per hop.
//
// Reference data flow being replaced (sgl/models/base_model.py:23-36):
//     feats = graph_op.propagate(adj, x)      K hops, each a full [N, d] matrix kept on the host (base_op.py:29-36)
//     out   = msg_op.aggregate(feats)         a second pass over all K+1 matrices (message_op/*.py)
// Here every hop kernel finishes a row ONCE and, while the row is still in registers (emit_row, spmm_common.cuh):
//   * scales it by deg^(r-1) -- the values of A^ are never materialised in FAST mode: the stream holds the raw weights of
//     (A+I)^T (nothing at all when they are all 1: 4 bytes per edge) and the hop input is pre-scaled by deg^(-r)
//     (utils.py:76-88:  A^ = diag(dL) (A+I)^T diag(dR));  PPR's teleport term alpha*x_i is added in the same flush;
//   * stores hop k only where the caller wants it (a model that aggregates needs none of the intermediate hops);
//   * writes the next hop's input (pre-scaled) into an internal ping-pong slab;
//   * folds the row into the running aggregate: sum / mean / weighted as ONE fire-and-forget L2 reduction per row
//     (red.global.add.v4.f32 -- no read, no stall; one IEEE add per element per hop in hop order = the reference's
//     left-to-right sum), concat / last by storing into the right place, max / min / NAFS over-smoothing-distance weights
//     by a read-modify-write in the flush (correct, but each row waits for its read: the host picks the separate
//     aggregation kernels for those where that is faster, see sgl_b200/sgap.py).
// Algorithmic bytes per hop fall from  8 nnz + 8 N d  (+ (K'+1) 4 N d for the separate aggregation pass) to
// 4..8 nnz + 8 N d + 8 N d (aggregate read-modify-write), with no aggregation pass at all.
#include <math.h>
#include <stdlib.h>

#include "spmm_common.cuh"
#include "trace.cuh"

namespace sglb200 {

// out[i, :] = in[i, :] * (scale ? scale[i] : 1)     -- the pre-scaled input of hop 1 / a plain strided copy
__global__ void __launch_bounds__(256) scale_rows_kernel(const float *__restrict__ in, int64_t ld_in, const float *__restrict__ scale,
                                                         float *__restrict__ out, int64_t ld_out, int64_t n, int d)
{
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / d;
        const int col = (int)(i - row * d);
        const float v = in[row * ld_in + col];
        out[row * ld_out + col] = scale ? __fmul_rn(v, scale[row]) : v;
    }
}

// agg[i, :] /= div  (mean: ONE true division after the left-to-right sum, mean_message_op.py:10)
__global__ void __launch_bounds__(256) divide_rows_kernel(float *__restrict__ agg, int64_t ld, int64_t n, int d, float div)
{
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / d;
        const int col = (int)(i - row * d);
        agg[row * ld + col] = __fdiv_rn(agg[row * ld + col], div);
    }
}

// hop 0 of the running aggregate (the input features themselves), one warp per row:
//   sum/mean: 0 + x   weighted: x * w0   max/min: x   osd: |x|, c0 = <x,x>/(|x|+eps)/(|x|+eps), num = e^{c0} x, den = e^{c0}
__global__ void __launch_bounds__(256) agg_init_kernel(const float *__restrict__ x, int64_t ldx, int64_t n, int d, int op, float w0,
                                                       float div, float *__restrict__ agg, int64_t ld_agg,
                                                       float *__restrict__ x_norm, float *__restrict__ den, int osd_final)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const float *xr = x + row * ldx;
    float *ar = agg + row * ld_agg;
    float e0 = 1.0f;
    if (op == EPI_AGG_OSD) {
        float nx = 0.0f;
        for (int c = lane; c < d; c += 32) nx = fmaf(xr[c], xr[c], nx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nx += __shfl_xor_sync(0xffffffffu, nx, o);
        const float norm = sqrtf(nx) + 1e-10f;
        e0 = expf(__fdiv_rn(__fdiv_rn(nx, norm), norm));
        if (lane == 0) {
            x_norm[row] = norm;
            den[row] = e0;
        }
        if (osd_final) e0 = 1.0f;   // K = 0: the only hop has weight 1
    }
    for (int c = lane; c < d; c += 32) {
        const float v = xr[c];
        float r;
        switch (op) {
        case EPI_AGG_SUM: r = __fadd_rn(0.0f, v); break;
        case EPI_AGG_WEIGHTED: r = __fmul_rn(v, w0); break;
        case EPI_AGG_OSD: r = __fmul_rn(e0, v); break;
        default: r = v; break;
        }
        if (div != 0.0f) r = __fdiv_rn(r, div);
        ar[c] = r;
    }
}

static int ensure_ping(sglb200_graph *g, size_t floats, int need_aux, cudaStream_t stream)
{
    if (floats > g->ping_floats) {
        // stream-ordered growth: earlier passes enqueued on this stream finish with the old slabs first
        for (int k = 0; k < 2; ++k) {
            if (g->ping[k]) SGL_CUDA_CHECK(cudaFreeAsync(g->ping[k], stream));
            g->ping[k] = nullptr;
        }
        g->bytes_resident -= 2 * g->ping_floats * sizeof(float);
        g->ping_floats = 0;
        for (int k = 0; k < 2; ++k) SGL_CUDA_CHECK(cudaMallocAsync(&g->ping[k], floats * sizeof(float), stream));
        g->ping_floats = floats;
        g->bytes_resident += 2 * floats * sizeof(float);
    }
    if (need_aux && !g->aux) {
        SGL_CUDA_CHECK(cudaMalloc(&g->aux, sizeof(float) * 2 * (size_t)(g->n_rows > 0 ? g->n_rows : 1)));
        g->bytes_resident += sizeof(float) * 2 * (size_t)g->n_rows;
    }
    return SGLB200_OK;
}

static int copy_rows(const float *in, int64_t ld_in, const float *scale, float *out, int64_t ld_out, int64_t n, int d,
                     cudaStream_t stream)
{
    const int64_t total = n * d;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    scale_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, ld_in, scale, out, ld_out, n, d);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

}  // namespace sglb200

using namespace sglb200;

extern "C" {

int sglb200_spmm_axpby(sglb200_graph_t g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode, float alpha,
                       const float *res, int64_t ld_res, int do_clamp, float lo, float hi, void *stream)
{
    clear_error();
    SGL_REQUIRE(g && X && Y, "spmm_axpby: NULL argument");
    SGL_REQUIRE(alpha != 0.0f, "spmm_axpby: alpha must be non-zero");
    SGL_REQUIRE(d >= 0 && d <= 512, "spmm_axpby: feature widths up to 512");
    SGL_REQUIRE(!res || ld_res >= d, "spmm_axpby: ld_res too small");
    Epilogue e = {};
    e.acc_scale = alpha;
    e.add_rows = res;
    e.ld_add = ld_res;
    e.clamp = do_clamp;
    e.clamp_lo = lo;
    e.clamp_hi = hi;
    return spmm_launch_ex(g, X, ldx, Y, ldy, d, mode, 0, 0, -1, &e, 0, (cudaStream_t)stream);
}

int sglb200_propagate_fused(sglb200_graph_t g, const float *X, int64_t ldx, float *const *hops_out, int64_t ld_hops, int d, int K,
                            int mode, int agg_op, int agg_start, int agg_end, const float *agg_weights, float *agg_out,
                            int64_t ld_out, int fuse_norm, void *stream_)
{
    clear_error();
    TraceRange range("sglb200_propagate_fused");
    SGL_REQUIRE(g && X, "propagate_fused: NULL argument");
    SGL_REQUIRE(K >= 0 && d >= 0, "propagate_fused: negative size");
    SGL_REQUIRE(g->n_rows == g->n_cols, "propagate_fused: operator must be square");
    SGL_REQUIRE(mode == SGLB200_MODE_FAST || mode == SGLB200_MODE_EXACT, "propagate_fused: unknown mode %d", mode);
    SGL_REQUIRE(agg_op >= -1 && agg_op <= SGLB200_AGG_LAST, "propagate_fused: unknown aggregation %d", agg_op);
    SGL_REQUIRE(agg_op < 0 || agg_out != nullptr, "propagate_fused: aggregation needs an output buffer");
    SGL_REQUIRE(d <= 512, "propagate_fused: feature widths up to 512 (wider: sglb200_propagate + sglb200_aggregate)");
    const int64_t n = g->n_rows;
    if (n == 0 || d == 0) return SGLB200_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool ranged = agg_op == SGLB200_AGG_SUM || agg_op == SGLB200_AGG_MEAN || agg_op == SGLB200_AGG_MAX ||
                        agg_op == SGLB200_AGG_MIN || agg_op == SGLB200_AGG_WEIGHTED || agg_op == SGLB200_AGG_CONCAT;
    if (ranged) {
        SGL_REQUIRE(0 <= agg_start && agg_start < agg_end && agg_end <= K + 1, "propagate_fused: hop range [%d, %d) outside [0, %d]",
                    agg_start, agg_end, K + 1);
        SGL_REQUIRE(agg_op != SGLB200_AGG_WEIGHTED || agg_weights, "propagate_fused: weighted aggregation needs weights");
        SGL_REQUIRE(ld_out >= (agg_op == SGLB200_AGG_CONCAT ? (int64_t)(agg_end - agg_start) * d : (int64_t)d),
                    "propagate_fused: ld_out too small");
    } else if (agg_op >= 0) {
        SGL_REQUIRE(ld_out >= d, "propagate_fused: ld_out too small");
    }
    auto user_hop = [&](int k) -> float * { return hops_out ? hops_out[k] : nullptr; };
    SGL_REQUIRE(!hops_out || ld_hops >= d, "propagate_fused: hop row stride too small");
    // FAST mode with scaling vectors in the handle: the normalisation is fused (values never materialised)
    const bool fused = fuse_norm && mode == SGLB200_MODE_FAST && g->has_scaling;
    // an internal slab is needed wherever a hop's input is not something the caller keeps
    auto kept = [&](int k) {
        return user_hop(k) != nullptr || (agg_op == SGLB200_AGG_CONCAT && k >= agg_start && k < agg_end) ||
               (agg_op == SGLB200_AGG_LAST && k == K);
    };
    bool need_ping = fused && K > 0;
    for (int k = 1; k <= K; ++k)
        if (!kept(k)) need_ping = true;
    {
        const int st = ensure_ping(g, need_ping ? (size_t)n * (size_t)d : 0, agg_op == SGLB200_AGG_OSD, stream);
        if (st != SGLB200_OK) return st;
    }
    int epi_op = EPI_AGG_NONE;
    switch (agg_op) {
    case SGLB200_AGG_SUM: case SGLB200_AGG_MEAN: epi_op = EPI_AGG_SUM; break;
    case SGLB200_AGG_WEIGHTED: epi_op = EPI_AGG_WEIGHTED; break;
    case SGLB200_AGG_MAX: epi_op = EPI_AGG_MAX; break;
    case SGLB200_AGG_MIN: epi_op = EPI_AGG_MIN; break;
    case SGLB200_AGG_OSD: epi_op = EPI_AGG_OSD; break;
    default: break;
    }
    const int first = epi_op == EPI_AGG_OSD ? 0 : agg_start, last = epi_op == EPI_AGG_OSD ? K : agg_end - 1;
    const float mean_div = agg_op == SGLB200_AGG_MEAN ? (float)(agg_end - agg_start) : 0.0f;

    // ---- hop 0: the caller's copy, the concat block, the aggregate's first term, the pre-scaled input of hop 1 ----------
    if (user_hop(0) && user_hop(0) != X) {
        const int st = copy_rows(X, ldx, nullptr, user_hop(0), ld_hops, n, d, stream);
        if (st != SGLB200_OK) return st;
    }
    if (agg_op == SGLB200_AGG_CONCAT && agg_start == 0) {
        const int st = copy_rows(X, ldx, nullptr, agg_out, ld_out, n, d, stream);
        if (st != SGLB200_OK) return st;
    }
    if (agg_op == SGLB200_AGG_LAST && K == 0) {
        const int st = copy_rows(X, ldx, nullptr, agg_out, ld_out, n, d, stream);
        if (st != SGLB200_OK) return st;
    }
    if (epi_op != EPI_AGG_NONE && first == 0) {
        const float w0 = agg_weights ? agg_weights[0] : 1.0f;
        agg_init_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(
            X, ldx, n, d, epi_op, w0, last == 0 ? mean_div : 0.0f, agg_out, ld_out, g->aux, g->aux ? g->aux + n : nullptr,
            epi_op == EPI_AGG_OSD && K == 0);
        SGL_CUDA_CHECK(cudaGetLastError());
    }
    const bool running_sum = epi_op == EPI_AGG_SUM || epi_op == EPI_AGG_WEIGHTED;
    if (running_sum && first >= 1) {
        // no hop-0 term: start from zeros so that every included hop is a pure add (0 + f_s = f_s exactly, like python's sum())
        SGL_CUDA_CHECK(cudaMemset2DAsync(agg_out, (size_t)ld_out * sizeof(float), 0, (size_t)d * sizeof(float), (size_t)n, stream));
    }
    const float *in = X;
    int64_t ld_in = ldx;
    if (fused && K > 0) {
        const int st = copy_rows(X, ldx, g->col_scale, g->ping[0], d, n, d, stream);   // Z_0 = dR (.) X
        if (st != SGLB200_OK) return st;
        in = g->ping[0];
        ld_in = d;
    }
    // ---- hops 1..K ------------------------------------------------------------------------------------------------------
    HopTimer timer("propagate_fused");
    timer.mark(stream);
    for (int k = 1; k <= K; ++k) {
        Epilogue e = {};
        bool need_epi = false;
        float *Y = user_hop(k);
        int64_t ldy = ld_hops;
        if (agg_op == SGLB200_AGG_CONCAT && k >= agg_start && k < agg_end) {
            SGL_REQUIRE(Y == nullptr, "propagate_fused: concat writes the hops into the concat slab; pass no separate hop buffers");
            Y = agg_out + (int64_t)(k - agg_start) * d;
            ldy = ld_out;
        }
        if (agg_op == SGLB200_AGG_LAST && k == K) {
            SGL_REQUIRE(Y == nullptr || Y == agg_out, "propagate_fused: LAST writes hop K into agg_out");
            Y = agg_out;
            ldy = ld_out;
        }
        float *next_in = nullptr;   // where hop k+1 reads from
        int64_t ld_next = d;
        if (fused) {
            e.row_scale = g->row_scale;
            if (g->self_coef) {
                e.self_coef = g->self_coef;
                e.self_x = in;
                e.ld_self = ld_in;
            }
            need_epi = true;
            if (k < K) {
                e.Z = g->ping[k & 1];
                e.ldz = d;
                e.z_scale = g->col_scale;
                next_in = e.Z;
            }
        } else if (k < K) {
            if (Y) {
                next_in = Y;
                ld_next = ldy;
            } else {
                Y = g->ping[k & 1];       // nobody keeps this hop: it only feeds the next one
                ldy = d;
                next_in = Y;
            }
        }
        if (epi_op != EPI_AGG_NONE && k >= first && k <= last) {
            e.agg_op = epi_op;
            e.agg_init = (k == first && !running_sum) ? 1 : 0;   // first > 0 here (hop 0: agg_init_kernel); running sums start from zeros
            e.agg_w = agg_weights ? agg_weights[k] : 1.0f;
            e.agg_div = 0.0f;   // mean: the running sum uses L2 reductions (no read); the division is one light pass at the end
            e.agg = agg_out;
            e.ld_agg = ld_out;
            if (epi_op == EPI_AGG_OSD) {
                e.x0 = X;
                e.ldx0 = ldx;
                e.x0_norm = g->aux;
                e.den = g->aux + n;
                e.osd_final = (k == K) ? 1 : 0;
            }
            need_epi = true;
        }
        if (Y == nullptr && !need_epi) {
            // nothing consumes this hop (possible only for k == K without outputs): still compute into the slab
            Y = g->ping[k & 1];
            ldy = d;
        }
        const int st = spmm_launch_ex(g, in, ld_in, Y, ldy, d, mode, 0, 0, -1, need_epi ? &e : nullptr, fused ? 1 : 0, stream);
        if (st != SGLB200_OK) return st;
        timer.mark(stream);
        in = next_in;
        ld_in = ld_next;
    }
    timer.report();
    if (mean_div != 0.0f && last >= 1) {
        const int64_t total = n * d;
        int64_t blocks = (total + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        divide_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(agg_out, ld_out, n, d, mean_div);
        SGL_CUDA_CHECK(cudaGetLastError());
    }
    return SGLB200_OK;
}

}  // extern "C"
