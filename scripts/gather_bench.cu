// gather_bench.cu -- microbenchmark: how fast can a B200 gather random 512-byte rows (d = 128 floats)?
// Compares three staging mechanisms for the SpMM inner loop on the same index stream:
//   ldg     : LDG.128 into registers, U rows in flight per warp
//   ldgsts  : cp.async (LDGSTS.128) into a per-warp shared-memory ring, consumed with LDS.128
//   bulk    : cp.async.bulk (TMA 1-D bulk copy, UBLKCP), one 512 B row per LANE per instruction, mbarrier completion
//   gather4 : cp.async.bulk.tensor.2d.tile::gather4 (UTMALDG, sm_100): FOUR rows of a 2-D tensor map per instruction
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_bench scripts/gather_bench.cu
// Run  :  ./gather_bench [n_rows] [n_idx] [skew 0|1]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int D = 128;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------------------
template <int U, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gather_ldg(const float *__restrict__ X, const int *__restrict__ idx,
                                                         int64_t per_warp, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int *my = idx + w * per_warp;
    const char *xb = reinterpret_cast<const char *>(X) + lane * 16;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 1
    for (int64_t b = 0; b < per_warp; b += 32) {
        const int c = my[b + lane];
#pragma unroll 1
        for (int k = 0; k < 32; k += U) {
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned cc = __shfl_sync(FULL, c, k + u);
                x[u] = __ldg(reinterpret_cast<const float4 *>(xb + (uint64_t)cc * (D * 4)));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w;
            }
        }
    }
    reinterpret_cast<float4 *>(out + w * D)[lane] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// per-warp ring of STAGES x ROWS rows in shared memory, filled with LDGSTS (one warp instruction = one 512 B row)
template <int STAGES, int ROWS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gather_ldgsts(const float *__restrict__ X, const int *__restrict__ idx,
                                                            int64_t per_warp, float *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * WARPS + wib;
    const int *my = idx + w * per_warp;
    float *ring = reinterpret_cast<float *>(smem) + (size_t)wib * STAGES * ROWS * D;
    const char *xb = reinterpret_cast<const char *>(X) + lane * 16;
    float4 acc = make_float4(0, 0, 0, 0);
    const int64_t n_groups = per_warp / ROWS;
    auto issue = [&](int64_t g) {
        const int stage = (int)(g % STAGES);
        const int c = (lane < ROWS) ? my[g * ROWS + lane] : 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const unsigned cc = __shfl_sync(FULL, c, r);
            const uint32_t dst = smem_u32(ring + ((size_t)stage * ROWS + r) * D) + lane * 16;
            const char *src = xb + (uint64_t)cc * (D * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    };
    for (int s = 0; s < STAGES - 1 && s < n_groups; ++s) issue(s);
#pragma unroll 1
    for (int64_t g = 0; g < n_groups; ++g) {
        if (g + STAGES - 1 < n_groups) issue(g + STAGES - 1);
        else asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1));
        __syncwarp();
        const float4 *rows = reinterpret_cast<const float4 *>(ring + (size_t)(g % STAGES) * ROWS * D);
#pragma unroll 4
        for (int r = 0; r < ROWS; ++r) {
            const float4 v = rows[r * (D / 4) + lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __syncwarp();
    }
    reinterpret_cast<float4 *>(out + w * D)[lane] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// per-warp ring, filled by TMA bulk copies: every lane copies ONE whole 512 B row per instruction (32 rows / stage)
template <int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gather_bulk(const float *__restrict__ X, const int *__restrict__ idx,
                                                          int64_t per_warp, float *__restrict__ out)
{
    constexpr int ROWS = 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * WARPS + wib;
    const int *my = idx + w * per_warp;
    float *ring = reinterpret_cast<float *>(smem) + (size_t)wib * STAGES * ROWS * D;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * STAGES * ROWS * D * 4) + wib * STAGES;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    float4 acc = make_float4(0, 0, 0, 0);
    const int64_t n_groups = per_warp / ROWS;
    auto issue = [&](int64_t g) {
        const int stage = (int)(g % STAGES);
        const uint32_t bar = smem_u32(bars + stage);
        const unsigned c = (unsigned)my[g * ROWS + lane];
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWS * D * 4));
        __syncwarp();
        const uint32_t dst = smem_u32(ring + ((size_t)stage * ROWS + lane) * D);
        const char *src = reinterpret_cast<const char *>(X) + (uint64_t)c * (D * 4);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"(D * 4), "r"(bar)
                     : "memory");
    };
    for (int s = 0; s < STAGES - 1 && s < n_groups; ++s) issue(s);
#pragma unroll 1
    for (int64_t g = 0; g < n_groups; ++g) {
        if (g + STAGES - 1 < n_groups) issue(g + STAGES - 1);
        const int stage = (int)(g % STAGES);
        const uint32_t bar = smem_u32(bars + stage);
        const uint32_t parity = (uint32_t)((g / STAGES) & 1);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done)
                         : "r"(bar), "r"(parity)
                         : "memory");
        }
        const float4 *rows = reinterpret_cast<const float4 *>(ring + (size_t)stage * ROWS * D);
#pragma unroll 4
        for (int r = 0; r < ROWS; ++r) {
            const float4 v = rows[r * (D / 4) + lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __syncwarp();
    }
    reinterpret_cast<float4 *>(out + w * D)[lane] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// per-warp ring filled by TMA gather4: lanes 0..ISSUERS-1 each fetch FOUR 512 B rows per instruction through a 2-D tensor
// map of X (box = {D, 1}); ROWS = 4 * ISSUERS rows per stage
template <int STAGES, int ISSUERS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gather_g4(const __grid_constant__ CUtensorMap tmap, const int *__restrict__ idx,
                                                        int64_t per_warp, float *__restrict__ out)
{
    constexpr int ROWS = 4 * ISSUERS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * WARPS + wib;
    const int *my = idx + w * per_warp;
    float *ring = reinterpret_cast<float *>(smem) + (size_t)wib * STAGES * ROWS * D;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * STAGES * ROWS * D * 4) + wib * STAGES;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    float4 acc = make_float4(0, 0, 0, 0);
    const int64_t n_groups = per_warp / ROWS;
    auto issue = [&](int64_t g) {
        const int stage = (int)(g % STAGES);
        const uint32_t bar = smem_u32(bars + stage);
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWS * D * 4));
        __syncwarp();
        if (lane < ISSUERS) {
            const int4 c = *reinterpret_cast<const int4 *>(my + g * ROWS + lane * 4);
            const uint32_t dst = smem_u32(ring + ((size_t)stage * ROWS + lane * 4) * D);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                         " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(&tmap), "r"(bar), "r"(0), "r"(c.x),
                         "r"(c.y), "r"(c.z), "r"(c.w)
                         : "memory");
        }
    };
    for (int s = 0; s < STAGES - 1 && s < n_groups; ++s) issue(s);
#pragma unroll 1
    for (int64_t g = 0; g < n_groups; ++g) {
        if (g + STAGES - 1 < n_groups) issue(g + STAGES - 1);
        const int stage = (int)(g % STAGES);
        const uint32_t bar = smem_u32(bars + stage);
        const uint32_t parity = (uint32_t)((g / STAGES) & 1);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done)
                         : "r"(bar), "r"(parity)
                         : "memory");
        }
        const float4 *rows = reinterpret_cast<const float4 *>(ring + (size_t)stage * ROWS * D);
#pragma unroll 4
        for (int r = 0; r < ROWS; ++r) {
            const float4 v = rows[r * (D / 4) + lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __syncwarp();
    }
    reinterpret_cast<float4 *>(out + w * D)[lane] = acc;
}

__global__ void stream_read(const float4 *__restrict__ X, int64_t n4, float *out)
{
    float4 acc = make_float4(0, 0, 0, 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = X[i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x == 12345.678f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

template <typename F> static float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char **argv)
{
    const int64_t n_rows = argc > 1 ? atoll(argv[1]) : 169343;
    int64_t n_idx = argc > 2 ? atoll(argv[2]) : (1 << 24);
    const int skew = argc > 3 ? atoi(argv[3]) : 0;
    float *X, *out;
    int *idx;
    CK(cudaMalloc(&X, n_rows * D * 4));
    CK(cudaMemset(X, 0, n_rows * D * 4));
    std::vector<int> h(n_idx);
    std::mt19937_64 rng(1);
    for (int64_t i = 0; i < n_idx; ++i) {
        if (skew) {  // R-MAT style marginal: each of 18 bits is 1 with probability 0.24, hashed
            uint64_t v = 0;
            uint64_t r = rng();
            for (int b = 0; b < 18; ++b) { v = (v << 1) | ((r & 0xff) < 61); r >>= 3; if (b % 16 == 15) r = rng(); }
            h[i] = (int)(((v * 0x9E3779B1ull) & 0x7fffffff) % n_rows);
        } else {
            h[i] = (int)(rng() % n_rows);
        }
    }
    CK(cudaMalloc(&idx, n_idx * 4));
    CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, (size_t)(1 << 20) * D * 4));
    const double bytes = (double)n_idx * D * 4;
    printf("table %.1f MB, %lld gathers of 512 B (%.2f GB), skew=%d\n", n_rows * D * 4 / 1e6, (long long)n_idx, bytes / 1e9, skew);
    {
        float ms = time_ms([&] { stream_read<<<148 * 16, 256>>>(reinterpret_cast<float4 *>(X), n_rows * D / 4, out); });
        printf("%-34s %8.3f ms  %8.1f GB/s (table streamed once)\n", "stream_read", ms, n_rows * D * 4 / ms / 1e6);
    }
#define RUN_LDG(U, WARPS, PW)                                                                         \
    {                                                                                                 \
        const int64_t warps = n_idx / (PW);                                                           \
        float ms = time_ms([&] { gather_ldg<U, WARPS><<<(unsigned)(warps / WARPS), WARPS * 32>>>(X, idx, PW, out); }); \
        printf("ldg    U=%-2d warps/cta=%d per_warp=%-5d %8.3f ms  %8.1f GB/s\n", U, WARPS, PW, ms, bytes / ms / 1e6); \
    }
    RUN_LDG(4, 8, 512)
    RUN_LDG(8, 8, 512)
    RUN_LDG(16, 8, 512)
    RUN_LDG(32, 8, 512)
    RUN_LDG(8, 4, 512)
    RUN_LDG(8, 8, 128)
    RUN_LDG(16, 8, 128)
    RUN_LDG(8, 8, 2048)
#define RUN_STS(STAGES, ROWS, WARPS, PW)                                                              \
    {                                                                                                 \
        const int64_t warps = n_idx / (PW);                                                           \
        const size_t sm = (size_t)WARPS * STAGES * ROWS * D * 4;                                      \
        CK(cudaFuncSetAttribute(gather_ldgsts<STAGES, ROWS, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        float ms = time_ms([&] { gather_ldgsts<STAGES, ROWS, WARPS><<<(unsigned)(warps / WARPS), WARPS * 32, sm>>>(X, idx, PW, out); }); \
        printf("ldgsts stages=%d rows=%-2d warps/cta=%d smem=%3zuKB per_warp=%-5d %8.3f ms  %8.1f GB/s\n", STAGES, ROWS, WARPS, sm >> 10, PW, ms, bytes / ms / 1e6); \
    }
    RUN_STS(2, 8, 8, 512)
    RUN_STS(3, 8, 8, 512)
    RUN_STS(4, 8, 8, 512)
    RUN_STS(2, 16, 8, 512)
    RUN_STS(3, 16, 4, 512)
    RUN_STS(4, 16, 4, 512)
    RUN_STS(2, 32, 4, 512)
    RUN_STS(4, 8, 8, 128)
#define RUN_BULK(STAGES, WARPS, PW)                                                                   \
    {                                                                                                 \
        const int64_t warps = n_idx / (PW);                                                           \
        const size_t sm = (size_t)WARPS * STAGES * 32 * D * 4 + WARPS * STAGES * 8;                   \
        CK(cudaFuncSetAttribute(gather_bulk<STAGES, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        float ms = time_ms([&] { gather_bulk<STAGES, WARPS><<<(unsigned)(warps / WARPS), WARPS * 32, sm>>>(X, idx, PW, out); }); \
        printf("bulk   stages=%d warps/cta=%d smem=%3zuKB per_warp=%-5d %8.3f ms  %8.1f GB/s\n", STAGES, WARPS, sm >> 10, PW, ms, bytes / ms / 1e6); \
    }
    RUN_BULK(2, 4, 512)
    RUN_BULK(2, 2, 512)
    RUN_BULK(3, 2, 512)
    RUN_BULK(2, 1, 512)
    RUN_BULK(3, 4, 512)
    RUN_BULK(2, 4, 128)
    RUN_BULK(2, 4, 2048)
    {
        // 2-D tensor map of X: inner dim D floats, outer dim n_rows, box {D, 1} (gather4 fetches 4 such boxes)
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        CUtensorMap tmap;
        cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)n_rows};
        cuuint64_t strides[1] = {(cuuint64_t)D * 4};
        cuuint32_t box[2] = {(cuuint32_t)D, 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
#define RUN_G4(STAGES, ISSUERS, WARPS, PW)                                                            \
    {                                                                                                 \
        const int64_t warps = n_idx / (PW);                                                           \
        const size_t sm = (size_t)WARPS * STAGES * 4 * ISSUERS * D * 4 + WARPS * STAGES * 8;          \
        CK(cudaFuncSetAttribute(gather_g4<STAGES, ISSUERS, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        float ms = time_ms([&] { gather_g4<STAGES, ISSUERS, WARPS><<<(unsigned)(warps / WARPS), WARPS * 32, sm>>>(tmap, idx, PW, out); }); \
        printf("gather4 stages=%d rows/stage=%-2d warps/cta=%d smem=%3zuKB per_warp=%-5d %8.3f ms  %8.1f GB/s\n", STAGES, 4 * ISSUERS, WARPS, sm >> 10, PW, ms, bytes / ms / 1e6); \
    }
        RUN_G4(2, 8, 4, 512)
        RUN_G4(2, 8, 2, 512)
        RUN_G4(3, 8, 2, 512)
        RUN_G4(2, 4, 4, 512)
        RUN_G4(3, 4, 4, 512)
        RUN_G4(4, 2, 4, 512)
        RUN_G4(4, 2, 8, 512)
        RUN_G4(2, 2, 8, 512)
        RUN_G4(3, 1, 8, 512)
        RUN_G4(6, 1, 8, 512)
    }
    return 0;
}
