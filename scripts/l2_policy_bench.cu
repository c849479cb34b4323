// l2_policy_bench.cu -- microbenchmark: can a B200 keep a HOT set of feature rows resident in L2 while a cold stream of
// random row gathers passes through it?  This is the access pattern of the hop kernel on an HBM-resident skewed graph:
// a fraction f of the gathers hits H hub rows, the rest is uniform over a table many times larger than L2.
//   policy 0: plain LDG.128 (hardware replacement only)
//   policy 1: per-load createpolicy hints -- hub rows L2::evict_last, cold rows L2::evict_first
//   policy 2: policy 1 with cudaLimitPersistingL2CacheSize raised to the device maximum
//   policy 3: per-launch access policy window over the hub prefix (persisting), misses streaming
//   policy 4: hub rows evict_last only (cold rows unhinted), persisting limit raised
//   policy 5: cold rows evict_first only (hub rows unhinted)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/bin/l2_policy_bench scripts/l2_policy_bench.cu
// Run  : ./l2_policy_bench [row_floats] [n_rows] [hub_rows] [hub_fraction] [n_idx]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr unsigned FULL = 0xffffffffu;

template <int MODE>  // 0 none, 1 both hints, 4 hub only, 5 cold only
__global__ void __launch_bounds__(256) gather(const float *__restrict__ X, int row_bytes, const int *__restrict__ idx,
                                              int64_t per_warp, unsigned hub_rows, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int *my = idx + w * per_warp;
    const bool act = lane * 16 < row_bytes;
    const char *xb = reinterpret_cast<const char *>(X) + (act ? lane * 16 : 0);
    uint64_t pol_hub, pol_cold, pol_norm;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_hub));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_cold));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_norm));
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 1
    for (int64_t b = 0; b < per_warp; b += 32) {
        const int c = my[b + lane];
#pragma unroll 1
        for (int k = 0; k < 32; k += 8) {
            float4 x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned cc = __shfl_sync(FULL, c, k + u);
                const char *p = xb + (uint64_t)cc * (unsigned)row_bytes;
                if (MODE == 0) {
                    x[u] = __ldg(reinterpret_cast<const float4 *>(p));
                } else {
                    uint64_t pol = cc < hub_rows ? (MODE == 5 ? pol_norm : pol_hub) : (MODE == 4 ? pol_norm : pol_cold);
                    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                                 : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "l"(p), "l"(pol));
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
        }
    }
    if (act) reinterpret_cast<float4 *>(out + w * 128)[lane] = acc;
}

int main(int argc, char **argv)
{
    const int row_floats = argc > 1 ? atoi(argv[1]) : 100;
    const int64_t n_rows = argc > 2 ? atoll(argv[2]) : 2449029;
    const int64_t hub_rows = argc > 3 ? atoll(argv[3]) : 131072;
    const double f = argc > 4 ? atof(argv[4]) : 0.74;
    const int64_t n_idx = argc > 5 ? atoll(argv[5]) : (1 << 26);
    const int row_bytes = row_floats * 4;
    float *X, *out;
    int *idx;
    CK(cudaMalloc(&X, n_rows * row_bytes));
    CK(cudaMemset(X, 0, n_rows * row_bytes));
    std::vector<int> h(n_idx);
    std::mt19937_64 rng(1);
    for (int64_t i = 0; i < n_idx; ++i) {
        const double u = (rng() >> 11) * (1.0 / 9007199254740992.0);
        h[i] = u < f ? (int)(rng() % hub_rows) : (int)(hub_rows + rng() % (n_rows - hub_rows));
    }
    CK(cudaMalloc(&idx, n_idx * 4));
    CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
    const int64_t per_warp = 1024, warps = n_idx / per_warp;
    CK(cudaMalloc(&out, warps * 128 * 4));
    int max_persist = 0;
    CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, 0));
    printf("rows of %d B, table %.0f MB, hub set %lld rows = %.1f MB taking %.0f %% of %lld gathers; max persisting L2 %.1f MB\n",
           row_bytes, n_rows * (double)row_bytes / 1e6, (long long)hub_rows, hub_rows * (double)row_bytes / 1e6, 100 * f,
           (long long)n_idx, max_persist / 1e6);
    const double bytes = (double)n_idx * row_bytes;
    const double ideal_dram = (1 - f) * bytes;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int policy = 0; policy <= 5; ++policy) {
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (policy == 2 || policy == 3 || policy == 4) ? (size_t)max_persist : 0));
        CK(cudaCtxResetPersistingL2Cache());
        auto launch = [&]() {
            const unsigned blocks = (unsigned)(warps / 8);
            if (policy == 3) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(blocks);
                cfg.blockDim = dim3(256);
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
                attr[0].val.accessPolicyWindow.base_ptr = X;
                attr[0].val.accessPolicyWindow.num_bytes = (size_t)hub_rows * row_bytes;
                attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
                attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                CK(cudaLaunchKernelEx(&cfg, gather<0>, (const float *)X, row_bytes, (const int *)idx, per_warp, (unsigned)hub_rows, out));
            } else if (policy == 0) {
                gather<0><<<blocks, 256>>>(X, row_bytes, idx, per_warp, (unsigned)hub_rows, out);
            } else if (policy == 4) {
                gather<4><<<blocks, 256>>>(X, row_bytes, idx, per_warp, (unsigned)hub_rows, out);
            } else if (policy == 5) {
                gather<5><<<blocks, 256>>>(X, row_bytes, idx, per_warp, (unsigned)hub_rows, out);
            } else {
                gather<1><<<blocks, 256>>>(X, row_bytes, idx, per_warp, (unsigned)hub_rows, out);
            }
        };
        launch();
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(a);
            launch();
            cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            best = std::min(best, ms);
        }
        CK(cudaGetLastError());
        printf("policy %d: %8.3f ms  %8.1f GB/s gathered  (ideal DRAM traffic %.2f GB -> %.3f ms at 6.5 TB/s)\n", policy, best,
               bytes / best / 1e6, ideal_dram / 1e9, ideal_dram / 6.5e9);
    }
    return 0;
}
