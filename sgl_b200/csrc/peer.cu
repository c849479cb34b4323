// peer.cu -- halo exchange over NVLink peer memory, without NCCL on the data path.
//
// The row partition (sgl_b200/dist.py) keeps two extended feature slabs per rank in memory that every other rank of
// the box has mapped (CUDA IPC).  After a tile range of the hop has produced its rows, push_rows_kernel copies the rows
// each peer references straight into that peer's halo block: one warp per row, 128-bit loads from local HBM (the rows
// were just written: L2 hits) and 128-bit stores over NVLink.  Completion is published with a system-scope release
// store into a flag word in the peer's memory; the next hop starts behind a kernel that acquires all peers' flags.
// Measured motivation (profiles/): NCCL all_to_all of ~100 MB chunks reaches 130-250 GB/s between two B200s and
// serialises on its own stream (2.2 ms per hop for 417 MB), NVLink peer stores are limited by the link (~700 GB/s).
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace sglb200 {

constexpr int kPushWarps = 8;

template <int VEC>
__global__ void __launch_bounds__(kPushWarps * 32)
    push_rows_kernel(const float *__restrict__ src, int64_t ld_src, int d, const int64_t *__restrict__ rows,
                     int64_t n_rows, float *__restrict__ dst, int64_t ld_dst)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kPushWarps + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * kPushWarps;
    // two rows in flight per warp
    for (int64_t i = warp0; i < n_rows; i += 2 * n_warps) {
        const int64_t i2 = i + n_warps;
        const bool two = i2 < n_rows;
        const float *a = src + rows[i] * ld_src;
        const float *b = src + (two ? rows[i2] : rows[i]) * ld_src;
        float *da = dst + i * ld_dst;
        float *db = dst + i2 * ld_dst;
        for (int c = lane * VEC; c < d; c += 32 * VEC) {
            if constexpr (VEC == 4) {
                const float4 va = *reinterpret_cast<const float4 *>(a + c);
                const float4 vb = *reinterpret_cast<const float4 *>(b + c);
                *reinterpret_cast<float4 *>(da + c) = va;
                if (two) *reinterpret_cast<float4 *>(db + c) = vb;
            } else {
                const float va = a[c], vb = b[c];
                da[c] = va;
                if (two) db[c] = vb;
            }
        }
    }
}

// one thread per peer: publish `value` in that peer's flag word (after everything this stream has written)
__global__ void signal_peers_kernel(unsigned long long *const *flag_ptrs, int n, unsigned long long value)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_ptrs[i]), "l"(value) : "memory");
}

// spins until every one of the n local flag words has reached `value` -- but never forever: a peer that died would
// otherwise hang this GPU.  After `timeout_ns` (globaltimer) the kernel gives up and raises the error word, which the
// next sglb200_wait_flags / sglb200_peer_status call reports (the hop results of that step are then undefined).
__global__ void wait_flags_kernel(const unsigned long long *flags, int n, unsigned long long value,
                                  unsigned long long timeout_ns, unsigned int *error_word)
{
    const int i = threadIdx.x;
    if (i < n) {
        unsigned long long seen, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned int polls = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + i) : "memory");
            if (seen >= value) break;
            if ((++polls & 1023u) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > timeout_ns) {
                    atomicOr(error_word, 1u << (i & 31));
                    break;
                }
            }
        } while (true);
    }
    __syncthreads();
}

}  // namespace sglb200

using namespace sglb200;

extern "C" {

int sglb200_ipc_alloc(int64_t bytes, void **ptr, unsigned char handle[64])
{
    clear_error();
    SGL_REQUIRE(bytes > 0 && ptr && handle, "ipc_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    {
        const int st = check_device();
        if (st != SGLB200_OK) return st;
    }
    void *p = nullptr;
    SGL_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        SGL_CUDA_CHECK(e);
    }
    memcpy(handle, &h, 64);
    *ptr = p;
    return SGLB200_OK;
}

int sglb200_ipc_open(const unsigned char handle[64], void **ptr)
{
    clear_error();
    SGL_REQUIRE(ptr && handle, "ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SGL_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SGLB200_OK;
}

int sglb200_ipc_close(void *ptr)
{
    clear_error();
    if (ptr) SGL_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return SGLB200_OK;
}

int sglb200_ipc_free(void *ptr)
{
    clear_error();
    if (ptr) SGL_CUDA_CHECK(cudaFree(ptr));
    return SGLB200_OK;
}

int sglb200_push_rows(const float *src, int64_t ld_src, int d, const int64_t *rows, int64_t n_rows, float *dst,
                      int64_t ld_dst, int max_blocks, void *stream)
{
    clear_error();
    SGL_REQUIRE(n_rows >= 0 && d >= 0, "push_rows: negative size");
    if (n_rows == 0 || d == 0) return SGLB200_OK;
    SGL_REQUIRE(src && rows && dst && ld_src >= d && ld_dst >= d, "push_rows: bad argument");
    int64_t blocks = (n_rows + 2 * kPushWarps - 1) / (2 * kPushWarps);
    const int cap = max_blocks > 0 ? max_blocks : 148 * 2;
    if (blocks > cap) blocks = cap;
    const bool vec4 = d % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    if (vec4)
        push_rows_kernel<4><<<(unsigned)blocks, kPushWarps * 32, 0, (cudaStream_t)stream>>>(src, ld_src, d, rows, n_rows, dst, ld_dst);
    else
        push_rows_kernel<1><<<(unsigned)blocks, kPushWarps * 32, 0, (cudaStream_t)stream>>>(src, ld_src, d, rows, n_rows, dst, ld_dst);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

int sglb200_signal_peers(unsigned long long *const *flag_ptrs_dev, int n, unsigned long long value, void *stream)
{
    clear_error();
    SGL_REQUIRE(n >= 0 && (n == 0 || flag_ptrs_dev), "signal_peers: bad argument");
    if (n == 0) return SGLB200_OK;
    signal_peers_kernel<<<1, 32 * ((n + 31) / 32), 0, (cudaStream_t)stream>>>(flag_ptrs_dev, n, value);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

// per-device error word of the flag waits (mapped pinned host memory: readable without synchronising the device)
static unsigned int *peer_error_word(int create)
{
    static unsigned int *words[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!words[dev] && create) {
        unsigned int *w = nullptr;
        if (cudaHostAlloc(&w, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        *w = 0;
        words[dev] = w;
    }
    return words[dev];
}

int sglb200_wait_flags(const unsigned long long *flags_dev, int n, unsigned long long value, void *stream)
{
    clear_error();
    SGL_REQUIRE(n >= 0 && n <= 1024 && (n == 0 || flags_dev), "wait_flags: bad argument");
    if (n == 0) return SGLB200_OK;
    unsigned int *err = peer_error_word(1);
    SGL_REQUIRE(err != nullptr, "wait_flags: cannot allocate the error word");
    if (*err != 0) {
        set_error("wait_flags: an earlier wait timed out (peer mask 0x%x): a peer rank is dead or stalled", *err);
        return SGLB200_ERR_CUDA;
    }
    static unsigned long long timeout_ns = 0;
    if (timeout_ns == 0) {
        const char *e = getenv("SGLB200_PEER_TIMEOUT_MS");
        timeout_ns = (unsigned long long)(e ? atoll(e) : 30000) * 1000000ULL;   // default 30 s
    }
    unsigned int *err_dev = nullptr;
    SGL_CUDA_CHECK(cudaHostGetDevicePointer(&err_dev, err, 0));
    wait_flags_kernel<<<1, 32 * ((n + 31) / 32), 0, (cudaStream_t)stream>>>(flags_dev, n, value, timeout_ns, err_dev);
    SGL_CUDA_CHECK(cudaGetLastError());
    return SGLB200_OK;
}

/* 0 when no flag wait has timed out on the current device, else the mask of flag slots that did (and clears it) */
int sglb200_peer_status(void)
{
    unsigned int *err = peer_error_word(0);
    if (!err) return 0;
    const int v = (int)*err;
    *err = 0;
    return v;
}

}  // extern "C"
