// Shared host/device helpers for the STiL head kernels (sm_100a only).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/stil_head.h"

namespace stil {

// ---- thread-local error message behind stil_last_error()
void set_error(const char* fmt, ...);
const char* last_error();

#define STIL_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            ::stil::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)

#define STIL_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::stil::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return STIL_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define STIL_LAUNCH_CHECK() STIL_CUDA(cudaGetLastError())

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- bump allocator over a caller-provided workspace (the extension allocates nothing)
struct Workspace {
    char* base;
    int64_t size, off;
    Workspace(void* p, int64_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
    template <class T>
    T* take(int64_t count) {
        int64_t bytes = round_up(count * (int64_t)sizeof(T), 256);
        char* p = base ? base + off : nullptr;
        off += bytes;
        return reinterpret_cast<T*>(p);
    }
    bool ok() const { return base == nullptr || off <= size; }
};

#ifdef __CUDACC__
// Launch with programmatic dependent launch allowed: the kernel may start while its stream predecessor is
// still draining; kernels call griddepcontrol.wait before touching global memory.
// All kernels of the library ask for the same (maximum-shared) L1/shared split as the GEMM kernel needs, so that
// back-to-back launches never make the SMs reconfigure their carve-out.
template <class K>
inline void prefer_max_shared(K kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// cudaFuncSetAttribute acts on the CURRENT device: opt a kernel into its dynamic shared memory (and the max-shared carve-out)
// once per device.  `mask` is a function-local static of the call site (one per kernel instantiation), bit = device ordinal.
inline cudaError_t ensure_dynamic_smem(std::atomic<unsigned long long>& mask, const void* kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    mask.fetch_or(bit, std::memory_order_release);
    return cudaSuccess;
}

bool pdl_enabled();            // api.cu (stil_debug_pdl)

// cluster_x > 1: the grid is launched as thread-block clusters of cluster_x consecutive blocks (gridDim.x % cluster_x == 0)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
    static const bool no_pdl = [] { const char* e = getenv("STIL_NO_PDL"); return e && e[0] == '1'; }();
    if (!no_pdl && pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // STIL_NO_PDL=1 / stil_debug_pdl(0) (debug / measurement): plain stream serialisation for every launch, so that a
    // profiler's per-kernel duration is the kernel's own time (with PDL a dependent's duration includes its wait)
    static const bool no_pdl = [] { const char* e = getenv("STIL_NO_PDL"); return e && e[0] == '1'; }();
    cfg.numAttrs = (no_pdl || !pdl_enabled()) ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers
// Waits on a PEER rank (arrival flags, LL words of the peer-memory exchanges) are bounded by WALL-CLOCK time, not by a spin
// count: skew between ranks at the head step is not bounded by any earlier collective (a data-loader stall at an epoch
// boundary, a checkpoint written by rank 0, a lazy graph capture on one rank), so the default is minutes — NCCL would simply
// wait; a peer that is really gone still becomes a CUDA error instead of a hung GPU.  Override at build time with
// -DSTIL_PEER_TIMEOUT_S=<seconds> (0 = wait for ever).
#ifndef STIL_PEER_TIMEOUT_S
#define STIL_PEER_TIMEOUT_S 600
#endif
__device__ __forceinline__ unsigned long long peer_wait_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// call once per spin with the iteration count; reads the clock every 4096 spins only
__device__ __forceinline__ void peer_wait_check(unsigned long long& spins, unsigned long long& t0) {
    if ((++spins & 4095ull) != 0) return;
    if (STIL_PEER_TIMEOUT_S == 0) return;
    const unsigned long long now = peer_wait_now();
    if (t0 == 0) { t0 = now; return; }
    if (now - t0 > (unsigned long long)STIL_PEER_TIMEOUT_S * 1000000000ull) __trap();
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int W>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float ld_as_float(const void* p, int dtype, int64_t i) {
    return dtype == STIL_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i])
                              : static_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_from_float(void* p, int dtype, int64_t i, float v) {
    if (dtype == STIL_BF16)
        static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else
        static_cast<float*>(p)[i] = v;
}

// Deterministic grid-wide scalar sum: every block stores its partial, the last block to take a
// ticket adds them in index order.  `ticket` must be zero on entry and is reset for the next launch.
__device__ __forceinline__ void ticket_sum(float block_value, float* partials, unsigned int* ticket, float* out,
                                           float scale) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_value;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        float s = 0.f;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += reinterpret_cast<volatile float*>(partials)[b];
        *out = s * scale;
        *ticket = 0u;
    }
}

// Sum `v` over the block (blockDim.x multiple of 32, <= 1024); result valid in thread 0.
__device__ __forceinline__ float block_sum(float v) {
    __shared__ float red[32];
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float s = 0.f;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        s = l < nw ? red[l] : 0.f;
        s = warp_sum(s);
    }
    return s;
}
#endif  // __CUDACC__

}  // namespace stil
