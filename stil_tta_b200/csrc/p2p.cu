// Low-latency exchange between the GPUs of one node over peer-mapped memory (NVLink): every rank writes its slot
// straight into every peer's buffer with 16-byte remote stores, publishes an arrival flag, and waits for the flags of
// all peers — an all-gather whose cost is a few microseconds for the 4 KB - 256 KB messages of the STiL head
// (embeddings for the global-batch InfoNCE, LSE vectors, prototype partial sums).  Buffers come from cudaMalloc and
// are shared with CUDA IPC so the mechanism needs nothing but the CUDA runtime.
#include <algorithm>
#include <cstring>

#include "internal.h"

namespace stil {
namespace {

constexpr int kMaxWorld = 8;
constexpr int kMaxSeg = 4;
constexpr int kP2PBlock = 256;

struct P2PExchange {
    unsigned char* base[kMaxWorld];   // every rank's buffer (own entry: the local pointer)
    int world, rank;
    long long flags_offset;           // u64 flags[channels][kMaxWorld] inside every buffer
    long long ctrl_offset;            // local control words: u64 seq[channels], u32 done[channels]
    int channel;
    int nseg;
    const unsigned char* src[kMaxSeg];
    long long nbytes[kMaxSeg];        // multiples of 16
    long long dst_offset[kMaxSeg];    // where this rank's slot lives inside every buffer
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kP2PBlock) p2p_exchange_kernel(const P2PExchange X) {
    unsigned char* self = X.base[X.rank];
    unsigned long long* seq_ctr = reinterpret_cast<unsigned long long*>(self + X.ctrl_offset) + X.channel;
    unsigned int* done_ctr = reinterpret_cast<unsigned int*>(self + X.ctrl_offset + 64 * sizeof(unsigned long long)) + X.channel;
    const unsigned long long seq = *seq_ctr + 1;   // stable for the whole kernel: only its last block bumps it
    // ---- push this rank's segments into every buffer (self included), 16 bytes per store
    for (int s = 0; s < X.nseg; ++s) {
        const uint4* src = reinterpret_cast<const uint4*>(X.src[s]);
        const long long n16 = X.nbytes[s] >> 4;
        for (int p = 0; p < X.world; ++p) {
            uint4* dst = reinterpret_cast<uint4*>(X.base[(X.rank + p) % X.world] + X.dst_offset[s]);
            for (long long i = (long long)blockIdx.x * kP2PBlock + threadIdx.x; i < n16; i += (long long)gridDim.x * kP2PBlock)
                dst[i] = src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x == 0) is_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    // ---- last block: publish arrival to every peer, then wait until every peer's data has landed here
    if ((int)threadIdx.x < X.world) {
        const int p = threadIdx.x;
        unsigned long long* their_flag =
            reinterpret_cast<unsigned long long*>(X.base[p] + X.flags_offset) + X.channel * kMaxWorld + X.rank;
        __threadfence_system();
        st_release_sys(their_flag, seq);
        const unsigned long long* my_flag =
            reinterpret_cast<const unsigned long long*>(self + X.flags_offset) + X.channel * kMaxWorld + p;
        unsigned long long spins = 0;
        while (ld_acquire_sys(my_flag) < seq) {
            if (++spins > (1ull << 26)) __trap();   // a lost peer becomes a CUDA error, not a hung GPU
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *done_ctr = 0u;
        *seq_ctr = seq;
        __threadfence_system();
    }
}

}  // namespace
}  // namespace stil

using namespace stil;

extern "C" {

STIL_API int stil_p2p_alloc(int64_t bytes, void** ptr) {
    STIL_REQUIRE(ptr && bytes > 0, STIL_E_ARG, "p2p_alloc: bad arguments");
    STIL_CUDA(cudaMalloc(ptr, (size_t)bytes));
    STIL_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    STIL_CUDA(cudaDeviceSynchronize());
    return STIL_OK;
}
STIL_API int stil_p2p_free(void* ptr) {
    STIL_CUDA(cudaFree(ptr));
    return STIL_OK;
}
STIL_API int stil_p2p_export(void* ptr, uint8_t* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    STIL_CUDA(cudaIpcGetMemHandle(&h, ptr));
    std::memcpy(handle64, &h, 64);
    return STIL_OK;
}
STIL_API int stil_p2p_import(const uint8_t* handle64, void** peer_ptr) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    STIL_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return STIL_OK;
}
STIL_API int stil_p2p_close(void* peer_ptr) {
    STIL_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return STIL_OK;
}

STIL_API int stil_p2p_exchange(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                               int channel, int nseg, const void* const* src, const int64_t* nbytes,
                               const int64_t* dst_offset, void* stream) {
    STIL_REQUIRE(bases && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, STIL_E_ARG, "p2p_exchange: bad world/rank");
    STIL_REQUIRE(nseg >= 1 && nseg <= kMaxSeg && channel >= 0 && channel < 8, STIL_E_ARG, "p2p_exchange: bad nseg/channel");
    P2PExchange X;
    std::memset(&X, 0, sizeof(X));
    long long most = 0;
    for (int p = 0; p < world; ++p) {
        STIL_REQUIRE(bases[p] != nullptr, STIL_E_ARG, "p2p_exchange: null peer buffer %d", p);
        X.base[p] = static_cast<unsigned char*>(bases[p]);
    }
    for (int s = 0; s < nseg; ++s) {
        STIL_REQUIRE(src[s] && nbytes[s] > 0 && nbytes[s] % 16 == 0 && dst_offset[s] % 16 == 0 &&
                         (reinterpret_cast<uintptr_t>(src[s]) & 15) == 0,
                     STIL_E_ALIGN, "p2p_exchange: segment %d must be 16-byte aligned and sized", s);
        X.src[s] = static_cast<const unsigned char*>(src[s]);
        X.nbytes[s] = nbytes[s];
        X.dst_offset[s] = dst_offset[s];
        most = std::max<long long>(most, nbytes[s]);
    }
    X.world = world; X.rank = rank; X.flags_offset = flags_offset; X.ctrl_offset = ctrl_offset;
    X.channel = channel; X.nseg = nseg;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(32, most / (16 * kP2PBlock * 2)));
    p2p_exchange_kernel<<<blocks, kP2PBlock, 0, static_cast<cudaStream_t>(stream)>>>(X);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

}  // extern "C"
