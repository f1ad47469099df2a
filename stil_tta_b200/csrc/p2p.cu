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
    int wait;                         // 1: all-gather (retire once every peer's segments have landed here); 0: push only
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

// Tail shared by every pushing kernel: once all blocks' remote stores are fenced, the last block publishes the
// arrival flag (the channel's sequence number) in every rank's buffer and, for an all-gather, waits for the peers'.
struct P2PChannel {
    unsigned char* base[kMaxWorld];
    int world, rank;
    long long flags_offset, ctrl_offset;
    int channel;
};
__device__ __forceinline__ void red_release_sys_add(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Push protocol (no waiting, ONE system-scope release per block on the critical path — a fence.sys costs 2-4 us on
// NVLink, so nothing here chains two of them): after the block barrier, thread p adds 1 to this rank's arrival counter in
// peer p's buffer with a release reduction (cumulative over the block's stores).  A consumer waits until the counter
// reaches the channel's target, which the last block to finish advances locally by the number of blocks.
__device__ __forceinline__ void p2p_publish(const P2PChannel& C) {
    unsigned char* self = C.base[C.rank];
    unsigned long long* target = reinterpret_cast<unsigned long long*>(self + C.ctrl_offset) + C.channel;
    unsigned int* done_ctr = reinterpret_cast<unsigned int*>(self + C.ctrl_offset + 64 * sizeof(unsigned long long)) + C.channel;
    __syncthreads();
    if ((int)threadIdx.x < C.world)
        red_release_sys_add(reinterpret_cast<unsigned long long*>(C.base[threadIdx.x] + C.flags_offset) +
                                C.channel * kMaxWorld + C.rank, 1ull);
    if (threadIdx.x == 0) {
        const unsigned int nblk = gridDim.x * gridDim.y;
        if (atomicAdd(done_ctr, 1u) == nblk - 1) {
            *done_ctr = 0u;
            *target = *target + nblk;      // read by the consumers launched after this kernel
        }
    }
}

// Wait until every peer's push number *seq_ctr on `channel` has landed in the local buffer (for consumers that are not
// flag-aware themselves)
__global__ void p2p_wait_kernel(const P2PChannel C) {
    unsigned char* self = C.base[C.rank];
    const unsigned long long seq = *(reinterpret_cast<const unsigned long long*>(self + C.ctrl_offset) + C.channel);
    if ((int)threadIdx.x < C.world) {
        const unsigned long long* my_flag =
            reinterpret_cast<const unsigned long long*>(self + C.flags_offset) + C.channel * kMaxWorld + threadIdx.x;
        unsigned long long spins = 0, t0 = 0;
        while (ld_acquire_sys(my_flag) < seq) peer_wait_check(spins, t0);
    }
}

// Rows of [feat_i | feat_t] and their inverse L2 norms, written straight into every rank's gathered buffers (one warp
// per local row): the pack, the normalisation pass and the all-gather of the global-batch InfoNCE in one kernel.
struct PushEmbed {
    P2PChannel C;
    const void* feat_i;
    const void* feat_t;
    int dtype, rows, dim;
    long long ab_offset;     // byte offset of the gathered [n, 2*dim] matrix inside every buffer
    long long ra_offset, rb_offset;   // byte offsets of the gathered inverse-norm vectors [n] f32
    int row0;                // global index of local row 0
};
__global__ void __launch_bounds__(kP2PBlock) p2p_push_embeddings_kernel(const PushEmbed X) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (kP2PBlock / 32) + warp;
    if (r < X.rows) {
        const int esz = X.dtype == STIL_BF16 ? 2 : 4;
        const long long row_bytes = (long long)X.dim * esz;        // multiple of 16
        const int n16 = (int)(row_bytes >> 4);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const unsigned char* src = static_cast<const unsigned char*>(side == 0 ? X.feat_i : X.feat_t) + (long long)r * row_bytes;
            float ss = 0.f;
            for (int i = lane; i < n16; i += 32) {
                const uint4 v = reinterpret_cast<const uint4*>(src)[i];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                if (X.dtype == STIL_BF16) {
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const float lo = __uint_as_float(w[h] << 16), hi = __uint_as_float(w[h] & 0xffff0000u);
                        ss += lo * lo + hi * hi;
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < 4; ++h) ss += __uint_as_float(w[h]) * __uint_as_float(w[h]);
                }
                for (int p = 0; p < X.C.world; ++p) {
                    unsigned char* dst = X.C.base[(X.C.rank + p) % X.C.world] + X.ab_offset +
                                         ((long long)(X.row0 + r) * 2 + side) * row_bytes;
                    reinterpret_cast<uint4*>(dst)[i] = v;
                }
            }
            ss = warp_sum(ss);
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);   // F.normalize eps (clip_loss.py:29-30)
            if (lane < X.C.world) {
                float* dst = reinterpret_cast<float*>(X.C.base[(X.C.rank + lane) % X.C.world] +
                                                      (side == 0 ? X.ra_offset : X.rb_offset)) + X.row0 + r;
                *dst = inv;
            }
        }
    }
    p2p_publish(X.C);
}

// Row LSEs of the local rows of both InfoNCE sides, merged from the GEMM_STATS partials and written into every rank's
// gathered LSE vectors (one warp per row): the statistics merge and the LSE all-gather in one kernel.
struct PushLse {
    P2PChannel C;
    const unsigned long long* tag;   // LL step tag: low 32 bits of this word (the embeddings channel's arrival target)
    const float* pmax[2];
    const float* psum[2];    // [slots, m] per side
    int slots, m;
    long long lse_offset[2]; // byte offsets of the gathered lse_row / lse_col vectors [n] f32
    int row0;
};
__global__ void __launch_bounds__(kP2PBlock) p2p_push_lse_kernel(const PushLse X) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gr = blockIdx.x * (kP2PBlock / 32) + warp;
    if (gr < 2 * X.m) {
        const int side = gr / X.m, i = gr - side * X.m;
        const float* pm = X.pmax[side];
        const float* ps = X.psum[side];
        float vm[2], vs[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int t = lane + 32 * q;
            vm[q] = t < X.slots ? pm[(long long)t * X.m + i] : -INFINITY;
            vs[q] = t < X.slots ? ps[(long long)t * X.m + i] : 0.f;
        }
        float mx = fmaxf(vm[0], vm[1]);
        for (int t = lane + 64; t < X.slots; t += 32) mx = fmaxf(mx, pm[(long long)t * X.m + i]);
        mx = warp_max(mx);
        float sm = vs[0] * expf(vm[0] - mx) + vs[1] * expf(vm[1] - mx);
        for (int t = lane + 64; t < X.slots; t += 32) sm += ps[(long long)t * X.m + i] * expf(pm[(long long)t * X.m + i] - mx);
        sm = warp_sum(sm);
        const float lse = mx + logf(sm);
        // flag-less: value and step tag travel in ONE 8-byte store; readers spin on the word itself
        const unsigned long long word = ((unsigned long long)(unsigned int)(*X.tag) << 32) | __float_as_uint(lse);
        if (lane < X.C.world)
            reinterpret_cast<unsigned long long*>(X.C.base[(X.C.rank + lane) % X.C.world] + X.lse_offset[side])[X.row0 + i] = word;
    }
}

// generic push of up to kMaxSeg byte segments (same arrival-counter protocol)
__global__ void __launch_bounds__(kP2PBlock) p2p_push_kernel(const P2PExchange X) {
    for (int s = 0; s < X.nseg; ++s) {
        const uint4* src = reinterpret_cast<const uint4*>(X.src[s]);
        const long long n16 = X.nbytes[s] >> 4;
        for (int p = 0; p < X.world; ++p) {
            uint4* dst = reinterpret_cast<uint4*>(X.base[(X.rank + p) % X.world] + X.dst_offset[s]);
            for (long long i = (long long)blockIdx.x * kP2PBlock + threadIdx.x; i < n16; i += (long long)gridDim.x * kP2PBlock)
                dst[i] = src[i];
        }
    }
    P2PChannel C;
    for (int p = 0; p < kMaxWorld; ++p) C.base[p] = X.base[p];
    C.world = X.world; C.rank = X.rank; C.flags_offset = X.flags_offset; C.ctrl_offset = X.ctrl_offset; C.channel = X.channel;
    p2p_publish(C);
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
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        __threadfence_system();      // one system-scope fence per block (cumulative over the block's stores)
        is_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (!is_last) return;
    // ---- last block: publish arrival to every peer, then wait until every peer's data has landed here
    if ((int)threadIdx.x < X.world) {
        const int p = threadIdx.x;
        unsigned long long* their_flag =
            reinterpret_cast<unsigned long long*>(X.base[p] + X.flags_offset) + X.channel * kMaxWorld + X.rank;
        __threadfence_system();
        st_release_sys(their_flag, seq);
        if (X.wait) {
            const unsigned long long* my_flag =
                reinterpret_cast<const unsigned long long*>(self + X.flags_offset) + X.channel * kMaxWorld + p;
            unsigned long long spins = 0, t0 = 0;   // wall-clock bound (common.cuh): a lost peer becomes a CUDA error
            while (ld_acquire_sys(my_flag) < seq) peer_wait_check(spins, t0);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *done_ctr = 0u;
        *seq_ctr = seq;
        __threadfence_system();
    }
}

int fill_channel(P2PChannel& C, void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                 int channel) {
    STIL_REQUIRE(bases && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, STIL_E_ARG, "p2p: bad world/rank");
    STIL_REQUIRE(channel >= 0 && channel < 8, STIL_E_ARG, "p2p: bad channel");
    std::memset(&C, 0, sizeof(C));
    for (int p = 0; p < world; ++p) {
        STIL_REQUIRE(bases[p] != nullptr, STIL_E_ARG, "p2p: null peer buffer %d", p);
        C.base[p] = static_cast<unsigned char*>(bases[p]);
    }
    C.world = world; C.rank = rank; C.flags_offset = flags_offset; C.ctrl_offset = ctrl_offset; C.channel = channel;
    return STIL_OK;
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

static int p2p_exchange_impl(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                             int channel, int nseg, const void* const* src, const int64_t* nbytes,
                             const int64_t* dst_offset, void* stream, int wait) {
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
    X.channel = channel; X.nseg = nseg; X.wait = wait;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(32, most / (16 * kP2PBlock * 2)));
    if (wait)
        p2p_exchange_kernel<<<blocks, kP2PBlock, 0, static_cast<cudaStream_t>(stream)>>>(X);
    else
        p2p_push_kernel<<<blocks, kP2PBlock, 0, static_cast<cudaStream_t>(stream)>>>(X);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

STIL_API int stil_p2p_exchange(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                               int channel, int nseg, const void* const* src, const int64_t* nbytes,
                               const int64_t* dst_offset, void* stream) {
    return p2p_exchange_impl(bases, world, rank, flags_offset, ctrl_offset, channel, nseg, src, nbytes, dst_offset, stream, 1);
}
STIL_API int stil_p2p_push(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset, int channel,
                           int nseg, const void* const* src, const int64_t* nbytes, const int64_t* dst_offset, void* stream) {
    return p2p_exchange_impl(bases, world, rank, flags_offset, ctrl_offset, channel, nseg, src, nbytes, dst_offset, stream, 0);
}
STIL_API int stil_p2p_wait(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset, int channel,
                           void* stream) {
    P2PChannel C;
    int rc = fill_channel(C, bases, world, rank, flags_offset, ctrl_offset, channel);
    if (rc) return rc;
    p2p_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(C);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}
STIL_API int stil_p2p_push_embeddings(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                                      int channel, const void* feat_i, const void* feat_t, int dtype, int64_t rows,
                                      int64_t dim, int64_t row0, int64_t ab_offset, int64_t ra_offset, int64_t rb_offset,
                                      void* stream) {
    PushEmbed X;
    std::memset(&X, 0, sizeof(X));
    int rc = fill_channel(X.C, bases, world, rank, flags_offset, ctrl_offset, channel);
    if (rc) return rc;
    STIL_REQUIRE(dtype == STIL_F32 || dtype == STIL_BF16, STIL_E_DTYPE, "p2p_push_embeddings: bad dtype");
    const int per16 = dtype == STIL_BF16 ? 8 : 4;
    STIL_REQUIRE(feat_i && feat_t && rows >= 1 && dim >= per16 && dim % per16 == 0 && ab_offset % 16 == 0 &&
                     ra_offset % 4 == 0 && rb_offset % 4 == 0 && (reinterpret_cast<uintptr_t>(feat_i) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(feat_t) & 15) == 0,
                 STIL_E_ALIGN, "p2p_push_embeddings: rows must be 16-byte granular and aligned");
    X.feat_i = feat_i; X.feat_t = feat_t; X.dtype = dtype; X.rows = (int)rows; X.dim = (int)dim;
    X.ab_offset = ab_offset; X.ra_offset = ra_offset; X.rb_offset = rb_offset; X.row0 = (int)row0;
    p2p_push_embeddings_kernel<<<(unsigned)ceil_div(rows, kP2PBlock / 32), kP2PBlock, 0, static_cast<cudaStream_t>(stream)>>>(X);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}
STIL_API int stil_p2p_push_lse(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                               int tag_channel, const void* infonce_workspace, int64_t m, int64_t n, int64_t dim, int dtype,
                               int64_t row0, int64_t lse_row_offset, int64_t lse_col_offset, void* stream) {
    PushLse X;
    std::memset(&X, 0, sizeof(X));
    int rc = fill_channel(X.C, bases, world, rank, flags_offset, ctrl_offset, tag_channel);
    if (rc) return rc;
    STIL_REQUIRE(infonce_workspace && m >= 1 && n >= m && lse_row_offset % 8 == 0 && lse_col_offset % 8 == 0 && tag_channel >= 0 &&
                     tag_channel < 8,
                 STIL_E_ARG, "p2p_push_lse: bad arguments");
    X.tag = reinterpret_cast<const unsigned long long*>(X.C.base[rank] + ctrl_offset) + tag_channel;
    int slots = 0;
    infonce_stat_partials(const_cast<void*>(infonce_workspace), m, n, dim, dtype, X.pmax, X.psum, &slots);
    X.slots = slots; X.m = (int)m; X.row0 = (int)row0;
    X.lse_offset[0] = lse_row_offset; X.lse_offset[1] = lse_col_offset;
    p2p_push_lse_kernel<<<(unsigned)ceil_div(2 * m, kP2PBlock / 32), kP2PBlock, 0, static_cast<cudaStream_t>(stream)>>>(X);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

}  // extern "C"
