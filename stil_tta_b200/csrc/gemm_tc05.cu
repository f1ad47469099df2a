// Similarity GEMM with fused softmax-statistics / gradient epilogues on tcgen05 (sm_100a).
//
//   L[i,j] = alpha * sx[i] * sy[j] * sum_{(p,q) in pairs} <X[i,p,:], Y[j,q,:]>
//
// One CTA per 128x128 tile.  Warp 0 streams 128x64 bf16 operand boxes with TMA (128-byte swizzle)
// through an mbarrier ring (6 stages with one CTA per SM, 3 with two), warp 1 issues tcgen05.mma (M=128,
// N=128, K=16, bf16 -> fp32 in TMEM), the remaining 16 (or 8) warps drain the accumulator with tcgen05.ld
// (thread = row, one or two 32-column chunks per warp) and run one of three epilogues (see GemmMode in internal.h).  Used for: the InfoNCE logits of utils/clip_loss.py:33 and
// their row/column log-sum-exp (:36-37), the prototype logits of utils/prototype_loss.py:26 and
// STiLModel.py:293, and both GEMMs of their backward passes (G = dLoss/dLogits is formed on chip from a
// recomputed tile and written as bf16 hi/lo; dX = G·Y reads Y in place as an MN-major operand and applies
// the backward of F.normalize in its epilogue).
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <mutex>

#include "internal.h"
#include "tc05.cuh"

namespace stil {

namespace {

// Ring depth by occupancy: small grids (<= one CTA per SM) run one CTA per SM with a 6-deep ring (the whole K loop of
// the reference sizes is in flight at once); grids of more than 148 tiles run TWO CTAs per SM with 3-deep rings so that
// one CTA's epilogue overlaps the other's TMA/MMA main loop (the accumulator is single-buffered).
constexpr int kMaxStages = 6;
// TMA warp, MMA warp, then the epilogue warps: 16 when one CTA owns the SM (each thread drains ONE 32-column chunk of its
// row — the epilogue is the long pole of the small, latency-bound launches), 8 when two CTAs share it (two chunks each)
constexpr int epi_warps_for(int occ) { return occ == 1 ? 16 : 8; }
constexpr int threads_for(int occ) { return 64 + 32 * epi_warps_for(occ); }
constexpr int kStageBytes = (kTileM + kTileN) * kTileK * 2;  // 32 KiB
constexpr int smem_bytes_for(int stages) {
    return stages * kStageBytes + 1024 /*align slack*/ + 3072 /*scales, dot partials*/ + 256 /*barriers*/;
}
constexpr int kStageFloats = 32 * 33;   // per-warp staging tile for coalesced epilogue stores (reuses the ring)
constexpr uint32_t kTmemCols = 128;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// log-sum-exp of one row/column from its (max, sum) partials laid out [slots, stride]; eight slots are loaded
// together so the merge costs one memory round trip
__device__ __forceinline__ float merge_partials(const float* pmax, const float* psum, int slots, long long stride,
                                                int idx) {
    float run_m = -INFINITY, run_s = 0.f;
    for (int t0 = 0; t0 < slots; t0 += 16) {
        float m[16], s[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const bool ok = t0 + j < slots;
            m[j] = ok ? __ldcg(pmax + (long long)(t0 + j) * stride + idx) : -INFINITY;
            s[j] = ok ? __ldcg(psum + (long long)(t0 + j) * stride + idx) : 0.f;
        }
        float nm = run_m;
#pragma unroll
        for (int j = 0; j < 16; ++j) nm = fmaxf(nm, m[j]);
        if (nm == -INFINITY) continue;
        float acc = run_s * __expf(run_m - nm);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += s[j] * __expf(m[j] - nm);
        run_m = nm;
        run_s = acc;
    }
    return run_m + logf(run_s);
}

__device__ __forceinline__ void load32_as_float(const void* base, int dtype, long long off, int nv, float (&x)[32]) {
    if (dtype == STIL_BF16) {
        const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(base) + off;
        if (nv == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint4 v = *reinterpret_cast<const uint4*>(p + j);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    x[j + 2 * h] = __uint_as_float(w[h] << 16);
                    x[j + 2 * h + 1] = __uint_as_float(w[h] & 0xffff0000u);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = j < nv ? __bfloat162float(p[j]) : 0.f;
        }
    } else {
        const float* p = static_cast<const float*>(base) + off;
        if (nv == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(p + j);
                x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = j < nv ? p[j] : 0.f;
        }
    }
}

__device__ __forceinline__ void store32_from_float(void* base, int dtype, long long off, int nv, const float (&x)[32]) {
    if (dtype == STIL_BF16) {
        __nv_bfloat16* p = static_cast<__nv_bfloat16*>(base) + off;
        if (nv == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t w[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const __nv_bfloat162 b = __floats2bfloat162_rn(x[j + 2 * h], x[j + 2 * h + 1]);
                    w[h] = *reinterpret_cast<const uint32_t*>(&b);
                }
                *reinterpret_cast<uint4*>(p + j) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nv) p[j] = __float2bfloat16_rn(x[j]);
        }
    } else {
        float* p = static_cast<float*>(base) + off;
        if (nv == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nv) p[j] = x[j];
        }
    }
}

// Optional timeline instrumentation (stil_debug_trace): when a buffer is installed every CTA stores %globaltimer
// stamps of its phases: [launch_id][cta][8] = start, prologue done, last TMA issued, last MMA committed, accumulator
// ready (epilogue), epilogue math done, epilogue end, mode.
unsigned long long* g_trace_host = nullptr;   // host copy; travels to the kernel in its parameters
constexpr int kTraceCtas = 64, kTraceSlots = 8, kTraceLaunches = 64;
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define STIL_TRACE(slot)                                                                              \
    do {                                                                                              \
        if (trace && blockIdx.x < kTraceCtas) trace[(blockIdx.x) * kTraceSlots + (slot)] = gtime();   \
    } while (0)

// Arrival flags of the peer-memory all-gathers (csrc/p2p.cu): a consumer tile waits only for the peers whose rows it reads,
// so the transfer overlaps the tiles that need local data.  Wall-clock-bounded wait (common.cuh, STIL_PEER_TIMEOUT_S).
__device__ __forceinline__ void wait_peer_rows(const unsigned long long* flags, const unsigned long long* seq_ptr, int rpp,
                                               int r0, int r1) {
    unsigned long long seq;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seq) : "l"(seq_ptr) : "memory");
    // acquire loads, polled by ONE thread per role (a trailing fence.sys instead costs microseconds: measured)
    for (int p = r0 / rpp; p <= r1 / rpp; ++p) {
        unsigned long long v, spins = 0, t0 = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + p) : "memory");
            if (v >= seq) break;
            peer_wait_check(spins, t0);
            __nanosleep(20);
        }
    }
}

// "LL" words of a flag-less peer exchange (csrc/p2p.cu, p2p_push_lse_kernel): value in the low 32 bits, the step tag in
// the high 32 bits of ONE 8-byte store; the reader spins on the word itself until the tag matches — no fence anywhere.
__device__ __forceinline__ float ll_read(const float* base, long long idx, const unsigned long long* tag_ptr) {
    unsigned long long tag;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(tag) : "l"(tag_ptr) : "memory");
    const unsigned int want = (unsigned int)tag;
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(base) + idx;
    unsigned long long v, spins = 0, t0 = 0;
    for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
        if ((unsigned int)(v >> 32) == want) break;
        peer_wait_check(spins, t0);
        __nanosleep(32);
    }
    return __uint_as_float((unsigned int)v);
}

// GEMM_STORE post-ops: 1 = exp(value) (embedding graph, comatch_model.py:309-311), 2 = diagonal forced to 1
// (pseudo-label graph, comatch_model.py:299-300)
__device__ __forceinline__ float store_post(float v, int op, int row, int col) {
    if (op == 1) return fast_exp2(v * kLog2e);
    if (op == 2) return row == col ? 1.f : v;
    return v;
}

template <int MODE, int OCC>
__global__ void __launch_bounds__(threads_for(OCC), OCC) gemm_tc05_kernel(const __grid_constant__ GemmLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kStages = OCC == 1 ? kMaxStages : 3;
    constexpr int kEpiWarps = epi_warps_for(OCC);
    constexpr int kEpiThreads = 32 * kEpiWarps;
    constexpr int kParts = kEpiWarps / 4;        // column parts of the tile (one warp per TMEM lane quarter each)
    constexpr int kCPW = 4 / kParts;             // 32-column chunks per warp
    // carve: [stages x (A 16K | B 16K)] 1024-aligned, then scales, then barriers
    const uint32_t raw = tc05::smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* tiles = smem_raw + pad;
    float* col_scale = reinterpret_cast<float*>(tiles + kStages * kStageBytes);  // [128]
    float* col_lse = col_scale + kTileN;                                          // [128]
    float* dot_part = col_lse + kTileN;                                           // [kParts <= 4][128]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(dot_part + 4 * kTileM);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned long long* trace = L.trace ? L.trace + (size_t)(L.trace_id % kTraceLaunches) * kTraceCtas * kTraceSlots : nullptr;
    if (threadIdx.x == 0) { STIL_TRACE(0); if (trace && blockIdx.x < kTraceCtas) trace[blockIdx.x * kTraceSlots + 7] = MODE; }

    // ---- which job / tile
    int jid = 0;
#pragma unroll
    for (int j = 1; j < kMaxGemmJobs; ++j)
        if (j < L.njobs && (int)blockIdx.x >= L.job[j].tile_begin) jid = j;
    const GemmJob& J = L.job[jid];
    int t = blockIdx.x - J.tile_begin;
    const int ksplit = (MODE == GEMM_STORE && J.ksplit > 1) ? J.ksplit : 1;
    // cluster split-K (see GemmLaunch::cluster_k): the ksplit CTAs of a tile are consecutive blocks = one cluster
    const bool clustered = MODE == GEMM_STORE && L.cluster_k > 1;
    const int ks = t % ksplit;
    t /= ksplit;
    const int tm = t / J.tiles_n, tn = t % J.tiles_n;
    const int m0 = tm * kTileM, n0 = tn * kTileN;
    const int kper_all = (J.D + kTileK - 1) / kTileK;
    // this CTA's slice of the contraction (whole range unless split)
    const int kchunk = (kper_all + ksplit - 1) / ksplit;
    const int kk0 = ks * kchunk;
    const int kper = max(0, min(kchunk, kper_all - kk0));
    const int nkb = J.npair * kper;

    if (warp == 0 && lane == 0) {
        tc05::tma_prefetch_desc(&J.tmx);
        tc05::tma_prefetch_desc(&J.tmy);
        if (MODE == GEMM_GRAD || J.out_tma) tc05::tma_prefetch_desc(&J.tmg);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                tc05::mbar_init(&full_bar[s], 1);
                tc05::mbar_init(&empty_bar[s], 1);
            }
            tc05::mbar_init(tmem_full_bar, 1);
            tc05::fence_mbar_init();
        }
        __syncwarp();
        tc05::tmem_alloc(tmem_slot, kTmemCols);
        tc05::tmem_relinquish();
    }
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch.  Every thread that touches global memory the stream predecessor may have written
    // waits for it first; `launch_dependents` is issued AFTER that wait, so by the time the next kernel of the chain
    // starts, everything two or more kernels back is complete.  That is what lets the TMA producer stream operands the
    // immediate predecessor does not write (J.early_x / J.early_y) before its own wait: the main loop of this kernel then
    // overlaps the predecessor's execution and only the epilogue depends on it.
    if (warp == 0) {
        // ===================== TMA producer =====================
        // The first pass over the ring is issued LANE-PARALLEL: lane s fills stage s.  A bulk-tensor instruction costs
        // ~0.1 us of issue time; one thread issuing the 12-18 loads of a three-segment operand back to back spent 1.2-1.7 us
        // on the critical path of the first kernel of the step.
        {
            const bool mn = MODE == GEMM_STORE && J.y_mn_major != 0;
            const bool xmn = MODE == GEMM_STORE && J.x_mn_major != 0;
            auto load_x = [&](int kb, int s) {
                const int p = kb / kper, kk = kk0 + kb - p * kper;
                uint8_t* a_dst = tiles + s * kStageBytes;
                if (!xmn) {
                    tc05::tma_load_3d(a_dst, &J.tmx, &full_bar[s], kk * kTileK, m0, J.xseg[p]);
                } else {
                    // X tile [64 contraction rows x 128 M] as two 64x64 boxes (M chunks 8 KiB apart), like the MN-major Y tile
                    tc05::tma_load_3d(a_dst, &J.tmx, &full_bar[s], m0, kk * kTileK, J.xseg[p]);
                    tc05::tma_load_3d(a_dst + 64 * kTileK * 2, &J.tmx, &full_bar[s], m0 + 64, kk * kTileK, J.xseg[p]);
                }
            };
            auto load_y = [&](int kb, int s) {
                const int p = kb / kper, kk = kk0 + kb - p * kper;
                uint8_t* b_dst = tiles + s * kStageBytes + kTileM * kTileK * 2;
                if (!mn) {
                    tc05::tma_load_3d(b_dst, &J.tmy, &full_bar[s], kk * kTileK, n0, J.yseg[p]);
                } else {
                    // Y tile [64 contraction rows x 128 N] as two 64x64 boxes (N chunks 8 KiB apart)
                    tc05::tma_load_3d(b_dst, &J.tmy, &full_bar[s], n0, kk * kTileK, J.yseg[p]);
                    tc05::tma_load_3d(b_dst + 64 * kTileK * 2, &J.tmy, &full_bar[s], n0 + 64, kk * kTileK, J.yseg[p]);
                }
            };
            // operands that do not depend on the predecessor go out before the wait
            const int first = min(nkb, kStages);
            const bool ex = J.early_x != 0, ey = J.early_y != 0;
            const bool mine = lane < first;
            if ((ex || ey) && mine) {
                tc05::mbar_arrive_expect_tx(&full_bar[lane], kStageBytes);
                if (ex) load_x(lane, lane);
                if (ey) load_y(lane, lane);
            }
            if (!(ex && ey)) {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                if (J.wait_flags && J.wait_y) {
                    // the Y rows of this tile were stored by their owner rank(s): wait for exactly those arrivals (one lane,
                    // the warp barrier passes the acquire on), then order it before the TMA (async proxy) reads
                    if (lane == 0) wait_peer_rows(J.wait_flags, J.wait_seq, J.wait_rows_per_peer, n0, min(n0 + kTileN, J.N) - 1);
                    __syncwarp();
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                if (mine) {
                    if (!(ex || ey)) tc05::mbar_arrive_expect_tx(&full_bar[lane], kStageBytes);
                    if (!ex) load_x(lane, lane);
                    if (!ey) load_y(lane, lane);
                }
            }
            __syncwarp();
            {
                // the whole (converged) warp walks the ring and ONE elected lane issues: the operands of the bulk-tensor
                // instructions are then warp-uniform for the compiler (uniform registers); inside an `if (lane == 0)`
                // branch every issue went through an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall
                const bool leader = tc05::elect_one();
                int s = 0;
                uint32_t ph = 1;            // second pass over the ring waits for the first release of each slot
                for (int kb = first; kb < nkb; ++kb) {
                    tc05::mbar_wait(&empty_bar[s], ph ^ 1);
                    if (leader) {
                        tc05::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                        load_x(kb, s);
                        load_y(kb, s);
                    }
                    __syncwarp();
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                if (lane == 0) STIL_TRACE(2);
            }
        }
        if (clustered) {   // the two cluster barriers of the split-K reduction (every thread of the cluster takes part)
            tc05::cluster_arrive(); tc05::cluster_wait();
            tc05::cluster_arrive(); tc05::cluster_wait();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Converged warp, one elected lane issues: descriptors and TMEM addresses stay in uniform registers and the four
        // tcgen05.mma of a stage go out back to back.  Inside an `if (lane == 0)` branch the compiler wrapped every MMA in
        // an ELECT / 4 x R2UR.BROADCAST / BRA.U.ANY waterfall: ~100 cycles of dependent issue per 64 cycles of tensor work
        // (profiles/r2_bank_timeline.txt has the switch-off experiments that found it).
        {
            const bool leader = tc05::elect_one();
            const bool mn = MODE == GEMM_STORE && J.y_mn_major != 0;
            const bool xmn = MODE == GEMM_STORE && J.x_mn_major != 0;
            // bits 15 / 16: A / B given MN-major
            const uint32_t idesc = tc05::make_idesc_bf16_f32(kTileM, kTileN) | (mn ? (1u << 16) : 0u) | (xmn ? (1u << 15) : 0u);
            // K-major: 16 bf16 = 32 B inside the swizzle atom (+2 in the >>4 address field);
            // MN-major: 16 contraction rows = 2 KiB (+128)
            const uint32_t a_step = xmn ? 128u : 2u;
            const uint32_t b_step = mn ? 128u : 2u;
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                tc05::mbar_wait(&full_bar[s], ph);
                tc05::fence_after_sync();
                const uint32_t a_addr = tc05::smem_u32(tiles + s * kStageBytes);
                const uint32_t b_addr = a_addr + kTileM * kTileK * 2;
                const uint64_t a_desc = xmn ? tc05::make_mnmajor_sw128_desc(a_addr, 64 * kTileK * 2)
                                            : tc05::make_kmajor_sw128_desc(a_addr);
                const uint64_t b_desc = mn ? tc05::make_mnmajor_sw128_desc(b_addr, 64 * kTileK * 2)
                                           : tc05::make_kmajor_sw128_desc(b_addr);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)
                        tc05::mma_f16_ss(tmem_base, a_desc + a_step * k, b_desc + b_step * k, idesc, (kb | k) ? 1u : 0u);
                    // frees the smem stage when these MMAs retire — only if the producer will refill it: a tcgen05.commit
                    // costs ~0.2 us of (serialised) completion tracking (bank_sweep.cu experiments), and behind a queue of
                    // useless ones the tile's final commit reaches the epilogue late
                    if (kb + kStages < nkb) tc05::mma_commit(&empty_bar[s]);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
            if (leader) {
                if (nkb > 0) tc05::mma_commit(tmem_full_bar);
                else tc05::mbar_arrive(tmem_full_bar);   // empty slice of a split contraction: nothing to wait for
            }
            __syncwarp();
            if (lane == 0) STIL_TRACE(3);
        }
        if (clustered) {
            tc05::cluster_arrive(); tc05::cluster_wait();
            tc05::cluster_arrive(); tc05::cluster_wait();
        }
    } else {
        // ===================== epilogue: kEpiWarps warps, thread = accumulator row, kParts warps per TMEM lane
        // quarter (each owns kCPW 32-column chunks of the tile) =====================
        const int e = threadIdx.x - 64;      // 0..kEpiThreads-1
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int part = (warp - 2) >> 2;    // which kCPW*32 columns
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < J.M;
        const int ncols = min(kTileN, J.N - n0);
        const float alpha = J.alpha;
        // J.early_stats: the row statistics (and plain row/column scales) were final two kernels ago — merge them before
        // waiting for the predecessor, overlapping its execution
        float lse_x_early = 0.f;
        const bool pre_stats = MODE == GEMM_GRAD && J.early_stats && J.lse_x == nullptr && J.px_max != nullptr;
        if (pre_stats && row_ok) lse_x_early = merge_partials(J.px_max, J.px_sum, J.px_tiles, J.M, row);
        asm volatile("griddepcontrol.wait;" ::: "memory");               // predecessor's results visible from here on
        if (e == 0) {
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // next kernel may begin its prologue
            STIL_TRACE(1);
        }
        if (J.wait_flags) {
            // column scales gathered from peer ranks: ONE thread waits for the owners of this tile's columns, the barrier
            // passes the acquire on to the others, which then read through L2
            if (e == 0) wait_peer_rows(J.wait_flags, J.wait_seq, J.wait_rows_per_peer, n0, min(n0 + kTileN, J.N) - 1);
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        }
        if (e < kTileN) {
            const int col = n0 + e;
            float cs = 0.f, cl = 0.f;
            if (col < J.N) {
                cs = alpha * (J.sy ? __ldcg(J.sy + col) : 1.f);
                if (MODE == GEMM_GRAD) {
                    if (J.lse_y) cl = J.lse_ll_tag ? ll_read(J.lse_y, col, J.lse_ll_tag) : __ldcg(J.lse_y + col);
                    else if (J.py_max) cl = merge_partials(J.py_max, J.py_sum, J.py_tiles, J.N, col);
                }
            }
            if (MODE == GEMM_STORE) cl = (J.col_bias && col < J.N) ? __ldg(J.col_bias + col) : 0.f;   // Linear bias (f-3)
            col_scale[e] = cs;
            col_lse[e] = cl;
        }
        float rs = (row_ok && J.sx) ? J.sx[row] : 1.f;
        if (MODE == GEMM_STORE && J.sx_recip) rs = 1.0f / rs;
        float lse_x = 0.f, u = 0.f, d = 0.f, gs = 1.f;
        int tgt = -1;
        if (MODE == GEMM_GRAD && row_ok) {
            lse_x = J.lse_x ? (J.lse_ll_tag ? ll_read(J.lse_x, row, J.lse_ll_tag) : J.lse_x[row])
                    : pre_stats ? lse_x_early : merge_partials(J.px_max, J.px_sum, J.px_tiles, J.M, row);
            tgt = J.tgt_vec ? J.tgt_vec[row] : row + J.tgt_offset;
            if (J.w_z) {
                // prototype CE coefficient from the picked logit (utils/prototype_loss.py:28,37-39)
                const float p = expf(J.w_z[(long long)row * J.w_ldz + tgt] - lse_x);
                u = (J.w_conf[row] ? J.w_coef : 0.f) * p / (p + 1e-7f);
                d = u;
            } else {
                u = J.u_vec ? J.u_vec[row] : J.u_scalar;
                d = J.u_vec ? u : J.d_scalar;
            }
            gs = J.gscale ? *J.gscale : 1.f;
            if (J.g_row_scale) gs *= rs;
        }
        const float v = (MODE == GEMM_GRAD && (J.lse_y || J.py_max)) ? J.v_scalar : 0.f;
        const bool fused_fin = MODE == GEMM_STORE && J.fin_dx != nullptr;
        const float fsx = (fused_fin && row_ok && J.fin_sx) ? J.fin_sx[row] : 0.f;
        // this thread's columns of x for the normalise-backward, fetched while the MMAs run
        float xrow[kCPW][32];
        if (MODE == GEMM_STORE && fused_fin && J.fin_sx) {
#pragma unroll
            for (int cc = 0; cc < kCPW; ++cc) {
                const int c = part * kCPW + cc;
                const int nv = max(0, min(32, ncols - c * 32));
                if (row_ok && nv > 0)
                    load32_as_float(J.fin_x, J.fin_x_dtype, (long long)row * J.fin_ldx + n0 + c * 32, nv, xrow[cc]);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xrow[cc][j] = 0.f;
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // epilogue-only named barrier: col_scale / col_lse ready

        if (e == 0) STIL_TRACE(4);   // epilogue prologue (scales, merges, coefficients) done
        tc05::mbar_wait(tmem_full_bar, 0);
        tc05::fence_after_sync();
        if (e == 0) STIL_TRACE(5);   // accumulator ready

        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int c_begin = part * kCPW, c_end = c_begin + kCPW;   // this warp's 32-column chunks

        bool cluster_follower = false;
        if (MODE == GEMM_STORE && clustered) {
            // Partial tile [128 x 128] fp32 in this CTA's (now idle) operand ring, stored COLUMN-major ([col][row]): the 32
            // threads of a warp (consecutive rows) then touch 32 consecutive words, on the writing side (conflict-free shared
            // stores) and, more importantly, on the reading side — distributed shared memory behaves like global memory: a
            // row-major tile read thread = row (512-byte stride between lanes) moved 17 GB/s and took 11 us per leader
            // (profiles/r2_cluster_splitk.txt).  The raw accumulator travels: scales are linear, applied once by the leader.
            const int r_in = q * 32 + lane;
            if (ks != 0) {
                cluster_follower = true;
                float* mine = reinterpret_cast<float*>(tiles);
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t acc[32];
                    if (nkb > 0) {
                        tc05::tmem_ld_32x32b_x32(taddr + c * 32, acc);
                        tc05::tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = 0u;     // empty slice: the accumulator was never written
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) mine[(c * 32 + j) * kTileM + r_in] = __uint_as_float(acc[j]);
                }
                tc05::cluster_arrive(); tc05::cluster_wait();     // partial published
                tc05::cluster_arrive(); tc05::cluster_wait();     // the leader has read it: the CTA may retire
            } else {
                tc05::cluster_arrive(); tc05::cluster_wait();     // every follower's partial is in its shared memory
                const uint32_t local = tc05::smem_u32(reinterpret_cast<float*>(tiles) + r_in);
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t acc[32];
                    if (nkb > 0) {
                        tc05::tmem_ld_32x32b_x32(taddr + c * 32, acc);
                        tc05::tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = 0u;
                    }
                    for (int peer = 1; peer < ksplit; ++peer) {       // fixed order: deterministic
                        const uint32_t remote = tc05::map_to_cta(local, (uint32_t)peer) + (uint32_t)(c * 32 * kTileM * 4);
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = tc05::ld_dsmem_f32(remote + (uint32_t)(j * kTileM * 4));
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + v[j]);
                    }
                    tc05::tmem_st_32x32b_x32(taddr + c * 32, acc);   // the summed tile replaces this CTA's accumulator
                }
                tc05::tmem_st_wait();
                tc05::cluster_arrive();                               // followers may retire (waited for at the very end)
            }
        }

        // forward of Linear + F.normalize (STiLModel.py:182-192): pass 1 = squared norm of the biased row (tile spans N)
        float fwd_inv = 1.f;
        if (MODE == GEMM_STORE && J.fwd_norm && !cluster_follower) {
            float part_ss = 0.f;
#pragma unroll
            for (int cc = 0; cc < kCPW; ++cc) {
                const int c = c_begin + cc;
                if (c * 32 < ncols) {   // warp-uniform
                    uint32_t acc[32];
                    tc05::tmem_ld_32x32b_x32(taddr + c * 32, acc);
                    tc05::tmem_ld_wait();
                    const int nv = min(32, ncols - c * 32);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(acc[j]) * rs * col_scale[c * 32 + j] + col_lse[c * 32 + j];
                        if (j < nv) part_ss += v * v;
                    }
                }
            }
            dot_part[part * kTileM + q * 32 + lane] = part_ss;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            float ss = 0.f;
#pragma unroll
            for (int pp = 0; pp < kParts; ++pp) ss += dot_part[pp * kTileM + q * 32 + lane];
            fwd_inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);       // F.normalize eps
            if (part == 0 && row_ok && J.fwd_inv_norm) J.fwd_inv_norm[row] = fwd_inv;
        }

        float fin_dot = 0.f;
        if (MODE == GEMM_STORE && fused_fin && J.fin_sx && !cluster_follower) {
            // pass 1 of the normalise-backward: <xh, g> over the whole row (the tile spans all of N)
            float part_dot = 0.f;
#pragma unroll
            for (int cc = 0; cc < kCPW; ++cc) {
                const int c = c_begin + cc;
                if (c * 32 < ncols) {   // warp-uniform
                    uint32_t acc[32];
                    tc05::tmem_ld_32x32b_x32(taddr + c * 32, acc);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) part_dot += xrow[cc][j] * __uint_as_float(acc[j]);
                }
            }
            dot_part[part * kTileM + q * 32 + lane] = part_dot * fsx * alpha * rs;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
#pragma unroll
            for (int pp = 0; pp < kParts; ++pp) fin_dot += dot_part[pp * kTileM + q * 32 + lane];
        }

#pragma unroll 1
        for (int c = c_begin; c < (cluster_follower ? c_begin : c_end); ++c) {
            float run_max = -INFINITY, run_sum = 0.f;
            if (c * 32 >= ncols) {   // warp-uniform: an empty chunk contributes the neutral statistics (-inf, 0)
                if (MODE == GEMM_STATS && row_ok) {
                    J.part_max[(long long)(tn * 4 + c) * J.M + row] = run_max;
                    J.part_sum[(long long)(tn * 4 + c) * J.M + row] = run_sum;
                }
                continue;
            }
            uint32_t acc[32];
            tc05::tmem_ld_32x32b_x32(taddr + c * 32, acc);
            tc05::tmem_ld_wait();
            const int nv = min(32, ncols - c * 32);
            float l[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) l[j] = __uint_as_float(acc[j]) * rs * col_scale[c * 32 + j];

            if (MODE == GEMM_STORE && (J.col_bias || J.fwd_norm)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) l[j] += col_lse[c * 32 + j];
                if (J.fwd_norm) {
                    if (J.fwd_raw && row_ok) {
                        float* rw = J.fwd_raw + (long long)row * J.fwd_ld_raw + n0 + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nv) rw[j] = l[j];
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) l[j] *= fwd_inv;
                }
            }

            if (MODE == GEMM_STATS && J.cpart_sum != nullptr) {
                // single pass over the symmetric problem: e = exp(L - shift) once per element, row sum in the thread,
                // column sums over this warp's 32 rows by a shuffle transpose-reduce (31 shuffles per 32 x 32 block)
                const float shift_l2 = J.sym_shift * kLog2e;
                float ev[32];
                float rsum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    ev[j] = (row_ok && j < nv) ? fast_exp2(fmaf(l[j], kLog2e, -shift_l2)) : 0.f;
                    rsum += ev[j];
                }
                if (row_ok) {
                    J.part_max[(long long)(tn * 4 + c) * J.M + row] = J.sym_shift;
                    J.part_sum[(long long)(tn * 4 + c) * J.M + row] = rsum;
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    const bool up = (lane & o) != 0;
#pragma unroll
                    for (int i = 0; i < o; ++i) {
                        const float send = up ? ev[i] : ev[i + o];
                        const float keep = up ? ev[i + o] : ev[i];
                        ev[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
                // lane j now holds the sum of column c*32 + j over the 32 rows of this TMEM lane quarter
                if (lane < nv) {
                    const long long slot = (long long)(tm * 4 + q) * J.N + n0 + c * 32 + lane;
                    J.cpart_max[slot] = J.sym_shift;
                    J.cpart_sum[slot] = ev[0];
                }
            } else if (MODE == GEMM_STATS) {
                float cmax = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) cmax = fmaxf(cmax, l[j]);
                const float new_max = fmaxf(run_max, cmax);
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) s += fast_exp2((l[j] - new_max) * kLog2e);
                run_sum = run_sum * fast_exp2((run_max - new_max) * kLog2e) + s;
                run_max = new_max;
                if (row_ok) {   // one (max, sum) partial per 32-column chunk
                    J.part_max[(long long)(tn * 4 + c) * J.M + row] = run_max;
                    J.part_sum[(long long)(tn * 4 + c) * J.M + row] = run_sum;
                }
            }
            if (MODE == GEMM_STORE && fused_fin) {
                if (J.fin_sx && row_ok) {
                    if (kCPW == 1 || c == c_begin) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) l[j] = fsx * (l[j] - fsx * xrow[0][j] * fin_dot);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) l[j] = fsx * (l[j] - fsx * xrow[kCPW - 1][j] * fin_dot);
                    }
                }
                if (J.out_tma) {
                    // fp32 dX: swizzled box in the idle operand ring, one bulk tensor store per box after the loop.  Thread =
                    // row stores straight to global are 32 scattered 16-byte pieces per instruction: ~2 us of LSU time per
                    // tile, on the last kernel of the step's critical chain (profiles/r2_gemm_timeline_c2.txt: epilogue 5 us)
                    const int r_in = q * 32 + lane;
                    uint8_t* box = tiles + c * (kTileM * 128) + r_in * 128;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4)
                        *reinterpret_cast<float4*>(box + ((k4 ^ (r_in & 7)) * 16)) = make_float4(l[4 * k4], l[4 * k4 + 1], l[4 * k4 + 2], l[4 * k4 + 3]);
                    continue;
                }
                if (row_ok) store32_from_float(J.fin_dx, J.fin_dx_dtype, (long long)row * J.fin_ld_dx + n0 + c * 32, nv, l);
            } else if (MODE == GEMM_STORE && ksplit > 1 && J.slice_stride == 0 && !clustered) {
                // split contraction: add this CTA's partial tile (an empty slice adds nothing)
                if (row_ok && nkb > 0) {
                    float* dst = J.out + (long long)row * J.ld_out + n0 + c * 32;
                    if (nv == 32 && (J.ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(J.out) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)   // one 16-byte reduction per 4 columns
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(l[j]),
                                         "f"(l[j + 1]), "f"(l[j + 2]), "f"(l[j + 3])
                                         : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nv) atomicAdd(dst + j, l[j]);
                    }
                }
            } else if ((MODE == GEMM_STATS || MODE == GEMM_STORE) && J.out) {
                if (MODE == GEMM_STORE && J.post_op) {
                    // graph epilogues of the CoMatch block (comatch_model.py:299-300, 309-311)
#pragma unroll
                    for (int j = 0; j < 32; ++j) l[j] = store_post(l[j], J.post_op, row, n0 + c * 32 + j);
                }
                if (J.out_tma) {
                    // fp32 tile -> 128-byte-swizzled [128 rows x 32 floats] boxes in the (idle) operand ring -> ONE bulk tensor
                    // store per box after the loop: the per-lane path below needed 32 store instructions per 32 x 32 block and
                    // made the epilogue the long pole of the bank logits (202 us for 2 x 448 x 65536, profiles/r2_bank_timeline.txt)
                    const bool empty = MODE == GEMM_STORE && ksplit > 1 && nkb == 0 && !clustered;
                    const int r_in = q * 32 + lane;
                    uint8_t* box = tiles + c * (kTileM * 128) + r_in * 128;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const int off = (k4 ^ (r_in & 7)) * 16;
                        *reinterpret_cast<float4*>(box + off) = empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                                                                      : make_float4(l[4 * k4], l[4 * k4 + 1], l[4 * k4 + 2], l[4 * k4 + 3]);
                    }
                    continue;
                }
                // coalesced store through a per-warp staging tile (the operand ring is idle once the accumulator is
                // complete): registers (thread = row) -> smem [32][33] -> 32 rows of up to 128 contiguous bytes; no
                // alignment or leading-dimension requirement on `out`
                float* stg = reinterpret_cast<float*>(tiles) + (warp - 2) * kStageFloats;
                // split contraction with per-slice outputs (deterministic: the consumer adds the slices in order);
                // an empty slice stores zeros (its accumulator was never written)
                const bool empty_slice = MODE == GEMM_STORE && ksplit > 1 && nkb == 0 && !clustered;
#pragma unroll
                for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = empty_slice ? 0.f : l[j];
                __syncwarp();
                {
                    const int rbase = m0 + q * 32;
                    float* obase = J.out + ((MODE == GEMM_STORE && !clustered) ? (long long)ks * J.slice_stride : 0ll) +
                                   (long long)rbase * J.ld_out + n0 + c * 32 + lane;
                    const int rmax = min(32, J.M - rbase);
                    if (lane < nv) {
#pragma unroll 8
                        for (int r = 0; r < rmax; ++r) obase[(long long)r * J.ld_out] = stg[r * 33 + lane];
                    }
                }
                __syncwarp();
            }
            if (MODE == GEMM_GRAD) {
                const bool want_lo = J.g_nseg > 1;
                // G leaves through shared memory and TMA: thread = row stores straight to global would issue 32 scattered
                // 16-byte sectors per instruction and make the LSU the long pole of the epilogue (2 us per tile).  Each
                // thread writes its 64 bytes into a 128-byte-swizzled [128 rows x 64 columns] box (conflict-free), one
                // thread then issues a bulk tensor store per box and segment.  Rows >= M and columns >= ld_g are clipped
                // by the tensor map; columns in [N, ld_g) hold finite garbage the dX GEMM never reads.
                uint32_t hi_pk[16], lo_pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float g2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int jj = j + h;
                        float g = u * fast_exp2((l[jj] - lse_x) * kLog2e);
                        if (v != 0.f) g += v * fast_exp2((l[jj] - col_lse[c * 32 + jj]) * kLog2e);
                        if (n0 + c * 32 + jj == tgt) g -= d;
                        g2[h] = g * col_scale[c * 32 + jj] * gs;
                    }
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(g2[0], g2[1]);
                    const float2 hf = __bfloat1622float2(hh);
                    const __nv_bfloat162 ll = __floats2bfloat162_rn(g2[0] - hf.x, g2[1] - hf.y);
                    hi_pk[j / 2] = *reinterpret_cast<const uint32_t*>(&hh);
                    lo_pk[j / 2] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                const int r_in = q * 32 + lane;                 // row inside the tile
                const int box = c >> 1, k0 = (c & 1) * 4;       // 64-column box, first 16-byte chunk of this warp's 32 columns
                uint8_t* hi_box = tiles + box * (kTileM * 128) + r_in * 128;
                uint8_t* lo_box = hi_box + 2 * (kTileM * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int off = ((k0 + k) ^ (r_in & 7)) * 16;   // SWIZZLE_128B: chunk index xor (row mod 8)
                    *reinterpret_cast<uint4*>(hi_box + off) = make_uint4(hi_pk[4 * k], hi_pk[4 * k + 1], hi_pk[4 * k + 2], hi_pk[4 * k + 3]);
                    if (want_lo)
                        *reinterpret_cast<uint4*>(lo_box + off) = make_uint4(lo_pk[4 * k], lo_pk[4 * k + 1], lo_pk[4 * k + 2], lo_pk[4 * k + 3]);
                }
            }
        }
        if ((MODE == GEMM_STATS || MODE == GEMM_STORE) && J.out_tma && !cluster_follower &&
            (J.out != nullptr || (MODE == GEMM_STORE && J.fin_dx != nullptr))) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (TMA)
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (e == 0) {
                const int slice = (MODE == GEMM_STORE && !clustered && J.slice_stride > 0) ? ks : 0;
                for (int c = 0; c < 4; ++c)
                    if (c * 32 < ncols) tc05::tma_store_3d(&J.tmg, tiles + c * (kTileM * 128), n0 + c * 32, m0, slice);
                tc05::tma_store_commit_and_wait();
            }
        }
        if (MODE == GEMM_GRAD) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (TMA)
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (e == 0) {
                for (int sgm = 0; sgm < J.g_nseg; ++sgm)
                    for (int box = 0; box < 2; ++box)
                        if (box * 64 < ncols)
                            tc05::tma_store_3d(&J.tmg, tiles + (sgm * 2 + box) * (kTileM * 128), n0 + box * 64, m0, sgm);
                tc05::tma_store_commit_and_wait();
            }
        }
    }

    // ---- teardown: all TMEM reads done before dealloc
    if (MODE == GEMM_STORE && clustered && warp >= 2 && ks == 0) tc05::cluster_wait();   // second cluster barrier (leader side)
    tc05::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) STIL_TRACE(6);
    if (warp == 1) {
        tc05::fence_after_sync();
        tc05::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// =====================================================================================================================
// Fused backward: one CTA owns a 128-row block of dX and walks over its column tiles.  Per tile: TMA brings the Y rows
// (X is loaded once), MMA 1 recomputes the logits into a TMEM buffer, the epilogue warps turn them into G = dLoss/dLogits
// and write G as a swizzled K-major operand into shared memory, MMA 2 accumulates dX += G · Y with the SAME Y tile read
// as an MN-major operand.  G never leaves the SM, and one launch replaces the GRAD + STORE pair (one kernel boundary and
// one global round trip of G less on the critical chain of the step).  With two Y buffers the next tile's recompute runs
// under this tile's epilogue (two logits buffers in TMEM).  Column tiles are dealt round-robin to `nsplit` CTAs per row
// block for long rows (global batch); their partial dX tiles go to per-slice outputs that grad_finish_kernel adds.
// =====================================================================================================================
constexpr int kBwdEpiWarps = 16;
constexpr int kBwdEpiThreads = 32 * kBwdEpiWarps;
constexpr int kBwdThreads = 64 + kBwdEpiThreads;
constexpr int kBoxBytes = kTileM * 128;          // one [128 rows x 64 bf16] swizzled box
constexpr int kBwdMiscBytes = 2 * 2 * kTileN * 4 /*col scale, lse x2 buffers*/ + 4 * kTileM * 4 /*dot partials*/ + 256;
constexpr uint32_t kBwdTmemCols = 512;           // logits x2 (256 columns) + dX (<= 128 columns)

__global__ void __launch_bounds__(kBwdThreads, 1) gemm_bwd_kernel(const __grid_constant__ GemmLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc05::smem_u32(smem_raw);
    uint8_t* base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int jid = 0;
#pragma unroll
    for (int j = 1; j < kMaxGemmJobs; ++j)
        if (j < L.njobs && (int)blockIdx.x >= L.job[j].tile_begin) jid = j;
    const GemmJob& J = L.job[jid];
    unsigned long long* trace = L.trace ? L.trace + (size_t)(L.trace_id % kTraceLaunches) * kTraceCtas * kTraceSlots : nullptr;
    if (threadIdx.x == 0) { STIL_TRACE(0); if (trace && blockIdx.x < kTraceCtas) trace[blockIdx.x * kTraceSlots + 7] = GEMM_BWD; }
    const int t_cta = blockIdx.x - J.tile_begin;
    const int nsplit = J.nsplit > 1 ? J.nsplit : 1;
    const int split = t_cta % nsplit, tm = t_cta / nsplit;
    const int m0 = tm * kTileM;
    const int ntile = (J.tiles_n - split + nsplit - 1) / nsplit;     // this CTA's column tiles: split, split+nsplit, ...
    const int nbx = J.D / 64;                                        // 64-column boxes along the embedding dimension
    const int nbuf = J.bw_nbuf;

    uint8_t* xs = base;                                              // [bw_nx][nbx] boxes
    uint8_t* ys = base + J.bw_y_off;                                 // [nbuf][bw_ny][nbx] boxes
    uint8_t* gsm = base + J.bw_g_off;                                // [g_nseg][2] boxes (K-major A operand of MMA 2)
    float* col_scale = reinterpret_cast<float*>(base + J.bw_misc_off);   // [2][128]
    float* col_lse = col_scale + 2 * kTileN;                             // [2][128]
    float* dot_part = col_lse + 2 * kTileN;                              // [4][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(dot_part + 4 * kTileM);
    uint64_t* x_full = bars;
    uint64_t* y_full = bars + 1;      // [2]
    uint64_t* y_empty = bars + 3;     // [2]
    uint64_t* s_full = bars + 5;      // [2]
    uint64_t* s_empty = bars + 7;     // [2]
    uint64_t* g_full = bars + 9;
    uint64_t* g_empty = bars + 10;
    uint64_t* dx_full = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    if (warp == 0 && lane == 0) {
        tc05::tma_prefetch_desc(&J.tmx);
        tc05::tma_prefetch_desc(&J.tmy);
    }
    if (warp == 1) {
        if (lane == 0) {
            tc05::mbar_init(x_full, 1);
            for (int i = 0; i < 2; ++i) {
                tc05::mbar_init(&y_full[i], 1);
                tc05::mbar_init(&y_empty[i], 1);
                tc05::mbar_init(&s_full[i], 1);
                tc05::mbar_init(&s_empty[i], kBwdEpiWarps);
            }
            tc05::mbar_init(g_full, kBwdEpiWarps);
            tc05::mbar_init(g_empty, 1);
            tc05::mbar_init(dx_full, 1);
            tc05::fence_mbar_init();
        }
        __syncwarp();
        tc05::tmem_alloc(tmem_slot, kBwdTmemCols);
        tc05::tmem_relinquish();
    }
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t y_bytes = (uint32_t)(J.bw_ny * nbx * kBoxBytes);

    if (warp == 0) {
        // ===================== TMA producer (converged warp, one elected lane issues) =====================
        const bool leader = tc05::elect_one();
        {
            auto load_x = [&]() {
                if (!leader) return;
                tc05::mbar_arrive_expect_tx(x_full, (uint32_t)(J.bw_nx * nbx * kBoxBytes));
                for (int sg = 0; sg < J.bw_nx; ++sg)
                    for (int b = 0; b < nbx; ++b)
                        tc05::tma_load_3d(xs + (sg * nbx + b) * kBoxBytes, &J.tmx, x_full, b * 64, m0, sg);
            };
            auto load_y = [&](int i) {
                if (!leader) return;
                const int yb = i % nbuf, n0 = (split + i * nsplit) * kTileN;
                tc05::mbar_arrive_expect_tx(&y_full[yb], y_bytes);
                for (int sg = 0; sg < J.bw_ny; ++sg)
                    for (int b = 0; b < nbx; ++b)
                        tc05::tma_load_3d(ys + yb * J.bw_y_stride + (sg * nbx + b) * kBoxBytes, &J.tmy, &y_full[yb], b * 64, n0,
                                          sg);
            };
            const int first = min(ntile, nbuf);
            const bool early = J.early_x && J.early_y;   // operands not written by the stream predecessor: before the wait
            if (!early) asm volatile("griddepcontrol.wait;" ::: "memory");
            load_x();
            for (int i = 0; i < first; ++i) load_y(i);
            for (int i = first; i < ntile; ++i) {
                const int yb = i % nbuf;
                tc05::mbar_wait(&y_empty[yb], ((i / nbuf) & 1) ^ 1);
                load_y(i);
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, one elected lane issues) =====================
        const bool leader = tc05::elect_one();
        {
            const uint32_t idesc1 = tc05::make_idesc_bf16_f32(kTileM, kTileN);
            const uint32_t idesc2 = tc05::make_idesc_bf16_f32(kTileM, J.D) | (1u << 16);   // B = Y tile, MN-major
            const uint32_t xs_a = tc05::smem_u32(xs), ys_a = tc05::smem_u32(ys), g_a = tc05::smem_u32(gsm);
            // a completed phase stays complete: what ANY lane has seen holds for the warp (keeps the loop converged)
            auto mma1_ready = [&](int i) {
                return __any_sync(0xffffffffu, tc05::mbar_test(&y_full[i % nbuf], (i / nbuf) & 1)) &&
                       __any_sync(0xffffffffu, tc05::mbar_test(&s_empty[i & 1], ((i >> 1) & 1) ^ 1));
            };
            auto mma1 = [&](int i) {     // logits of tile i into TMEM buffer i & 1 (caller checked mma1_ready)
                const int yb = i % nbuf, sb = i & 1;
                tc05::fence_after_sync();
                if (!leader) return;
                uint32_t acc = 0;
                for (int p = 0; p < J.npair; ++p)
                    for (int b = 0; b < nbx; ++b) {
                        const uint64_t a_desc = tc05::make_kmajor_sw128_desc(xs_a + (J.xseg[p] * nbx + b) * kBoxBytes);
                        const uint64_t b_desc =
                            tc05::make_kmajor_sw128_desc(ys_a + yb * J.bw_y_stride + (J.yseg[p] * nbx + b) * kBoxBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            tc05::mma_f16_ss(tmem_base + sb * kTileN, a_desc + 2 * k, b_desc + 2 * k, idesc1, acc);
                            acc = 1;
                        }
                    }
                tc05::mma_commit(&s_full[sb]);
            };
            auto mma2 = [&](int i) {     // dX += G(tile i) · Y(tile i): contraction over the tile's 128 columns
                const int yb = i % nbuf;   // (caller saw g_full complete its phase i)
                tc05::fence_after_sync();
                if (!leader) return;
                for (int p = 0; p < J.npair2; ++p) {
                    const uint64_t b_base = tc05::make_mnmajor_sw128_desc(
                        ys_a + yb * J.bw_y_stride + (J.yseg2[p] * nbx) * kBoxBytes, kBoxBytes);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t a_desc = tc05::make_kmajor_sw128_desc(g_a + (J.gseg2[p] * 2 + kb) * kBoxBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc05::mma_f16_ss(tmem_base + 2 * kTileN, a_desc + 2 * k, b_base + 128u * (kb * 4 + k), idesc2,
                                             (i | p | kb | k) ? 1u : 0u);
                    }
                }
                // (only the releases somebody will wait for: every commit is ~0.2 us of serialised completion tracking)
                if (i + 1 < ntile) tc05::mma_commit(g_empty);                // G buffer free for the next tile
                if (i + nbuf < ntile) tc05::mma_commit(&y_empty[yb]);       // Y buffer free for the tile after next
                if (i == ntile - 1) tc05::mma_commit(dx_full);
            };
            tc05::mbar_wait(x_full, 0);
            // One thread serves both products: it issues whichever is ready (never parks on the Y tile of a later
            // recompute while the current tile's G is waiting for its dX product).  A recompute may run at most two
            // tiles ahead (two logits buffers).
            int n1 = 0, n2 = 0;
            unsigned long long spins = 0;
            while (n2 < ntile) {
                if (n1 < ntile && n1 < n2 + 2 && mma1_ready(n1)) { mma1(n1++); __syncwarp(); spins = 0; continue; }
                if (n2 < n1 && __any_sync(0xffffffffu, tc05::mbar_test(g_full, n2 & 1))) { mma2(n2++); __syncwarp(); spins = 0; continue; }
                if (++spins > (1ull << 28)) __trap();
            }
            if (ntile == 0 && leader) tc05::mbar_arrive(dx_full);
            __syncwarp();
        }
    } else {
        // ===================== epilogue: 16 warps, thread = row, one 32-column chunk per warp =====================
        const int e = threadIdx.x - 64;
        const int q = warp & 3;
        const int c = (warp - 2) >> 2;                 // chunk 0..3 of the 128-column tile
        const int r_in = q * 32 + lane;
        const int row = m0 + r_in;
        const bool row_ok = row < J.M;
        const float alpha = J.alpha;
        float lse_x_early = 0.f;
        const bool pre_stats = J.early_stats && J.lse_x == nullptr && J.px_max != nullptr;
        if (pre_stats && row_ok) lse_x_early = merge_partials(J.px_max, J.px_sum, J.px_tiles, J.M, row);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (e == 0) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); STIL_TRACE(1); }
        // ---- per-row quantities (as in gemm_tc05_kernel<GEMM_GRAD>)
        const float rs = (row_ok && J.sx) ? J.sx[row] : 1.f;
        float lse_x = 0.f, u = 0.f, d = 0.f, gs = 1.f;
        int tgt = -1;
        if (row_ok) {
            lse_x = J.lse_x ? (J.lse_ll_tag ? ll_read(J.lse_x, row, J.lse_ll_tag) : J.lse_x[row])
                    : pre_stats ? lse_x_early : merge_partials(J.px_max, J.px_sum, J.px_tiles, J.M, row);
            tgt = J.tgt_vec ? J.tgt_vec[row] : row + J.tgt_offset;
            if (J.w_z) {
                const float p = expf(J.w_z[(long long)row * J.w_ldz + tgt] - lse_x);
                u = (J.w_conf[row] ? J.w_coef : 0.f) * p / (p + 1e-7f);
                d = u;
            } else {
                u = J.u_vec ? J.u_vec[row] : J.u_scalar;
                d = J.u_vec ? u : J.d_scalar;
            }
            gs = J.gscale ? *J.gscale : 1.f;
        }
        const float v = (J.lse_y || J.py_max) ? J.v_scalar : 0.f;
        const bool want_lo = J.g_nseg > 1;
        // normalise-backward inputs of the final epilogue (this thread's 32 columns of x)
        const bool fused_fin = J.fin_dx != nullptr;
        const float fsx = (fused_fin && row_ok && J.fin_sx) ? J.fin_sx[row] : 0.f;
        const int dcols = max(0, min(32, J.D - c * 32));
        float xrow[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) xrow[j] = 0.f;
        if (fused_fin && J.fin_sx && row_ok && dcols > 0)
            load32_as_float(J.fin_x, J.fin_x_dtype, (long long)row * J.fin_ldx + c * 32, dcols, xrow);
        // column scale / LSE of a tile, fetched one tile ahead by the first 128 epilogue threads
        auto col_vals = [&](int i, float& cs, float& cl) {
            cs = 0.f; cl = 0.f;
            const int col = (split + i * nsplit) * kTileN + e;
            if (i < ntile && col < J.N) {
                cs = alpha * (J.sy ? __ldcg(J.sy + col) : 1.f);
                if (J.lse_y) cl = J.lse_ll_tag ? ll_read(J.lse_y, col, J.lse_ll_tag) : __ldcg(J.lse_y + col);
                else if (J.py_max) cl = merge_partials(J.py_max, J.py_sum, J.py_tiles, J.N, col);
            }
        };
        float ncs = 0.f, ncl = 0.f;
        if (e < kTileN) col_vals(0, ncs, ncl);
        if (e == 0) STIL_TRACE(2);      // row / column quantities loaded

        for (int i = 0; i < ntile; ++i) {
            const int sb = i & 1, n0 = (split + i * nsplit) * kTileN;
            if (e < kTileN) {
                col_scale[sb * kTileN + e] = ncs;
                col_lse[sb * kTileN + e] = ncl;
                col_vals(i + 1, ncs, ncl);            // next tile's values travel while this tile is processed
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kBwdEpiThreads) : "memory");
            tc05::mbar_wait(&s_full[sb], (i >> 1) & 1);
            tc05::fence_after_sync();
            uint32_t acc[32];
            tc05::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sb * kTileN + c * 32, acc);
            tc05::tmem_ld_wait();
            tc05::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(&s_empty[sb]);      // this warp is done with the logits buffer
            const float* csc = col_scale + sb * kTileN + c * 32;
            const float* cls_ = col_lse + sb * kTileN + c * 32;
            uint32_t hi_pk[16], lo_pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float g2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int jj = j + h;
                    const float l = __uint_as_float(acc[jj]) * rs * csc[jj];
                    float g = u * fast_exp2((l - lse_x) * kLog2e);
                    if (v != 0.f) g += v * fast_exp2((l - cls_[jj]) * kLog2e);
                    if (n0 + c * 32 + jj == tgt) g -= d;
                    g2[h] = row_ok ? g * csc[jj] * gs : 0.f;
                }
                const __nv_bfloat162 hh = __floats2bfloat162_rn(g2[0], g2[1]);
                const float2 hf = __bfloat1622float2(hh);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(g2[0] - hf.x, g2[1] - hf.y);
                hi_pk[j / 2] = *reinterpret_cast<const uint32_t*>(&hh);
                lo_pk[j / 2] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            if (i > 0) tc05::mbar_wait(g_empty, (i - 1) & 1);    // MMA 2 of the previous tile has read G
            {
                const int box = c >> 1, k0 = (c & 1) * 4;
                uint8_t* hi_box = gsm + box * kBoxBytes + r_in * 128;
                uint8_t* lo_box = hi_box + 2 * kBoxBytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int off = ((k0 + k) ^ (r_in & 7)) * 16;
                    *reinterpret_cast<uint4*>(hi_box + off) = make_uint4(hi_pk[4 * k], hi_pk[4 * k + 1], hi_pk[4 * k + 2], hi_pk[4 * k + 3]);
                    if (want_lo)
                        *reinterpret_cast<uint4*>(lo_box + off) = make_uint4(lo_pk[4 * k], lo_pk[4 * k + 1], lo_pk[4 * k + 2], lo_pk[4 * k + 3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> MMA (async proxy) reads
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(g_full);
            if (e == 0 && i == 0) STIL_TRACE(3);   // first G tile in shared memory
            if (e == 0 && i == ntile - 1) STIL_TRACE(4);   // last G tile in shared memory
        }

        // ---- final epilogue: dX tile out of TMEM
        tc05::mbar_wait(dx_full, 0);
        tc05::fence_after_sync();
        if (e == 0) STIL_TRACE(5);      // dX accumulator complete
        float l[32];
        if (dcols > 0 && ntile > 0) {
            uint32_t acc[32];
            tc05::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * kTileN + c * 32, acc);
            tc05::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) l[j] = __uint_as_float(acc[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) l[j] = 0.f;
        }
        if (fused_fin) {
            if (J.fin_sx) {
                float part_dot = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) part_dot += xrow[j] * l[j];
                dot_part[c * kTileM + r_in] = part_dot * fsx;
                asm volatile("bar.sync 1, %0;" ::"n"(kBwdEpiThreads) : "memory");
                const float fin_dot = dot_part[r_in] + dot_part[kTileM + r_in] + dot_part[2 * kTileM + r_in] + dot_part[3 * kTileM + r_in];
#pragma unroll
                for (int j = 0; j < 32; ++j) l[j] = fsx * (l[j] - fsx * xrow[j] * fin_dot);
            }
            if (row_ok && dcols > 0) store32_from_float(J.fin_dx, J.fin_dx_dtype, (long long)row * J.fin_ld_dx + c * 32, dcols, l);
        } else if (row_ok && dcols > 0) {
            // partial (or plain fp32) dX: slice `split` of the output, added by grad_finish_kernel in index order
            float* dst = J.out + (long long)split * J.slice_stride + (long long)row * J.ld_out + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j + 3 < dcols) *reinterpret_cast<float4*>(dst + j) = make_float4(l[j], l[j + 1], l[j + 2], l[j + 3]);
        }
    }

    tc05::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) STIL_TRACE(6);
    if (warp == 1) {
        tc05::fence_after_sync();
        tc05::tmem_dealloc(tmem_base, kBwdTmemCols);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

}  // namespace

int make_operand_map(CUtensorMap* tm, const void* base, int64_t inner, int64_t rows, int64_t nseg,
                     int64_t row_stride, int64_t seg_stride, int box_rows, int box_segs) {
    EncodeTiledFn enc = get_encode_fn();
    STIL_REQUIRE(enc != nullptr, STIL_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    STIL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, STIL_E_ALIGN, "operand base %p not 16-byte aligned", base);
    STIL_REQUIRE((row_stride * 2) % 16 == 0 && (seg_stride * 2) % 16 == 0, STIL_E_ALIGN,
                 "operand strides (%lld, %lld elements) must be multiples of 8 bf16", (long long)row_stride,
                 (long long)seg_stride);
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nseg};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)seg_stride * 2};
    cuuint32_t box[3] = {kTileK, (cuuint32_t)box_rows, (cuuint32_t)box_segs};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    STIL_REQUIRE(r == CUDA_SUCCESS, STIL_E_CUDA,
                 "cuTensorMapEncodeTiled failed (%d) for [%lld x %lld x %lld], strides %lld/%lld", (int)r,
                 (long long)inner, (long long)rows, (long long)nseg, (long long)row_stride, (long long)seg_stride);
    return STIL_OK;
}

// fp32 output map [cols, rows, slices]: 32 x 128 x 1 boxes (128 bytes wide), 128-byte swizzle; false = `out` not eligible
bool make_out_map(CUtensorMap* tm, float* base, int64_t cols, int64_t rows, int64_t ld, int64_t slices, int64_t slice_stride) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0 || (slices > 1 && (slice_stride & 3) != 0)) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)std::max<int64_t>(slices, 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(slices > 1 ? slice_stride : ld * rows) * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)kTileM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void gemm_job_tiles(GemmLaunch& L) {
    int begin = 0;
    for (int j = 0; j < L.njobs; ++j) {
        GemmJob& J = L.job[j];
        J.tiles_m = (int)ceil_div(J.M, kTileM);
        J.tiles_n = (int)ceil_div(J.N, kTileN);
        J.tile_begin = begin;
        begin += J.mode == GEMM_BWD ? J.tiles_m * std::max(1, J.nsplit)
                                    : J.tiles_m * J.tiles_n * (J.mode == GEMM_STORE && J.ksplit > 1 ? J.ksplit : 1);
    }
    L.total_tiles = begin;
}

int gemm_set_trace(void* buf) {
    g_trace_host = static_cast<unsigned long long*>(buf);
    return STIL_OK;
}

template <int MODE, int OCC>
static int launch_gemm_mode(const GemmLaunch& L, cudaStream_t stream) {
    constexpr int smem = smem_bytes_for(OCC == 1 ? kMaxStages : 3);
    static std::atomic<unsigned long long> smem_set{0};
    STIL_CUDA(ensure_dynamic_smem(smem_set, reinterpret_cast<const void*>(&gemm_tc05_kernel<MODE, OCC>), smem));
    if (MODE == GEMM_STORE && L.cluster_k > 1) {
        STIL_CUDA(launch_pdl_cluster(gemm_tc05_kernel<MODE, OCC>, dim3(L.total_tiles), dim3(threads_for(OCC)), smem, stream,
                                     L.cluster_k, L));
        return STIL_OK;
    }
    STIL_CUDA(launch_pdl(gemm_tc05_kernel<MODE, OCC>, dim3(L.total_tiles), dim3(threads_for(OCC)), smem, stream, L));
    return STIL_OK;
}

// ---- fused backward: job construction and launch
int bwd_nsplit(int64_t n_cols, int64_t row_blocks_total) {
    // a CTA walks its column tiles one after the other (~1.5 us each): at most 4 per CTA, but no more CTAs than SMs
    const int64_t tiles = ceil_div(n_cols, kTileN);
    int64_t s = ceil_div(tiles, 4);
    while (s > 1 && s * row_blocks_total > 148) --s;
    return (int)std::max<int64_t>(1, std::min<int64_t>(s, tiles));
}

bool make_bwd_job(GemmJob& out, const GemmJob& grad, const GemmJob& store, int nsplit) {
    // EXPERIMENTAL, opt-in (STIL_FUSED_BWD=1).  Correct (tests/test_gpu_fused_bwd.py) but measured SLOWER than the
    // GRAD + STORE pair at every size tried: one thread=row pass over a 128x128 logits tile costs ~3 us on one SM, and
    // the fused kernel serialises those passes per row block (or pays a slice reduction kernel when the tiles are dealt
    // to separate CTAs) — C2 step 34.3 vs 30.9 us, one rank's N=8 InfoNCE 42.9 vs 36.9 us (profiles/README.md).
    static const bool on = [] { const char* e = getenv("STIL_FUSED_BWD"); return e && e[0] == '1'; }();
    if (!on) return false;
    if (grad.mode != GEMM_GRAD || store.mode != GEMM_STORE || !store.y_mn_major) return false;
    if (grad.D != 64 && grad.D != 128) return false;                 // whole embedding rows as 64-column boxes
    if (store.N != grad.D || store.D != grad.N || store.M != grad.M) return false;
    out = grad;
    out.mode = GEMM_BWD;
    out.npair2 = store.npair;
    int nx = 0, ny = 0;
    for (int p = 0; p < grad.npair; ++p) { nx = std::max(nx, grad.xseg[p] + 1); ny = std::max(ny, grad.yseg[p] + 1); }
    for (int p = 0; p < store.npair; ++p) {
        out.gseg2[p] = store.xseg[p];
        out.yseg2[p] = store.yseg[p];
        ny = std::max(ny, store.yseg[p] + 1);
        if (store.xseg[p] >= grad.g_nseg) return false;
    }
    out.bw_nx = nx; out.bw_ny = ny;
    const int nbx = grad.D / 64;
    const int x_bytes = nx * nbx * kBoxBytes, y_bytes = ny * nbx * kBoxBytes, g_bytes = grad.g_nseg * 2 * kBoxBytes;
    const int limit = 227 * 1024 - 1024 /*alignment slack*/ - kBwdMiscBytes;
    if (x_bytes + y_bytes + g_bytes > limit) return false;
    out.bw_nbuf = (x_bytes + 2 * y_bytes + g_bytes <= limit) ? 2 : 1;
    out.bw_y_off = x_bytes;
    out.bw_y_stride = y_bytes;
    out.bw_g_off = x_bytes + out.bw_nbuf * y_bytes;
    out.bw_misc_off = out.bw_g_off + g_bytes;
    out.nsplit = std::max(1, std::min(nsplit, (int)ceil_div(grad.N, kTileN)));
    if (ceil_div(ceil_div(grad.N, kTileN), out.nsplit) > 8) return false;   // a CTA walks its tiles serially: keep it short
    // outputs of the second product
    out.fin_dx = store.fin_dx; out.fin_dx_dtype = store.fin_dx_dtype; out.fin_ld_dx = store.fin_ld_dx;
    out.fin_x = store.fin_x; out.fin_x_dtype = store.fin_x_dtype; out.fin_ldx = store.fin_ldx; out.fin_sx = store.fin_sx;
    out.out = store.out; out.ld_out = store.ld_out; out.slice_stride = store.slice_stride;
    if (out.nsplit > 1 && (out.out == nullptr || out.slice_stride == 0)) return false;   // partial tiles need slices
    if (out.nsplit == 1 && out.fin_dx == nullptr && out.out == nullptr) return false;
    if (out.nsplit > 1) out.fin_dx = nullptr;
    return true;
}

static int launch_gemm_bwd(const GemmLaunch& L, cudaStream_t stream) {
    int smem = 0;
    for (int j = 0; j < L.njobs; ++j) smem = std::max(smem, L.job[j].bw_misc_off + kBwdMiscBytes + 1024);
    static std::atomic<unsigned long long> smem_set{0};
    STIL_CUDA(ensure_dynamic_smem(smem_set, reinterpret_cast<const void*>(&gemm_bwd_kernel), 227 * 1024));
    STIL_CUDA(launch_pdl(gemm_bwd_kernel, dim3(L.total_tiles), dim3(kBwdThreads), (size_t)smem, stream, L));
    return STIL_OK;
}

// All jobs of one launch share the epilogue mode (the kernel is specialised per mode to keep it small).
int launch_gemm(const GemmLaunch& L, cudaStream_t stream) {
    STIL_REQUIRE(L.njobs >= 1 && L.njobs <= kMaxGemmJobs, STIL_E_ARG, "gemm launch with %d jobs", L.njobs);
    if (L.total_tiles == 0) return STIL_OK;
    const int mode = L.job[0].mode;
    for (int j = 1; j < L.njobs; ++j)
        STIL_REQUIRE(L.job[j].mode == mode, STIL_E_ARG, "gemm launch mixes epilogue modes");
    static unsigned int launch_counter = 0;
    const_cast<GemmLaunch&>(L).trace_id = launch_counter++;
    const_cast<GemmLaunch&>(L).trace = g_trace_host;
    if (mode == GEMM_GRAD) {
        // bulk-store view of G: [ld_g columns, M rows, g_nseg segments], 64 x 128 boxes, 128-byte swizzle
        for (int j = 0; j < L.njobs; ++j) {
            GemmJob& J = const_cast<GemmLaunch&>(L).job[j];
            int rc = make_operand_map(&J.tmg, J.gop, J.ld_g, J.M, J.g_nseg, (int64_t)J.g_nseg * J.ld_g, J.ld_g, kTileM);
            if (rc) return rc;
        }
    }
    if (L.cluster_k > 1) {
        STIL_REQUIRE(mode == GEMM_STORE && L.cluster_k <= 8, STIL_E_ARG, "cluster split-K needs GEMM_STORE and cluster_k <= 8");
        for (int j = 0; j < L.njobs; ++j)
            STIL_REQUIRE(L.job[j].ksplit == L.cluster_k && L.job[j].slice_stride == 0, STIL_E_ARG,
                         "cluster split-K: every job's ksplit must equal cluster_k (no slice outputs)");
    }
    if (mode == GEMM_STATS || mode == GEMM_STORE) {
        static const bool no_tma_out = [] { const char* e = getenv("STIL_NO_TMA_OUT"); return e && e[0] == '1'; }();
        for (int j = 0; j < L.njobs; ++j) {
            GemmJob& J = const_cast<GemmLaunch&>(L).job[j];
            J.out_tma = 0;
            // plain fp32 outputs only (not the fused dX epilogue, not red.global accumulation into `out`)
            const bool sliced = mode == GEMM_STORE && J.ksplit > 1 && J.slice_stride > 0 && L.cluster_k <= 1;
            const bool accum = mode == GEMM_STORE && J.ksplit > 1 && J.slice_stride == 0 && L.cluster_k <= 1;
            // N % 4: a bulk tensor store clips out-of-range columns in 16-byte granules — with N = 130 it overwrote columns
            // 130-131, which belong to the neighbouring block of the CoMatch graphs (found by tests/test_gpu_banks.py)
            if (!no_tma_out && mode == GEMM_STORE && J.fin_dx && J.fin_dx_dtype == STIL_F32 && (J.N & 3) == 0) {
                // fused normalise-backward epilogue with an fp32 dX: the map is over the final gradient
                if (make_out_map(&J.tmg, static_cast<float*>(J.fin_dx), J.N, J.M, J.fin_ld_dx, 1, 0)) J.out_tma = 1;
                continue;
            }
            if (no_tma_out || !J.out || J.fin_dx || accum || J.fwd_norm || (J.N & 3) != 0) continue;
            if (make_out_map(&J.tmg, J.out, J.N, J.M, J.ld_out, sliced ? J.ksplit : 1, J.slice_stride)) J.out_tma = 1;
        }
    }
    if (mode == GEMM_BWD) return launch_gemm_bwd(L, stream);
    // more tiles than SMs: two CTAs per SM (3-deep rings) so epilogues overlap main loops
    const bool two = L.total_tiles > 148;
    if (mode == GEMM_STATS) return two ? launch_gemm_mode<GEMM_STATS, 2>(L, stream) : launch_gemm_mode<GEMM_STATS, 1>(L, stream);
    if (mode == GEMM_STORE) return two ? launch_gemm_mode<GEMM_STORE, 2>(L, stream) : launch_gemm_mode<GEMM_STORE, 1>(L, stream);
    return two ? launch_gemm_mode<GEMM_GRAD, 2>(L, stream) : launch_gemm_mode<GEMM_GRAD, 1>(L, stream);
}

}  // namespace stil
