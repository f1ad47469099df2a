// a7 bank sweep (SURVEY §8 a7 / e-C5; reference models/MatchModel/simmatch_model.py:268,281 — the two products
// `feat_kw @ bank` and `feat_qu @ bank` against a [dim, K_b] memory bank, and the backward product onto feat_qu):
// persistent tcgen05 kernels for the shapes where the general tiled GEMM (gemm_tc05.cu) is bound by operand re-fetch and
// per-tile latencies instead of the tensor pipe.
//
// The forward products have a SHORT contraction (dim <= 512) and a LONG bank axis (65536 columns).  With one CTA per
// 128 x 128 output tile every tile re-loads its 128 x dim feature rows (128 KiB at dim 512) next to its 128 KiB of bank
// columns, pays a CTA prologue (barrier init, TMEM allocation, tensor-map fetch, pipeline fill) per 1 us of tensor work
// and cannot overlap its epilogue with anything: 180 us for 2 x 448 x 65536 x 512 (profiles/r2_bank_timeline.txt).
//
// bank_logits_kernel: one CTA per SM.  A CTA owns a UNIT (teacher or student rows of one 128-row block) and walks over a
// range of 128-column bank blocks:
//   * the unit's rows live in TENSOR MEMORY (tcgen05.mma with the A operand from TMEM: 128 lanes x dim/2 columns of packed
//     bf16 pairs, written once per chunk with tcgen05.st), so shared memory belongs to the bank ring and the output
//     staging (rows in shared memory cost 128 KiB and left a 4-stage ring of 16 KiB stages).  A ring stage is 128 contraction
//     rows (32 KiB, 8 MMAs): a tcgen05.commit costs ~0.2 us of completion tracking, a 64-row stage is 0.13 us of tensor work
//     (the STIL_SWEEP_DEBUG switch-off experiments in profiles/r2_bank_timeline.txt);
//   * warp 0 streams bank blocks through the ring (TMA, the bank read in place as an MN-major operand), warp 1 issues the
//     MMAs into one of TWO 128-column TMEM accumulators, two groups of 4 warps drain the other one (tcgen05.ld ->
//     128-byte-swizzled fp32 boxes in shared memory -> bulk tensor stores; each group owns 64 of the 128 columns), so the
//     epilogue of block i runs under the MMAs of block i + 1;
//   * chunks (unit x column range) are dealt so that CTAs running at the same time sweep the SAME bank columns for
//     different units: the bank streams from HBM once per sweep and is shared through L2 (plus an L2 prefetch a few
//     blocks ahead of the ring).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "internal.h"
#include "tc05.cuh"

namespace stil {
namespace {

constexpr int kSweepGroups = 2;                    // epilogue groups, 64 columns of the block each
constexpr int kSweepEpiWarps = 4 * kSweepGroups;
constexpr int kSweepThreads = 64 + 32 * kSweepEpiWarps;   // TMA | MMA | 8 epilogue warps
constexpr int kBoxA = kTileM * 128;                // [128 rows x 64 bf16], 128-byte swizzle
constexpr int kStageK = 2 * kTileK;                // contraction rows per ring stage
constexpr int kStageB = kStageK * kTileN * 2;      // [128 contraction rows x 128 columns] as two 128 x 64 boxes (32 KiB)
constexpr int kOutBox = kTileM * 128;              // [128 rows x 32 fp32], 128-byte swizzle
constexpr int kOutBoxes = 2 * kSweepGroups;        // the whole 128 x 128 block
constexpr int kMaxStagesB = 6;
constexpr int kPrefetchAhead = 4;                  // bank blocks pulled into L2 ahead of the ring
constexpr int kSweepBarBytes = 256;
constexpr int kSmemLimit = 227 * 1024;
constexpr uint32_t kSweepTmemCols = 512;           // 2 accumulators (256 columns) + the unit's rows (dim / 2 <= 256 columns)
constexpr uint32_t kSweepTmemA = 256;

struct alignas(64) BankSweepLaunch {
    CUtensorMap tmb;      // bank [dim, k_shard] in place, MN-major: boxes 64 (columns) x 128 (contraction rows) [x 2 groups]
    CUtensorMap tmo[2];   // zt / zs fp32 [rows, ldz], boxes 32 x 128
    const __nv_bfloat16* feat[2];   // feat_ku / feat_qu [rows, ld]
    long long ld;
    int rows, k_shard, nbx /*stages per block = dim / 128*/, tiles_m, nblocks, ranges, stages, nchunks;
    int b_grouped;        // tmb is [64, dim, k_shard / 64] with box 64 x 64 x 2: one bulk load per stage
    int debug;            // STIL_SWEEP_DEBUG bits (timing experiments only): 1 no stores, 2 no bank loads, 4 no MMA, 8 no epilogue
};

struct Chunk {
    int job, m0, b0, b1;
};
__device__ __forceinline__ Chunk decode_chunk(const BankSweepLaunch& L, int c) {
    const int units = 2 * L.tiles_m;
    const int r = c / units, u = c - r * units;
    Chunk k;
    k.job = u & 1;
    k.m0 = (u >> 1) * kTileM;
    k.b0 = (int)(((long long)r * L.nblocks) / L.ranges);
    k.b1 = (int)(((long long)(r + 1) * L.nblocks) / L.ranges);
    return k;
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 lanes = rows, 16 contraction elements = 8 columns of packed bf16
// pairs) comes from tensor memory
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(kSweepThreads, 1) bank_logits_kernel(const __grid_constant__ BankSweepLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc05::smem_u32(smem_raw);
    uint8_t* base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* bs = base;                                   // [stages] bank ring
    uint8_t* os = bs + L.stages * kStageB;                // [kOutBoxes] output staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(os + kOutBoxes * kOutBox);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;                          // [kMaxStagesB]
    uint64_t* b_empty = b_full + kMaxStagesB;             // [kMaxStagesB]
    uint64_t* t_full = b_empty + kMaxStagesB;             // [2]
    uint64_t* t_empty = t_full + 2;                       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc05::tma_prefetch_desc(&L.tmb);
        tc05::tma_prefetch_desc(&L.tmo[0]);
        tc05::tma_prefetch_desc(&L.tmo[1]);
    }
    if (warp == 1) {
        if (lane == 0) {
            tc05::mbar_init(a_full, kSweepEpiWarps);
            for (int s = 0; s < kMaxStagesB; ++s) {
                tc05::mbar_init(&b_full[s], 1);
                tc05::mbar_init(&b_empty[s], 1);
            }
            for (int i = 0; i < 2; ++i) {
                tc05::mbar_init(&t_full[i], 1);
                tc05::mbar_init(&t_empty[i], kSweepEpiWarps);
            }
            tc05::fence_mbar_init();
        }
        __syncwarp();
        tc05::tmem_alloc(tmem_slot, kSweepTmemCols);
        tc05::tmem_relinquish();
    }
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int nbx = L.nbx, stages = L.stages;

    if (warp == 0) {
        // ===================== bank producer: every (chunk, block, contraction box) in MMA order =====================
        {
            const bool leader = tc05::elect_one();     // converged warp, one elected lane issues (see the MMA issuer)
            int s = 0;
            uint32_t ph = 0;
            const int units = 2 * L.tiles_m;
            for (int c = blockIdx.x; c < L.nchunks; c += gridDim.x) {
                const Chunk k = decode_chunk(L, c);
                const int u = c % units;
                // The CTAs that sweep this column range together (one per unit) take turns pulling the blocks
                // kPrefetchAhead ahead into L2: the ring then waits for L2 hits, not for DRAM.
                auto prefetch_block = [&](int nb) {
                    if (leader && nb < k.b1 && nb % units == u)
                        for (int kb = 0; kb < nbx; ++kb) {
                            if (L.b_grouped) {
                                tc05::tma_prefetch_l2_3d(&L.tmb, 0, kb * kStageK, nb * 2);
                            } else {
                                tc05::tma_prefetch_l2_3d(&L.tmb, nb * kTileN, kb * kStageK, 0);
                                tc05::tma_prefetch_l2_3d(&L.tmb, nb * kTileN + 64, kb * kStageK, 0);
                            }
                        }
                };
                for (int nb = k.b0; nb < k.b0 + kPrefetchAhead; ++nb) prefetch_block(nb);
                for (int nb = k.b0; nb < k.b1; ++nb) {
                    prefetch_block(nb + kPrefetchAhead);
                    for (int kb = 0; kb < nbx; ++kb) {
                        tc05::mbar_wait(&b_empty[s], ph ^ 1);
                        if (leader) {
                            if (L.debug & 2) {
                                tc05::mbar_arrive(&b_full[s]);
                            } else {
                                tc05::mbar_arrive_expect_tx(&b_full[s], kStageB);
                                uint8_t* dst = bs + s * kStageB;
                                if (L.b_grouped) {
                                    tc05::tma_load_3d(dst, &L.tmb, &b_full[s], 0, kb * kStageK, nb * 2);
                                } else {
                                    tc05::tma_load_3d(dst, &L.tmb, &b_full[s], nb * kTileN, kb * kStageK, 0);
                                    tc05::tma_load_3d(dst + 64 * kStageK * 2, &L.tmb, &b_full[s], nb * kTileN + 64, kb * kStageK, 0);
                                }
                            }
                        }
                        __syncwarp();
                        if (++s == stages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The WHOLE warp walks the loop (converged), one elected lane issues: every operand of tcgen05.mma / commit is then
        // warp-uniform for the compiler, which keeps descriptors and addresses in uniform registers.  Inside an
        // `if (lane == 0)` branch it emitted an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall of ~14 dependent instructions
        // per MMA: ~100 cycles of issue per 64 cycles of tensor work, i.e. the issuing thread was the bottleneck.
        {
            const uint32_t idesc = tc05::make_idesc_bf16_f32(kTileM, kTileN) | (1u << 16);   // B MN-major
            const uint32_t bs_a = tc05::smem_u32(bs);
            const bool leader = tc05::elect_one();
            int s = 0, ci = 0, ti = 0;
            uint32_t ph = 0;
            for (int c = blockIdx.x; c < L.nchunks; c += gridDim.x, ++ci) {
                const Chunk k = decode_chunk(L, c);
                tc05::mbar_wait(a_full, ci & 1);                              // the unit's rows are in tensor memory
                tc05::fence_after_sync();
                for (int nb = k.b0; nb < k.b1; ++nb, ++ti) {
                    const int sb = ti & 1;
                    tc05::mbar_wait(&t_empty[sb], ((ti >> 1) & 1) ^ 1);    // the epilogue has drained this accumulator
                    tc05::fence_after_sync();
                    for (int kb = 0; kb < nbx; ++kb) {
                        tc05::mbar_wait(&b_full[s], ph);
                        tc05::fence_after_sync();
                        // 64-column groups 16 KiB apart (LBO), 16 contraction rows = 2 KiB per MMA (+128 in the address field)
                        const uint64_t b_desc = tc05::make_mnmajor_sw128_desc(bs_a + s * kStageB, 64 * kStageK * 2);
                        if (leader) {
                            if (!(L.debug & 4)) {
#pragma unroll
                                for (int kk = 0; kk < kStageK / 16; ++kk)
                                    mma_f16_ts(tmem_base + sb * kTileN, tmem_base + kSweepTmemA + kb * (kStageK / 2) + kk * 8, b_desc + 128u * kk,
                                               idesc, (kb | kk) ? 1u : 0u);
                            }
                            tc05::mma_commit(&b_empty[s]);
                        }
                        __syncwarp();
                        if (++s == stages) { s = 0; ph ^= 1; }
                    }
                    if (leader) tc05::mma_commit(&t_full[sb]);
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================== epilogue: thread = accumulator row; group g owns columns [64 g, 64 g + 64) =====================
        const int ew = warp - 2;
        const int grp = ew >> 2;                // 0 / 1
        const int q = warp & 3;                 // TMEM lane quarter of this warp
        const int r_in = q * 32 + lane;
        const bool issuer = (ew & 3) == 0 && lane == 0;      // one thread per group issues its bulk stores
        uint8_t* gos = os + grp * 2 * kOutBox;
        int ti = 0, ci = 0;
        for (int c = blockIdx.x; c < L.nchunks; c += gridDim.x, ++ci) {
            const Chunk k = decode_chunk(L, c);
            // ---- the unit's rows -> tensor memory (A operand): lane = row, column j = elements (2j, 2j+1).  The MMAs of the
            // previous chunk are complete (its last accumulator was drained below).  The groups split the contraction boxes.
            {
                const int row = k.m0 + r_in;
                const __nv_bfloat16* src = L.feat[k.job] + (long long)row * L.ld;
                for (int kb = grp; kb < 2 * nbx; kb += kSweepGroups) {
                    uint32_t v[32];
                    if (row < L.rows) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint4 x = __ldg(reinterpret_cast<const uint4*>(src + kb * kTileK) + j);
                            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0u;
                    }
                    tc05::tmem_st_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kSweepTmemA + kb * 32, v);
                }
                tc05::tmem_st_wait();
                tc05::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(a_full);
            }
            for (int nb = k.b0; nb < k.b1; ++nb, ++ti) {
                const int sb = ti & 1, n0 = nb * kTileN;
                tc05::mbar_wait(&t_full[sb], (ti >> 1) & 1);
                tc05::fence_after_sync();
                uint32_t acc[2][32];
                if (!(L.debug & 8)) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc)
                        tc05::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sb * kTileN + (grp * 2 + cc) * 32, acc[cc]);
                    tc05::tmem_ld_wait();
                }
                tc05::fence_before_sync();      // this warp's columns are in registers: hand the accumulator back
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(&t_empty[sb]);
                if (L.debug & 8) continue;
                // the group's previous bulk stores must have read the staging boxes before they are overwritten
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint8_t* box = gos + cc * kOutBox + r_in * 128;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4)
                        *reinterpret_cast<uint4*>(box + ((k4 ^ (r_in & 7)) * 16)) =
                            make_uint4(acc[cc][4 * k4], acc[cc][4 * k4 + 1], acc[cc][4 * k4 + 2], acc[cc][4 * k4 + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if (issuer) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int col = n0 + (grp * 2 + cc) * 32;
                        if (col < L.k_shard && !(L.debug & 1)) tc05::tma_store_3d(&L.tmo[k.job], gos + cc * kOutBox, col, k.m0, 0);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }

    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc05::fence_after_sync();
        tc05::tmem_dealloc(tmem_base, kSweepTmemCols);
    }
}

// how many column ranges every unit is cut into: fill the SMs, few waves, chunks long enough to amortise the reload of
// the unit's rows (~2 block times)
int pick_ranges(int units, int nblocks, int sms) {
    int best = 1;
    double best_cost = 1e30;
    for (int r = 1; r <= std::min(nblocks, 1024); ++r) {
        const double chunks = (double)units * r;
        const double waves = std::ceil(chunks / sms);
        const double cost = waves * ((double)nblocks / r + 2.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = r; }
    }
    return best;
}

// =====================================================================================================================
// bank_dx_kernel: d_feat[rows, dim] (+)= G[rows, (hi, lo), K_b] · bankᵀ — the backward product of simmatch_model.py:281,
// a LONG contraction (the bank axis) onto a small output.  One CTA per SM owns a 128-row block and a range of the
// contraction and keeps the WHOLE [128 x dim] fp32 accumulator in tensor memory (dim <= 512 columns = all of TMEM), so a
// G tile is fetched once for all dim output columns (the tiled GEMM re-fetched it per 128-column tile: 4x at dim 512)
// and each stage carries [G hi | G lo | dim bank rows] x 64 contraction columns.  The partial tiles of the CTAs that
// share a row block are added with 16-byte reductions into the zero-initialised output.
constexpr int kDxEpiWarps = 4;
constexpr int kDxThreads = 64 + 32 * kDxEpiWarps;
constexpr int kDxMaxStages = 6;
constexpr int kDxPrefetchAhead = 6;
constexpr int kDxOutBoxes = 8;                    // output boxes staged per round of bulk reductions

struct alignas(64) BankDxLaunch {
    CUtensorMap tmg;      // G [rows, nseg, K_b] K-major, boxes 64 x 128
    CUtensorMap tmb;      // bank [dim, K_b] K-major, boxes 64 x 128
    CUtensorMap tmo;      // out fp32 [rows, ld_out], boxes 32 x 128 (bulk reductions)
    float* out;           // [rows, ld_out] fp32, zero-initialised
    const float* row_scale;   // optional upstream gradient per row
    long long ld_out;
    int rows, dim, nseg, nbn /*128-row bank boxes*/, kboxes, ksplit, stages;
    int debug;            // STIL_SWEEP_DEBUG bits: 16 no loads, 32 no MMA
    uint32_t tmem_cols;
};

__global__ void __launch_bounds__(kDxThreads, 1) bank_dx_kernel(const __grid_constant__ BankDxLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc05::smem_u32(smem_raw);
    uint8_t* base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int stage_bytes = (L.nseg + L.nbn) * kBoxA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + L.stages * stage_bytes);
    uint64_t* full = bars;                       // [kDxMaxStages]
    uint64_t* empty = bars + kDxMaxStages;       // [kDxMaxStages]
    uint64_t* acc_full = empty + kDxMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ks = blockIdx.x % L.ksplit, tm = blockIdx.x / L.ksplit;
    const int m0 = tm * kTileM;
    const int kb0 = (int)(((long long)ks * L.kboxes) / L.ksplit), kb1 = (int)(((long long)(ks + 1) * L.kboxes) / L.ksplit);
    if (warp == 0 && lane == 0) {
        tc05::tma_prefetch_desc(&L.tmg);
        tc05::tma_prefetch_desc(&L.tmb);
        tc05::tma_prefetch_desc(&L.tmo);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kDxMaxStages; ++s) {
                tc05::mbar_init(&full[s], 1);
                tc05::mbar_init(&empty[s], 1);
            }
            tc05::mbar_init(acc_full, 1);
            tc05::fence_mbar_init();
        }
        __syncwarp();
        tc05::tmem_alloc(tmem_slot, L.tmem_cols);
        tc05::tmem_relinquish();
    }
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // whole warp converged, one elected lane issues (uniform operands: see the MMA issuer of bank_logits_kernel)
        {
            const bool leader = tc05::elect_one();
            int s = 0;
            uint32_t ph = 0;
            const int tiles_m = (int)gridDim.x / L.ksplit;
            // L2 prefetch kDxPrefetchAhead boxes ahead of the ring: this CTA's G rows, and the bank boxes in turns with the
            // CTAs of the other row blocks that walk the same contraction range
            auto prefetch_box = [&](int kb) {
                if (kb >= kb1 || !leader) return;
                for (int sg = 0; sg < L.nseg; ++sg) tc05::tma_prefetch_l2_3d(&L.tmg, kb * kTileK, m0, sg);
                if (kb % tiles_m == tm)
                    for (int nb = 0; nb < L.nbn; ++nb) tc05::tma_prefetch_l2_3d(&L.tmb, kb * kTileK, nb * kTileN, 0);
            };
            for (int kb = kb0; kb < kb0 + kDxPrefetchAhead; ++kb) prefetch_box(kb);
            for (int kb = kb0; kb < kb1; ++kb) {
                prefetch_box(kb + kDxPrefetchAhead);
                tc05::mbar_wait(&empty[s], ph ^ 1);
                if (leader && (L.debug & 16)) {
                    tc05::mbar_arrive(&full[s]);
                } else if (leader) {
                    tc05::mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
                    uint8_t* dst = base + s * stage_bytes;
                    for (int sg = 0; sg < L.nseg; ++sg) tc05::tma_load_3d(dst + sg * kBoxA, &L.tmg, &full[s], kb * kTileK, m0, sg);
                    for (int nb = 0; nb < L.nbn; ++nb)
                        tc05::tma_load_3d(dst + (L.nseg + nb) * kBoxA, &L.tmb, &full[s], kb * kTileK, nb * kTileN, 0);
                }
                __syncwarp();
                if (++s == L.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = tc05::elect_one();
            int s = 0;
            uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                tc05::mbar_wait(&full[s], ph);
                tc05::fence_after_sync();
                const uint32_t st_a = tc05::smem_u32(base + s * stage_bytes);
                if (leader) {
                    for (int sg = 0; sg < ((L.debug & 32) ? 0 : L.nseg); ++sg) {
                        const uint64_t a_desc = tc05::make_kmajor_sw128_desc(st_a + sg * kBoxA);
                        for (int n0 = 0; n0 < L.dim; n0 += 256) {      // up to 256 output columns per instruction
                            const int nn = min(256, L.dim - n0);
                            const uint32_t idesc = tc05::make_idesc_bf16_f32(kTileM, (uint32_t)nn);
                            const uint64_t b_desc = tc05::make_kmajor_sw128_desc(st_a + (L.nseg + n0 / kTileN) * kBoxA);
#pragma unroll
                            for (int kk = 0; kk < kTileK / 16; ++kk)
                                tc05::mma_f16_ss(tmem_base + n0, a_desc + 2u * kk, b_desc + 2u * kk, idesc, (kb > kb0 || sg || kk) ? 1u : 0u);
                        }
                    }
                    tc05::mma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == L.stages) { s = 0; ph ^= 1; }
            }
            if (leader) {
                if (kb1 > kb0) tc05::mma_commit(acc_full);
                else tc05::mbar_arrive(acc_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < L.rows;
        const float rs = (L.row_scale && row_ok) ? L.row_scale[row] : 1.f;
        tc05::mbar_wait(acc_full, 0);
        tc05::fence_after_sync();
        if (kb1 > kb0) {
            // The partial tile leaves as BULK REDUCTIONS: registers -> 128-byte-swizzled [128 rows x 32 fp32] boxes in the
            // (now idle) operand ring -> cp.reduce.async.bulk.tensor (.add) into the zero-initialised output.  Per-thread
            // red.global.add.v4 (thread = row: 32 scattered 16-byte atomics per instruction) took 28 us for the 148 x 256 KiB
            // of partial tiles at C5, as long as the whole main loop.
            const int e = threadIdx.x - 64;
            const int r_in = q * 32 + lane;
            const int nbox = min(kDxOutBoxes, (L.stages * stage_bytes) / kOutBox);
            for (int c0 = 0; c0 < L.dim; c0 += 32 * nbox) {
                if (c0 > 0) {       // the previous round's reductions have read their boxes
                    if (e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kDxEpiWarps) : "memory");
                }
                const int nb = min(nbox, (L.dim - c0) / 32);
                for (int b = 0; b < nb; ++b) {
                    uint32_t acc[32];
                    tc05::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0 + b * 32, acc);
                    tc05::tmem_ld_wait();
                    uint8_t* box = base + b * kOutBox + r_in * 128;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4)
                        *reinterpret_cast<float4*>(box + ((k4 ^ (r_in & 7)) * 16)) =
                            make_float4(__uint_as_float(acc[4 * k4]) * rs, __uint_as_float(acc[4 * k4 + 1]) * rs,
                                        __uint_as_float(acc[4 * k4 + 2]) * rs, __uint_as_float(acc[4 * k4 + 3]) * rs);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kDxEpiWarps) : "memory");
                if (e == 0) {
                    for (int b = 0; b < nb; ++b) tc05::tma_reduce_add_3d(&L.tmo, base + b * kOutBox, c0 + b * 32, m0, 0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc05::fence_after_sync();
        tc05::tmem_dealloc(tmem_base, L.tmem_cols);
    }
}

}  // namespace

bool bank_logits_eligible(int dtype, int64_t rows, int64_t dim, int64_t k_shard, int64_t ldz) {
    static const bool off = [] { const char* e = std::getenv("STIL_BANK_SWEEP"); return e && e[0] == '0'; }();
    return !off && dtype == STIL_BF16 && rows >= 1 && dim >= kStageK && dim % kStageK == 0 && dim <= 512 && k_shard % 8 == 0 && ldz % 4 == 0;
}

int launch_bank_logits(const void* feat_ku, const void* feat_qu, int64_t rows, int64_t dim, int64_t ld, const void* bank,
                       int64_t ld_bank, int64_t k_shard, float* zt, float* zs, int64_t ldz, cudaStream_t stream) {
    BankSweepLaunch L;
    std::memset(&L, 0, sizeof(L));
    int rc;
    STIL_REQUIRE(ld % 8 == 0 && ((reinterpret_cast<uintptr_t>(feat_ku) | reinterpret_cast<uintptr_t>(feat_qu)) & 15) == 0, STIL_E_ALIGN,
                 "bank sweep: feature rows must be 16-byte aligned (ld %% 8 == 0)");
    L.b_grouped = k_shard % 64 == 0;
    if (L.b_grouped) {
        // [64 columns, dim rows, k_shard / 64 column groups]: one box 64 x 128 x 2 is a whole stage
        if ((rc = make_operand_map(&L.tmb, bank, 64, dim, k_shard / 64, ld_bank, 64, kStageK, 2))) return rc;
    } else if ((rc = make_operand_map(&L.tmb, bank, k_shard, dim, 1, ld_bank, ld_bank * dim, kStageK))) {
        return rc;
    }
    STIL_REQUIRE(make_out_map(&L.tmo[0], zt, k_shard, rows, ldz, 1, 0) && make_out_map(&L.tmo[1], zs, k_shard, rows, ldz, 1, 0),
                 STIL_E_ALIGN, "bank sweep: the logits buffers need a 16-byte aligned base and ld %% 4 == 0");
    static const int sms = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
        return n;
    }();
    L.feat[0] = static_cast<const __nv_bfloat16*>(feat_ku);
    L.feat[1] = static_cast<const __nv_bfloat16*>(feat_qu);
    L.ld = ld;
    L.rows = (int)rows;
    L.k_shard = (int)k_shard;
    L.nbx = (int)(dim / kStageK);
    L.tiles_m = (int)ceil_div(rows, kTileM);
    L.nblocks = (int)ceil_div(k_shard, kTileN);
    L.ranges = pick_ranges(2 * L.tiles_m, L.nblocks, sms);
    L.nchunks = 2 * L.tiles_m * L.ranges;
    static const int debug = [] { const char* e = std::getenv("STIL_SWEEP_DEBUG"); return e ? atoi(e) : 0; }();
    L.debug = debug;
    const int fixed = 1024 + kOutBoxes * kOutBox + kSweepBarBytes;
    L.stages = std::min(kMaxStagesB, (kSmemLimit - fixed) / kStageB);
    const int smem = fixed + L.stages * kStageB;
    static std::atomic<unsigned long long> smem_set{0};
    STIL_CUDA(ensure_dynamic_smem(smem_set, reinterpret_cast<const void*>(&bank_logits_kernel), kSmemLimit));
    bank_logits_kernel<<<std::min(sms, L.nchunks), kSweepThreads, smem, stream>>>(L);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

bool bank_dx_eligible(int dtype, int64_t rows, int64_t dim, int64_t k_shard, const float* out, int64_t ld_out) {
    static const bool off = [] { const char* e = std::getenv("STIL_BANK_SWEEP"); return e && e[0] == '0'; }();
    return !off && dtype == STIL_BF16 && rows >= 1 && dim >= 64 && dim % 64 == 0 && dim <= 512 && k_shard % 8 == 0 && ld_out % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

int launch_bank_dx(const __nv_bfloat16* gop, int64_t ldg, int g_nseg, int64_t rows, const void* bank, int64_t ld_bank, int64_t dim,
                   int64_t k_shard, const float* row_scale, float* out, int64_t ld_out, cudaStream_t stream, bool out_is_zero) {
    BankDxLaunch L;
    std::memset(&L, 0, sizeof(L));
    int rc;
    if ((rc = make_operand_map(&L.tmg, gop, k_shard, rows, g_nseg, (int64_t)g_nseg * ldg, ldg, 128))) return rc;
    if ((rc = make_operand_map(&L.tmb, bank, k_shard, dim, 1, ld_bank, ld_bank * dim, 128))) return rc;
    STIL_REQUIRE(make_out_map(&L.tmo, out, dim, rows, ld_out, 1, 0), STIL_E_ALIGN, "bank dX: output needs a 16-byte aligned base and ld %% 4 == 0");
    static const int sms = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
        return n;
    }();
    L.out = out; L.row_scale = row_scale; L.ld_out = ld_out;
    L.rows = (int)rows; L.dim = (int)dim; L.nseg = g_nseg;
    L.nbn = (int)ceil_div(dim, kTileN);
    L.kboxes = (int)ceil_div(k_shard, kTileK);
    const int tiles_m = (int)ceil_div(rows, kTileM);
    L.ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(L.kboxes, sms / tiles_m));
    const int stage_bytes = (L.nseg + L.nbn) * kBoxA;
    L.stages = std::min(kDxMaxStages, (kSmemLimit - 1024 - kSweepBarBytes) / stage_bytes);
    STIL_REQUIRE(L.stages >= 2, STIL_E_SHAPE, "bank dX: no room for two stages at dim %lld", (long long)dim);
    static const int debug = [] { const char* e = std::getenv("STIL_SWEEP_DEBUG"); return e ? atoi(e) : 0; }();
    L.debug = debug;
    L.tmem_cols = 32;
    while ((int64_t)L.tmem_cols < dim) L.tmem_cols <<= 1;
    const int smem = 1024 + kSweepBarBytes + L.stages * stage_bytes;
    static std::atomic<unsigned long long> smem_set{0};
    STIL_CUDA(ensure_dynamic_smem(smem_set, reinterpret_cast<const void*>(&bank_dx_kernel), kSmemLimit));
    // the partial tiles are ADDED: start from zero
    if (!out_is_zero) STIL_CUDA(cudaMemset2DAsync(out, ld_out * sizeof(float), 0, dim * sizeof(float), rows, stream));
    bank_dx_kernel<<<tiles_m * L.ksplit, kDxThreads, smem, stream>>>(L);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

}  // namespace stil
