// Row-wise (HBM-bound) kernels of the STiL head: operand preparation, CGPL+PGLS pseudo-labelling,
// soft-label argmax, statistics merge / loss terms, normalise-backward, masked soft-target CE and the
// segmented per-class prototype sums.  Warp-shuffle reductions, vectorised coalesced row access.
#include <algorithm>
#include <cstdlib>

#include "internal.h"

namespace stil {

namespace {

constexpr int kRowBlock = 256;  // 8 warps (upper bound; small problems use 2-warp blocks to spread over the SMs)

// Warp-per-row kernels are latency-bound at the reference sizes (a few hundred rows): fewer warps per block means
// more SMs busy and a whole issue slot per warp.
inline int row_block_threads(int64_t warps_needed) { return warps_needed <= 4096 ? 64 : kRowBlock; }

// =====================================================================================
// Operand preparation
// =====================================================================================
__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& l, __nv_bfloat16& ll) {
    h = __float2bfloat16_rn(x);
    float r = x - __bfloat162float(h);
    l = __float2bfloat16_rn(r);
    r -= __bfloat162float(l);
    ll = __float2bfloat16_rn(r);
}

constexpr int kPrepRows = kRowBlock / 32;  // one warp per row

// Programmatic dependent launch (no-ops unless the launch carries the attribute): let the next kernel of the
// chain start its prologue now, and wait for the previous one's results before touching global memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// exp(x - m), m * log2(e) formed once per row, as ONE fused multiply-add and ONE special-function instruction (ex2.approx: relative error <= 2^-22; the rounding
// of x * log2(e) adds |x| * 2^-24).  libdevice expf costs ~14 instructions per call and made the row kernels
// instruction-bound (ncu, profiles/r2_ncu_rows_before.txt: 65 % issue-slot utilisation at 38 % of the HBM peak); the softmax
// probabilities differ from torch's by <= 2e-6, far inside the 1e-5 ambiguity band of DESIGN.md §2.
__device__ __forceinline__ float exp_fast_shift(float x, float m_log2e) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(x, 1.4426950408889634f, -m_log2e)));
    return y;
}
__device__ __forceinline__ float log_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.6931471805599453f;
}

__global__ void __launch_bounds__(kRowBlock) prep_kernel(const __grid_constant__ PrepLaunch L) {
    pdl_wait();                 // predecessor complete and visible ...
    pdl_launch_dependents();    // ... before the next kernel of the chain may start (see gemm_tc05.cu)
    // reduction tickets of the later kernels of this op start from zero (workspace content is arbitrary)
    if (blockIdx.x == 0 && (int)threadIdx.x < L.n_zero) L.zero_words[threadIdx.x] = 0u;
    int jid = 0;
#pragma unroll
    for (int j = 1; j < kMaxPrepJobs; ++j)
        if (j < L.njobs && (int)blockIdx.x >= L.job[j].block_begin) jid = j;
    const PrepJob& J = L.job[jid];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = (blockIdx.x - J.block_begin) * kPrepRows + warp;
    if (r >= J.rows) return;
    float ss = 0.f;
    const int odim = J.op_dim > 0 ? J.op_dim : J.dim;
    const bool vec = (J.dim % 128 == 0) && (odim == J.dim) && (J.ld % 4 == 0) &&
                     (reinterpret_cast<uintptr_t>(J.x) % (J.dtype == STIL_BF16 ? 8 : 16) == 0);
    if (vec) {
        // each lane owns 4 consecutive elements of every 128-element slab
        for (int d0 = lane * 4; d0 < J.dim; d0 += 128) {
            float x[4];
            if (J.dtype == STIL_BF16) {
                const uint2 v = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(J.x) + (long long)r * J.ld + d0);
                x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xffff0000u);
                x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xffff0000u);
            } else {
                const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(J.x) + (long long)r * J.ld + d0);
                x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
            }
            ss += x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
            if (J.op) {
                __nv_bfloat16 h[4], l[4], ll[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split3(x[j], h[j], l[j], ll[j]);
                __nv_bfloat16* o = J.op + ((long long)r * J.nseg) * J.dim + d0;
                auto pack = [](const __nv_bfloat16 (&q)[4]) {
                    return make_uint2((uint32_t)__bfloat16_as_ushort(q[0]) | ((uint32_t)__bfloat16_as_ushort(q[1]) << 16),
                                      (uint32_t)__bfloat16_as_ushort(q[2]) | ((uint32_t)__bfloat16_as_ushort(q[3]) << 16));
                };
                *reinterpret_cast<uint2*>(o) = pack(h);
                if (J.nseg > 1) *reinterpret_cast<uint2*>(o + J.dim) = pack(l);
                if (J.nseg > 2) *reinterpret_cast<uint2*>(o + 2 * J.dim) = pack(ll);
            }
        }
    } else {
        for (int d = lane; d < odim; d += 32) {
            const float x = d < J.dim ? ld_as_float(J.x, J.dtype, (long long)r * J.ld + d) : 0.f;
            ss += x * x;
            if (J.op) {
                __nv_bfloat16 h, l, ll;
                split3(x, h, l, ll);
                __nv_bfloat16* o = J.op + ((long long)r * J.nseg) * odim + d;
                o[0] = h;
                if (J.nseg > 1) o[odim] = l;
                if (J.nseg > 2) o[2 * odim] = ll;
            }
        }
    }
    if (J.inv_norm) {
        ss = warp_sum(ss);
        if (lane == 0) J.inv_norm[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (clip_loss.py:29-30)
    }
}

// =====================================================================================
// CGPL + PGLS (STiLModel.py:262-298), LPR lanes per row, NV float2 items per lane
// =====================================================================================
struct CgplArgs {
    const void *y_m, *y_i, *y_t;
    int logit_dtype;
    long long ld_y;
    const float* tl;
    long long ld_t;
    int rows, k;
    float temperature, rate_pseudo, one_minus_rate, th1;
    float inv_temperature, third;   // 1.0f / T and 1.0f / 3.0f: tensor / python-scalar is a multiply by the fp32 reciprocal in
                                    // torch's CUDA eager kernels, which is where the reference trains
    int past_start;
    float* pseudo_label;
    long long ld_pl;
    float* prediction;
    long long ld_pred;
    float* max_prob;
    long long* max_idx;
    unsigned char *mask1, *case1, *case2_i, *case2_t, *case3;
    long long* top1;
    int* cls;
    unsigned char* conf;
    int vec_y, vec_t, vec_pl, vec_pred;  // 8-byte (f32) / 4-byte (bf16) pair access allowed
    // labelled rows ride along: cls = y_l, conf = (1 >= th1)  (one-hot rows of pseudo_label_all, STiLModel.py:321)
    const long long* y_l;
    int b_l;
    int* cls_l;
    unsigned char* conf_l;
    // optional replacement of softmax(y_m) as the thresholded `prediction` — the distribution-aligned probabilities of
    // STiLModel.py:276-277 (DA == True); the agreement cases and the pseudo label still come from the logits
    const float* pred_in;
    long long ld_pin;
    int vec_pin;
};

template <int LPR, int NV>
__device__ __forceinline__ void load_row(const void* base, int dtype, long long ld, int row, int k, int sub, int vec,
                                         float (&a)[2 * NV]) {
#pragma unroll
    for (int it = 0; it < NV; ++it) {
        const int e0 = 2 * (sub + LPR * it);
        float x0 = -INFINITY, x1 = -INFINITY;
        if (e0 + 1 < k && vec) {
            if (dtype == STIL_BF16) {
                const __nv_bfloat162 p =
                    *reinterpret_cast<const __nv_bfloat162*>(static_cast<const __nv_bfloat16*>(base) + (long long)row * ld + e0);
                x0 = __bfloat162float(p.x);
                x1 = __bfloat162float(p.y);
            } else {
                const float2 p = *reinterpret_cast<const float2*>(static_cast<const float*>(base) + (long long)row * ld + e0);
                x0 = p.x;
                x1 = p.y;
            }
        } else {
            if (e0 < k) x0 = ld_as_float(base, dtype, (long long)row * ld + e0);
            if (e0 + 1 < k) x1 = ld_as_float(base, dtype, (long long)row * ld + e0 + 1);
        }
        a[2 * it] = x0;
        a[2 * it + 1] = x1;
    }
}

template <int LPR, int NV>
__device__ __forceinline__ void store_row(float* base, long long ld, int row, int k, int sub, int vec,
                                          const float (&a)[2 * NV]) {
#pragma unroll
    for (int it = 0; it < NV; ++it) {
        const int e0 = 2 * (sub + LPR * it);
        if (e0 + 1 < k && vec) {
            *reinterpret_cast<float2*>(base + (long long)row * ld + e0) = make_float2(a[2 * it], a[2 * it + 1]);
        } else {
            if (e0 < k) base[(long long)row * ld + e0] = a[2 * it];
            if (e0 + 1 < k) base[(long long)row * ld + e0 + 1] = a[2 * it + 1];
        }
    }
}

// FAST path of the row kernels (fp32 rows, even k, 8-byte aligned rows — every reference shape): one float2 per slot, ONE
// predicate per slot and a 32-bit offset from a row pointer formed once; the generic load_row above costs two bounds tests,
// a dtype test and a 64-bit address per element (ncu, profiles/r2_ncu_rows_before.txt: ISETP + IMAD + BRA = 22 % of the
// instructions of an instruction-bound kernel)
template <int LPR, int NV>
__device__ __forceinline__ void load_row_fast(const float* base, long long ld, int row, int khalf, int sub, float (&a)[2 * NV]) {
    const float2* p = reinterpret_cast<const float2*>(base + (long long)row * ld);
#pragma unroll
    for (int it = 0; it < NV; ++it) {
        const int i2 = sub + LPR * it;
        const float2 v = i2 < khalf ? p[i2] : make_float2(-INFINITY, -INFINITY);
        a[2 * it] = v.x;
        a[2 * it + 1] = v.y;
    }
}
template <int LPR, int NV>
__device__ __forceinline__ void store_row_fast(float* base, long long ld, int row, int khalf, int sub, const float (&a)[2 * NV]) {
    float2* p = reinterpret_cast<float2*>(base + (long long)row * ld);
#pragma unroll
    for (int it = 0; it < NV; ++it) {
        const int i2 = sub + LPR * it;
        if (i2 < khalf) p[i2] = make_float2(a[2 * it], a[2 * it + 1]);
    }
}

// Reductions over the LPR lanes of a row.  LPR <= 32: shuffles inside the warp.  LPR > 32 (one row per BLOCK of LPR
// threads, used for small batches where the per-row dependency chain — not bandwidth — bounds the kernel): warp shuffles,
// then the warps' values through shared memory; every thread of the block takes part (control flow is row-uniform).
template <int LPR>
__device__ __forceinline__ float row_max(float v) {
    if constexpr (LPR <= 32) {
        return group_max<LPR>(v);
    } else {
        __shared__ float sh[LPR / 32];
        v = warp_max(v);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        float r = sh[0];
#pragma unroll
        for (int w = 1; w < LPR / 32; ++w) r = fmaxf(r, sh[w]);
        __syncthreads();
        return r;
    }
}
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
    if constexpr (LPR <= 32) {
        return group_sum<LPR>(v);
    } else {
        __shared__ float sh[LPR / 32];
        v = warp_sum(v);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        float r = sh[0];
#pragma unroll
        for (int w = 1; w < LPR / 32; ++w) r += sh[w];      // fixed order: deterministic
        __syncthreads();
        return r;
    }
}

// in: logits a (−inf padded). out: e = exp(a − max) in place; returns the row sum (all lanes of the group).
template <int LPR, int NV>
__device__ __forceinline__ float softmax_exp(float (&a)[2 * NV]) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2 * NV; ++j) m = fmaxf(m, a[j]);
    m = row_max<LPR>(m);
    float s = 0.f;
    const float ml = m * 1.4426950408889634f;
#pragma unroll
    for (int j = 0; j < 2 * NV; ++j) {
        a[j] = exp_fast_shift(a[j], ml);     // padding slots hold -inf -> 0
        s += a[j];
    }
    return row_sum<LPR>(s);
}

// first-index argmax across the LPR lanes of a row group
template <int LPR>
__device__ __forceinline__ void group_argmax(float& v, int& idx) {
    constexpr int W = LPR < 32 ? LPR : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
    if constexpr (LPR > 32) {
        __shared__ float shv[LPR / 32];
        __shared__ int shi[LPR / 32];
        if ((threadIdx.x & 31) == 0) { shv[threadIdx.x >> 5] = v; shi[threadIdx.x >> 5] = idx; }
        __syncthreads();
        v = shv[0];
        idx = shi[0];
#pragma unroll
        for (int w = 1; w < LPR / 32; ++w)
            if (shv[w] > v || (shv[w] == v && shi[w] < idx)) { v = shv[w]; idx = shi[w]; }
        __syncthreads();
    }
}

// argmax_k softmax(y)_k for a head whose probabilities are not needed: softmax is monotone, so this is the first
// index of the largest LOGIT.  (torch takes the argmax of the rounded probabilities, STiLModel.py:263; the two can
// only differ when an earlier, smaller logit rounds to the same probability — logits closer than one ulp, rows that
// DESIGN.md §2 classifies as ambiguous.)
template <int LPR, int NV>
__device__ __forceinline__ int argmax_logits(const float (&y)[2 * NV], int sub, int k) {
    // slots beyond k hold -inf (load_row), so they never win against a finite logit; ascending idx within a lane and the
    // strict > keep the first maximum.  The lane's first slot is the initial candidate (a lane whose slots are all
    // padding then offers (-inf, idx >= k), which loses every comparison in group_argmax).
    float bv = y[0];
    int bi = 2 * sub;
#pragma unroll
    for (int j = 1; j < 2 * NV; ++j) {
        const int idx = 2 * (sub + LPR * (j >> 1)) + (j & 1);
        if (y[j] > bv) {
            bv = y[j];
            bi = idx;
        }
    }
    group_argmax<LPR>(bv, bi);
    return bi < k ? bi : 0;      // a row of NaNs / -inf compares false everywhere: keep the index in range
}

template <int LPR, int NV, bool FAST>
__global__ void __launch_bounds__(kRowBlock, (NV <= 5 ? 3 : 1)) cgpl_pgls_kernel(const CgplArgs A) {
    constexpr int RPW = LPR <= 32 ? 32 / LPR : 0;
    pdl_wait();                 // predecessor complete and visible ...
    pdl_launch_dependents();    // ... before the next kernel of the chain may start (see gemm_tc05.cu)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.b_l; i += gridDim.x * blockDim.x) {
        // a label outside [0, k) (F.one_hot raises on it in the reference) must not become a row index downstream: the row
        // is kept in range and marked not confident, so it contributes nothing
        const long long y = A.y_l[i];
        const bool ok = y >= 0 && y < A.k;
        A.cls_l[i] = ok ? (int)y : 0;
        A.conf_l[i] = ok && (1.0f >= A.th1);
    }
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int sub = LPR <= 32 ? lane % LPR : (int)threadIdx.x;                       // LPR > 32: one row per block of
    const int row = LPR <= 32 ? warp_global * RPW + lane / (LPR <= 32 ? LPR : 1)      // LPR threads
                              : (int)blockIdx.x;
    // rows beyond the end keep running with a clamped index so that the full-warp shuffles stay converged
    const bool row_ok = row < A.rows;
    const int r = row_ok ? row : A.rows - 1;
    const int k = A.k;

    float ym[2 * NV], yi[2 * NV], yt[2 * NV], tp[2 * NV];
    if constexpr (FAST) {
        load_row_fast<LPR, NV>(static_cast<const float*>(A.y_m), A.ld_y, r, k >> 1, sub, ym);
        load_row_fast<LPR, NV>(static_cast<const float*>(A.y_i), A.ld_y, r, k >> 1, sub, yi);
        load_row_fast<LPR, NV>(static_cast<const float*>(A.y_t), A.ld_y, r, k >> 1, sub, yt);
        load_row_fast<LPR, NV>(A.tl, A.ld_t, r, k >> 1, sub, tp);
    } else {
        load_row<LPR, NV>(A.y_m, A.logit_dtype, A.ld_y, r, k, sub, A.vec_y, ym);
        load_row<LPR, NV>(A.y_i, A.logit_dtype, A.ld_y, r, k, sub, A.vec_y, yi);
        load_row<LPR, NV>(A.y_t, A.logit_dtype, A.ld_y, r, k, sub, A.vec_y, yt);
        load_row<LPR, NV>(A.tl, STIL_F32, A.ld_t, r, k, sub, A.vec_t, tp);   // all four rows in flight together
    }

    // ---- :262-263  top-1 of the three softmaxes
    float pm[2 * NV];  // becomes softmax(y_m) = `prediction` (:279) and the case-3 pseudo label (:273)
#pragma unroll
    for (int j = 0; j < 2 * NV; ++j) pm[j] = ym[j];
    const float sm = softmax_exp<LPR, NV>(pm);
    const float rsm = __frcp_rn(sm);   // p = e * (1/s): within 1 ulp of torch's e / s (same class as exp differences)
#pragma unroll
    for (int j = 0; j < 2 * NV; ++j) pm[j] = __fmul_rn(pm[j], rsm);
    // the same rule for the three heads: first index of the largest logit (see argmax_logits)
    const int top_m = argmax_logits<LPR, NV>(ym, sub, k);
    const int top_i = argmax_logits<LPR, NV>(yi, sub, k);
    const int top_t = argmax_logits<LPR, NV>(yt, sub, k);
    // ---- :264-267 agreement cases
    const bool mi = top_m == top_i, mt = top_m == top_t;
    const bool c1 = mi && mt, c2i = mi && !mt, c2t = mt && !mi;
    const bool c3 = !(c1 || c2i || c2t);

    // ---- :270-274 case-selected softmax of the averaged logits (eager rounding: no FMA contraction)
    float pl[2 * NV];
    if (c3) {
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j) pl[j] = pm[j];
    } else {
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j) {
            float a;
            if (c1)
                a = __fmul_rn(__fadd_rn(__fadd_rn(ym[j], yi[j]), yt[j]), A.third);
            else if (c2i)
                a = __fmul_rn(__fadd_rn(ym[j], yi[j]), 0.5f);
            else
                a = __fmul_rn(__fadd_rn(ym[j], yt[j]), 0.5f);
            pl[j] = a;
        }
        const float rs = __frcp_rn(softmax_exp<LPR, NV>(pl));
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j) pl[j] = __fmul_rn(pl[j], rs);
    }
    // ---- :293-294 teacher prototype probabilities
#pragma unroll
    for (int j = 0; j < 2 * NV; ++j) tp[j] = __fmul_rn(tp[j], A.inv_temperature);
    {
        const float rs = __frcp_rn(softmax_exp<LPR, NV>(tp));
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j) tp[j] = __fmul_rn(tp[j], rs);
    }
    // ---- :276-277 `prediction` given by the caller (distribution alignment) instead of softmax(y_m)
    if (A.pred_in) {
        load_row<LPR, NV>(A.pred_in, STIL_F32, A.ld_pin, r, k, sub, A.vec_pin, pm);
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j)
            if (pm[j] == -INFINITY) pm[j] = 0.f;     // padding slots beyond k
    }
    // ---- :295-298 smoothing mix, max/argmax, threshold
    // padding slots carry probability 0 in pm, pl and tp, strictly below the row maximum (>= 1/k): no index test needed
    float bv = -1.f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 2 * NV; ++j) {
        const int idx = 2 * (sub + LPR * (j >> 1)) + (j & 1);
        const float t = __fmul_rn(A.one_minus_rate, tp[j]);
        pl[j] = __fadd_rn(__fmul_rn(A.rate_pseudo, pl[j]), t);
        pm[j] = __fadd_rn(__fmul_rn(A.rate_pseudo, pm[j]), t);
        if (pm[j] > bv) {
            bv = pm[j];
            bi = idx;
        }
    }
    group_argmax<LPR>(bv, bi);
    if (bi >= k) { bi = 0; bv = __int_as_float(0x7fc00000); }   // NaN row: index stays in range, max_prob = NaN, mask1 = false
    const bool m1 = bv >= A.th1;

    if (!row_ok) return;
    if constexpr (FAST) {
        store_row_fast<LPR, NV>(A.pseudo_label, A.ld_pl, row, k >> 1, sub, pl);
        if (A.prediction) store_row_fast<LPR, NV>(A.prediction, A.ld_pred, row, k >> 1, sub, pm);
    } else {
        store_row<LPR, NV>(A.pseudo_label, A.ld_pl, row, k, sub, A.vec_pl, pl);
        if (A.prediction) store_row<LPR, NV>(A.prediction, A.ld_pred, row, k, sub, A.vec_pred, pm);
    }
    if (sub == 0) {
        if (A.max_prob) A.max_prob[row] = bv;
        A.max_idx[row] = bi;
        A.mask1[row] = m1;
        if (A.case1) A.case1[row] = c1;
        if (A.case2_i) A.case2_i[row] = c2i;
        if (A.case2_t) A.case2_t[row] = c2t;
        if (A.case3) A.case3[row] = c3;
        if (A.top1) {
            A.top1[row] = top_m;
            A.top1[A.rows + row] = top_i;
            A.top1[2LL * A.rows + row] = top_t;
        }
        if (A.cls) A.cls[row] = A.past_start ? bi : 0;
        if (A.conf) A.conf[row] = A.past_start ? (unsigned char)m1 : (unsigned char)(0.0f >= A.th1);
    }
}

// =====================================================================================
// label.max(1) with threshold (utils/prototype_loss.py:31-32, STiLModel.py:204-205)
// =====================================================================================
__global__ void __launch_bounds__(kRowBlock) label_argmax_kernel(const float* __restrict__ label, long long ld,
                                                                 int rows, int k, float th, int* cls,
                                                                 unsigned char* conf, float* max_prob) {
    const int row = blockIdx.x * (kRowBlock / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    bool has_nan = false;
    for (int j = lane; j < k; j += 32) {
        const float x = label[(long long)row * ld + j];
        if (x != x && !has_nan) {  // torch.max propagates the first NaN
            has_nan = true;
            bv = x;
            bi = j;
        }
        if (!has_nan && (x > bv || bi == 0x7fffffff)) {
            bv = x;
            bi = j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool onan = ov != ov, mnan = bv != bv;
        bool take;
        if (onan || mnan)
            take = onan && (!mnan || oi < bi);
        else
            take = ov > bv || (ov == bv && oi < bi);
        if (take) {
            bv = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        cls[row] = bi;
        conf[row] = bv >= th;
        if (max_prob) max_prob[row] = bv;
    }
}

__global__ void labelled_cls_kernel(const long long* y_l, int b_l, float th, int* cls, unsigned char* conf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b_l) {
        cls[i] = (int)y_l[i];
        conf[i] = 1.0f >= th;
    }
}

// four consecutive elements of an f32 / bf16 row as floats (16- / 8-byte load; the caller guarantees alignment)
__device__ __forceinline__ void ld4f(const void* base, int dtype, long long off, float (&x)[4]) {
    if (dtype == STIL_BF16) {
        const uint2 v = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(base) + off);
        x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xffff0000u);
        x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xffff0000u);
    } else {
        const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + off);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
}
__device__ __forceinline__ void st4f(void* base, int dtype, long long off, const float (&x)[4]) {
    if (dtype == STIL_BF16) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(base) + off) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    } else {
        *reinterpret_cast<float4*>(static_cast<float*>(base) + off) = make_float4(x[0], x[1], x[2], x[3]);
    }
}
__device__ __forceinline__ bool rows_vec4(const void* base, int dtype, long long ld) {
    const int esz = dtype == STIL_BF16 ? 2 : 4;
    return (reinterpret_cast<uintptr_t>(base) % (4 * esz)) == 0 && (ld % 4) == 0;
}

// =====================================================================================
// merge GEMM_STATS partials -> LSE, diagonal / picked logit, loss terms (deterministic reduction)
// =====================================================================================
__global__ void __launch_bounds__(kRowBlock) finish_kernel(const __grid_constant__ FinishLaunch L) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grow = blockIdx.x * (blockDim.x >> 5) + warp;
    float term[2] = {0.f, 0.f};
    pdl_wait();
    pdl_launch_dependents();
    if (grow < L.total_rows) {
        int jid = 0;
#pragma unroll
        for (int j = 1; j < kMaxFinishJobs; ++j)
            if (j < L.njobs && grow >= L.job[j].row_begin) jid = j;
        const FinishJob& J = L.job[jid];
        const int i = grow - J.row_begin;
        // every global load of the row is issued before the first reduction (one memory round trip, not five):
        // the first 64 statistics slots and 128 row elements live in registers
        const long long yrow = J.kind == 0 ? (long long)(i + J.y_offset) : (long long)J.cls[i];
        float pm[2], ps[2], xv[4], yv[4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int t = lane + 32 * q;
            pm[q] = t < J.tiles_n ? J.part_max[(long long)t * J.M + i] : -INFINITY;
            ps[q] = t < J.tiles_n ? J.part_sum[(long long)t * J.M + i] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int d = lane + 32 * q;
            xv[q] = d < J.dim ? ld_as_float(J.x, J.x_dtype, (long long)i * J.ldx + d) : 0.f;
            yv[q] = d < J.dim ? ld_as_float(J.y, J.y_dtype, yrow * J.ldy + d) : 0.f;
        }
        const float sxv = J.sx ? J.sx[i] : 1.f, syv = J.sy ? J.sy[i + J.y_offset] : 1.f;
        // merge (max, sum) partials over column tiles
        // slots beyond the first 64: (max, sum) pairs merged per lane in batches of four independent loads
        float em = -INFINITY, es = 0.f;
        for (int t0 = lane + 64; t0 < J.tiles_n; t0 += 128) {
            float bm[4], bs[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + 32 * q;
                bm[q] = t < J.tiles_n ? J.part_max[(long long)t * J.M + i] : -INFINITY;
                bs[q] = t < J.tiles_n ? J.part_sum[(long long)t * J.M + i] : 0.f;
            }
            const float nm = fmaxf(fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3])), em);
            if (nm == -INFINITY) continue;
            float acc = es * expf(em - nm);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += bs[q] * expf(bm[q] - nm);
            em = nm;
            es = acc;
        }
        float m = warp_max(fmaxf(fmaxf(pm[0], pm[1]), em));
        float s = ps[0] * expf(pm[0] - m) + ps[1] * expf(pm[1] - m) + (em == -INFINITY ? 0.f : es * expf(em - m));
        s = warp_sum(s);
        const float lse = m + logf(s);
        // dot product with the partner row
        float dot = xv[0] * yv[0] + xv[1] * yv[1] + xv[2] * yv[2] + xv[3] * yv[3];
        if (J.dim > 128) {
            // wide rows (D = 512 / 2048): 4-element vector loads, four independent pairs in flight per lane
            if ((J.dim & 3) == 0 && rows_vec4(J.x, J.x_dtype, J.ldx) && rows_vec4(J.y, J.y_dtype, J.ldy)) {
                for (int d0 = 128 + 4 * lane; d0 < J.dim; d0 += 512) {
                    float a[4][4], b[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int d = d0 + 128 * u;
                        if (d < J.dim) {
                            ld4f(J.x, J.x_dtype, (long long)i * J.ldx + d, a[u]);
                            ld4f(J.y, J.y_dtype, yrow * J.ldy + d, b[u]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) a[u][q] = b[u][q] = 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int q = 0; q < 4; ++q) dot += a[u][q] * b[u][q];
                }
            } else {
                for (int d = lane + 128; d < J.dim; d += 32)
                    dot += ld_as_float(J.x, J.x_dtype, (long long)i * J.ldx + d) * ld_as_float(J.y, J.y_dtype, yrow * J.ldy + d);
            }
        }
        dot = warp_sum(dot);
        const float z = dot * J.alpha * sxv * syv;
        if (lane == 0) {
            J.lse[i] = lse;
            if (J.kind == 0) {
                term[J.loss_slot] = J.coef * (lse - z);
            } else {
                const float p = expf(z - lse);
                const float c = J.conf[i] ? 1.f : 0.f;
                term[J.loss_slot] = -c * logf(p + 1e-7f) * J.coef;   // utils/prototype_loss.py:28,37-39
                J.w[i] = c * J.coef * p / (p + 1e-7f);
            }
        }
    }
    // block partial (two slots), then the last block adds the partials in order
    __shared__ float sred[2][kRowBlock / 32];
    __shared__ bool is_last;
    if (lane == 0) {
        sred[0][warp] = term[0];
        sred[1][warp] = term[1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            a += sred[0][w];
            b += sred[1][w];
        }
        L.block_partials[2 * blockIdx.x] = a;
        L.block_partials[2 * blockIdx.x + 1] = b;
        __threadfence();
        is_last = atomicAdd(L.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && warp == 0) {
        // deterministic: lane-strided partial sums in block order, then a fixed shuffle tree, per loss slot
        __threadfence();
        const float2* bp = reinterpret_cast<const float2*>(L.block_partials);
        // batches of 8 independent L2 loads per lane (one round trip per 256 blocks), added in a fixed order
        float sum2[2] = {0.f, 0.f};
        for (unsigned int b0 = 0; b0 < gridDim.x; b0 += 256) {
            float2 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned int b = b0 + lane + 32 * q;
                v[q] = b < gridDim.x ? __ldcg(bp + b) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) { sum2[0] += v[q].x; sum2[1] += v[q].y; }
        }
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
            float sum = warp_sum(sum2[slot]);
            bool used = false;
            for (int j = 0; j < L.njobs; ++j) used |= L.job[j].loss_slot == slot;
            if (used && lane == 0) L.out_loss[slot] = sum;
            if (slot == 0 && lane < L.ll_world)
                *L.ll_out[lane] = ((unsigned long long)(unsigned int)(*L.ll_tag) << 32) | __float_as_uint(sum);
        }
        if (lane == 0) *L.ticket = 0u;
    }
}

// =====================================================================================
// backward of F.normalize (or plain cast) on the GEMM2 output
// =====================================================================================
__global__ void __launch_bounds__(kRowBlock) grad_finish_kernel(const __grid_constant__ GradFinishLaunch L) {
    pdl_wait();
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grow = blockIdx.x * (blockDim.x >> 5) + warp;
    if (grow >= L.total_rows) return;
    int jid = 0;
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (j < L.njobs && grow >= L.job[j].row_begin) jid = j;
    const GradFinishJob& J = L.job[jid];
    const int i = grow - J.row_begin;
    const float* g0 = J.g + (long long)i * J.dim;
    const int ns = J.nslices > 1 ? J.nslices : 1;
    auto gsum = [&](int d) {      // slices of a split contraction, added in index order (deterministic)
        // eight independent (predicated) loads per batch, then the adds: a `v += load` loop serialises one memory round
        // trip per slice
        if (ns == 1) return g0[d];
        float v = 0.f;
        for (int k0 = 0; k0 < ns; k0 += 8) {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = (k0 + j < ns) ? __ldcg(g0 + (long long)(k0 + j) * J.slice_stride + d) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) v += t[j];
        }
        return v;
    };
    // wide rows (D = 512 / 2048; one warp per row): 4-element vector loads with four independent chunks in flight per lane —
    // the scalar form below serialises one memory round trip per 32 elements (215 us for 2 x 4096 x 2048, 28 % of a CLIPLoss
    // forward + backward at that shape)
    const bool wide = J.dim > 128 && (J.dim & 3) == 0 && (reinterpret_cast<uintptr_t>(J.g) & 15) == 0 && (J.slice_stride & 3) == 0 &&
                      rows_vec4(J.dx, J.dx_dtype, J.ld_dx) && (!J.sx || rows_vec4(J.x, J.x_dtype, J.ldx));
    if (wide) {
        auto gsum4 = [&](int d, float (&v)[4]) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(g0 + d));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            for (int k = 1; k < ns; ++k) {
                const float4 b = __ldcg(reinterpret_cast<const float4*>(g0 + (long long)k * J.slice_stride + d));
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
            }
        };
        const float sx = J.sx ? J.sx[i] : 0.f;
        float dot = 0.f;
        if (J.sx) {
            for (int d0 = 4 * lane; d0 < J.dim; d0 += 512) {
                float gv[4][4], xv[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int d = d0 + 128 * u;
                    if (d < J.dim) {
                        gsum4(d, gv[u]);
                        ld4f(J.x, J.x_dtype, (long long)i * J.ldx + d, xv[u]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) gv[u][q] = xv[u][q] = 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int q = 0; q < 4; ++q) dot += sx * xv[u][q] * gv[u][q];
            }
            dot = warp_sum(dot);
        }
        for (int d0 = 4 * lane; d0 < J.dim; d0 += 512) {
            float gv[4][4], xv[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = d0 + 128 * u;
                if (d < J.dim) {
                    gsum4(d, gv[u]);
                    if (J.sx) ld4f(J.x, J.x_dtype, (long long)i * J.ldx + d, xv[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = d0 + 128 * u;
                if (d < J.dim) {
                    float o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) o[q] = J.sx ? sx * (gv[u][q] - sx * xv[u][q] * dot) : gv[u][q];
                    st4f(J.dx, J.dx_dtype, (long long)i * J.ld_dx + d, o);
                }
            }
        }
        return;
    }
    if (J.sx) {
        const float sx = J.sx[i];
        // dim <= 128: the row's gradient stays in registers between the dot product and the projection
        float gv[4];
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int d = lane + 32 * q;
            gv[q] = d < J.dim ? gsum(d) : 0.f;
            if (d < J.dim) dot += sx * ld_as_float(J.x, J.x_dtype, (long long)i * J.ldx + d) * gv[q];
        }
        for (int d = lane + 128; d < J.dim; d += 32) dot += sx * ld_as_float(J.x, J.x_dtype, (long long)i * J.ldx + d) * gsum(d);
        dot = warp_sum(dot);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int d = lane + 32 * q;
            if (d < J.dim) {
                const float xh = sx * ld_as_float(J.x, J.x_dtype, (long long)i * J.ldx + d);
                st_from_float(J.dx, J.dx_dtype, (long long)i * J.ld_dx + d, sx * (gv[q] - xh * dot));
            }
        }
        for (int d = lane + 128; d < J.dim; d += 32) {
            const float xh = sx * ld_as_float(J.x, J.x_dtype, (long long)i * J.ldx + d);
            st_from_float(J.dx, J.dx_dtype, (long long)i * J.ld_dx + d, sx * (gsum(d) - xh * dot));
        }
    } else {
        for (int d = lane; d < J.dim; d += 32) st_from_float(J.dx, J.dx_dtype, (long long)i * J.ld_dx + d, gsum(d));
    }
}

// =====================================================================================
// segmented per-class sums (STiLModel.py:199-226, 380-381): one warp per class, rows in index order
// =====================================================================================
constexpr int kAccChunk = 2048;   // rows of (cls, conf) codes staged in shared memory per pass

__device__ __forceinline__ void load4_as_float(const void* feat, int dtype, long long off, int d, int dim, bool vec,
                                               float (&x)[4]) {
    if (vec) {
        if (dtype == STIL_BF16) {
            const uint2 v = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(feat) + off + d);
            x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xffff0000u);
            x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xffff0000u);
        } else {
            const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(feat) + off + d);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = (d + j < dim) ? ld_as_float(feat, dtype, off + d + j) : 0.f;
    }
}

// One block per class; the block's warps split the rows (contiguous ranges), each warp adds its matching rows in
// index order, and the per-warp partials are combined in warp order — the result does not depend on scheduling.
__global__ void __launch_bounds__(kRowBlock) proto_accumulate_kernel(const void* __restrict__ feat, int dtype, int rows,
                                                                     int dim, long long ld,
                                                                     const int* __restrict__ cls,
                                                                     const unsigned char* __restrict__ conf, int b_l,
                                                                     float repeat_ratio, int k, float* class_sum,
                                                                     float* class_count, float* psum, float* pcount) {
    __shared__ int codes[kAccChunk];                       // class of a confident row, -1 otherwise
    __shared__ float part[2][kRowBlock / 32][128];         // [labelled | unlabelled][warp][dim slab]
    __shared__ float cnt[2][kRowBlock / 32];
    const int c = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarp = blockDim.x >> 5;
    const bool vec = (dim % 4 == 0) && (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(feat) % 16 == 0);
    for (int d0 = 0; d0 < dim; d0 += 128) {
        float al[4] = {0.f, 0.f, 0.f, 0.f}, au[4] = {0.f, 0.f, 0.f, 0.f};
        float nl = 0.f, nu = 0.f;
        const int d = d0 + lane * 4;
        for (int base = 0; base < rows; base += kAccChunk) {
            const int nrow = min(kAccChunk, rows - base);
            __syncthreads();
            for (int i = threadIdx.x; i < nrow; i += blockDim.x) codes[i] = conf[base + i] ? cls[base + i] : -1;
            __syncthreads();
            // this warp's contiguous share of the chunk, in units of 32 rows
            const int groups = (nrow + 31) / 32;
            const int per = (groups + nwarp - 1) / nwarp;
            for (int gi = warp * per; gi < min(groups, (warp + 1) * per); ++gi) {
                const int r0 = gi * 32;
                const bool hit = (r0 + lane < nrow) && codes[r0 + lane] == c;
                unsigned int ballot = __ballot_sync(0xffffffffu, hit);
                while (ballot) {
                    // up to four matching rows in flight, added in ascending row order
                    int idx[4];
                    int n = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (ballot) {
                            idx[j] = base + r0 + __ffs(ballot) - 1;
                            ballot &= ballot - 1;
                            n = j + 1;
                        }
                    float x[4][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < n && d < dim) load4_as_float(feat, dtype, (long long)idx[j] * ld, d, dim, vec, x[j]);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < n) {
                            const bool lab = idx[j] < b_l;
                            if (lab) nl += 1.f; else nu += 1.f;
                            if (d < dim) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (lab) al[q] += x[j][q]; else au[q] += x[j][q];
                                }
                            }
                        }
                }
            }
        }
        // combine the warps' partials in warp order
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            part[0][warp][lane * 4 + q] = al[q];
            part[1][warp][lane * 4 + q] = au[q];
        }
        if (lane == 0) {
            cnt[0][warp] = nl;
            cnt[1][warp] = nu;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 128; t += blockDim.x) {
            const int dd = d0 + t;
            if (dd < dim) {
                float sl = 0.f, su = 0.f;
                for (int w = 0; w < nwarp; ++w) {
                    sl += part[0][w][t];
                    su += part[1][w][t];
                }
                const float v = __fadd_rn(__fdiv_rn(sl, repeat_ratio), su);         // :224
                class_sum[(long long)c * dim + dd] = v;
                if (psum) psum[(long long)c * dim + dd] += v;                        // :380
            }
        }
        if (d0 == 0 && threadIdx.x == 0) {
            float sl = 0.f, su = 0.f;
            for (int w = 0; w < nwarp; ++w) {
                sl += cnt[0][w];
                su += cnt[1][w];
            }
            const float v = __fadd_rn(__fdiv_rn(sl, repeat_ratio), su);             // :225
            class_count[c] = v;
            if (pcount) pcount[c] += v;                                              // :381
        }
    }
}

__global__ void proto_add_kernel(const float* class_sum, const float* class_count, int k, int dim, float* psum,
                                 float* pcount) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)k * dim) psum[i] += class_sum[i];
    if (i < k) pcount[i] += class_count[i];
}

__global__ void proto_add_gathered_kernel(const float* parts, int world, long long slot, int k, int dim, float* class_sum,
                                          float* class_count, float* psum, float* pcount,
                                          const unsigned long long* wait_flags, const unsigned long long* wait_target,
                                          const unsigned long long* loss_ll, const unsigned long long* loss_tag,
                                          float* loss_out) {
    // fused schedule of the data-parallel head: the partials were pushed by the peers (csrc/p2p.cu) — wait for every
    // rank's arrival counter here, at the consumer, instead of in a kernel of its own
    if (wait_flags) {
        if (threadIdx.x == 0) {
            // one polling thread per block, acquire loads with back-off
            const unsigned long long target = *wait_target;
            for (int w = 0; w < world; ++w) {
                unsigned long long v, spins = 0, t0 = 0;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(wait_flags + w) : "memory");
                    if (v >= target) break;
                    peer_wait_check(spins, t0);
                    __nanosleep(64);
                }
            }
        }
        __syncthreads();
    }
    if (loss_ll && blockIdx.x == 0 && threadIdx.x == 0) {
        // global InfoNCE loss = sum of the ranks' partials (LL words), in rank order
        const unsigned int want = (unsigned int)(*loss_tag);
        float s = 0.f;
        for (int w = 0; w < world; ++w) {
            unsigned long long v, spins = 0, t0 = 0;
            for (;;) {
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(loss_ll + w) : "memory");
                if ((unsigned int)(v >> 32) == want) break;
                peer_wait_check(spins, t0);
                __nanosleep(64);
            }
            s += __uint_as_float((unsigned int)v);
        }
        *loss_out = s;
    }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nsum = (long long)k * dim;
    if (i < nsum + k) {
        float s = 0.f;
        for (int w = 0; w < world; ++w) s += __ldcg(parts + (long long)w * slot + i);   // rank order: deterministic
        if (i < nsum) {
            class_sum[i] = s;
            psum[i] += s;
        } else {
            class_count[i - nsum] = s;
            pcount[i - nsum] += s;
        }
    }
}

__global__ void proto_finalize_kernel(float* prototypes, float* psum, float* pcount, int dim, int* empty) {
    const int c = blockIdx.x;
    const float cnt = pcount[c];
    __syncthreads();
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        prototypes[(long long)c * dim + d] = psum[(long long)c * dim + d] / cnt;   // STiLModel.py:413
        psum[(long long)c * dim + d] = 0.f;                                         // :414
    }
    if (threadIdx.x == 0) {
        pcount[c] = 0.f;                                                            // :415
        if (cnt < 1.f) atomicAdd(empty, 1);                                         // :411-412 (assert -> flag)
    }
}

// =====================================================================================
// masked soft-target CE, forward + gradient (STiLModel.py:301-303)
// =====================================================================================
struct SoftCeArgs {
    const void* y[3];
    int logit_dtype;
    long long ld_y;
    const float* pl;
    long long ld_pl;
    const unsigned char *mask1, *case1, *case2_i, *case2_t, *case3, *mask_random;
    int rows, k;
    float* dy[3];
    long long ld_g;
    float grad_scale;
    float* block_partials;  // [blocks*3]
    unsigned int* ticket;
    float* losses;          // [3]
    int vec_y, vec_pl, vec_g;
};

// block partial of the three losses, then the last block adds the partials in order (deterministic)
__device__ __forceinline__ void softce_block_finish(float (&lossv)[3], const SoftCeArgs& A) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 0; h < 3; ++h) lossv[h] = warp_sum(lossv[h]);     // the row leaders' terms of this warp
    __shared__ float sred[3][kRowBlock / 32];
    __shared__ bool is_last;
    if (lane == 0)
        for (int h = 0; h < 3; ++h) sred[h][warp] = lossv[h];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int h = 0; h < 3; ++h) {
            float a = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sred[h][w];
            A.block_partials[3 * blockIdx.x + h] = a;
        }
        __threadfence();
        is_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && warp == 0) {
        // deterministic: lane-strided partial sums in block order, then a fixed shuffle tree, per loss
        __threadfence();
        const float* bp = A.block_partials;
        float sum3[3] = {0.f, 0.f, 0.f};
        for (unsigned int b0 = 0; b0 < gridDim.x; b0 += 128) {      // four independent L2 loads per lane and loss per batch
            float v[4][3];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned int b = b0 + lane + 32 * q;
#pragma unroll
                for (int h = 0; h < 3; ++h) v[q][h] = b < gridDim.x ? __ldcg(bp + 3 * b + h) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int h = 0; h < 3; ++h) sum3[h] += v[q][h];
        }
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const float sum = warp_sum(sum3[h]);
            if (lane == 0) A.losses[h] = sum / (float)A.rows;   // .mean() over B_u
        }
        if (lane == 0) *A.ticket = 0u;
    }
}

// K = 2 (cardiac: binary CAD / infarction): a row is 8 bytes, so one thread takes TWO rows with 16-byte loads / stores and
// 2-byte flag loads — the generic one-thread-per-row form issues 13 memory instructions for 62 bytes and is LSU / issue bound
// (ncu, profiles/r2_ncu_rows2_before.txt: 17 % of the DRAM peak at 64 % issue utilisation)
__global__ void __launch_bounds__(kRowBlock) masked_softce_k2_kernel(const SoftCeArgs A) {
    float lossv[3] = {0.f, 0.f, 0.f};
    const float inv_rows = 1.0f / (float)A.rows;
    const int pairs = A.rows >> 1;
    const float4* pl4 = reinterpret_cast<const float4*>(A.pl);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += gridDim.x * blockDim.x) {
        const float4 pl = pl4[p];
        float4 y[3];
#pragma unroll
        for (int h = 0; h < 3; ++h) y[h] = reinterpret_cast<const float4*>(A.y[h])[p];
        const uchar2 m1 = reinterpret_cast<const uchar2*>(A.mask1)[p], c1 = reinterpret_cast<const uchar2*>(A.case1)[p];
        const uchar2 c2i = reinterpret_cast<const uchar2*>(A.case2_i)[p], c2t = reinterpret_cast<const uchar2*>(A.case2_t)[p];
        const uchar2 c3 = reinterpret_cast<const uchar2*>(A.case3)[p], mr = reinterpret_cast<const uchar2*>(A.mask_random)[p];
        float4 dy[3];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float p0 = r ? pl.z : pl.x, p1 = r ? pl.w : pl.y;
            const float fm1 = (r ? m1.y : m1.x) ? 1.f : 0.f, fc1 = (r ? c1.y : c1.x) ? 1.f : 0.f;
            const float fc2i = (r ? c2i.y : c2i.x) ? 1.f : 0.f, fc2t = (r ? c2t.y : c2t.x) ? 1.f : 0.f;
            const float fc3 = (r ? c3.y : c3.x) ? 1.f : 0.f, fmr = (r ? mr.y : mr.x) ? 1.f : 0.f;
            const float wgt[3] = {fm1 * fc1, fm1 * (fc1 + fc2t + fc3 * fmr), fm1 * (fc1 + fc2i + fc3 * (1.f - fmr))};
            const float spl = p0 + p1;
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const float y0 = r ? y[h].z : y[h].x, y1 = r ? y[h].w : y[h].y;
                const float m = fmaxf(y0, y1);
                const float ml = m * 1.4426950408889634f;
                const float e0 = exp_fast_shift(y0, ml), e1 = exp_fast_shift(y1, ml);
                const float s = e0 + e1;
                const float lse = m + log_fast(s);
                lossv[h] += (lse * spl - (p0 * y0 + p1 * y1)) * wgt[h];
                const float gs = A.grad_scale * wgt[h] * inv_rows;
                const float ps = gs * spl * __frcp_rn(s);
                const float d0 = fmaf(e0, ps, -gs * p0), d1 = fmaf(e1, ps, -gs * p1);
                if (r) { dy[h].z = d0; dy[h].w = d1; } else { dy[h].x = d0; dy[h].y = d1; }
            }
        }
#pragma unroll
        for (int h = 0; h < 3; ++h)
            if (A.dy[h]) reinterpret_cast<float4*>(A.dy[h])[p] = dy[h];
    }
    // odd row count: the last row alone
    if ((A.rows & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int row = A.rows - 1;
        const float p0 = A.pl[2 * row], p1 = A.pl[2 * row + 1];
        const float fm1 = A.mask1[row] ? 1.f : 0.f, fc1 = A.case1[row] ? 1.f : 0.f, fc2i = A.case2_i[row] ? 1.f : 0.f;
        const float fc2t = A.case2_t[row] ? 1.f : 0.f, fc3 = A.case3[row] ? 1.f : 0.f, fmr = A.mask_random[row] ? 1.f : 0.f;
        const float wgt[3] = {fm1 * fc1, fm1 * (fc1 + fc2t + fc3 * fmr), fm1 * (fc1 + fc2i + fc3 * (1.f - fmr))};
        const float spl = p0 + p1;
        for (int h = 0; h < 3; ++h) {
            const float* yy = static_cast<const float*>(A.y[h]) + 2 * row;
            const float m = fmaxf(yy[0], yy[1]), ml = m * 1.4426950408889634f;
            const float e0 = exp_fast_shift(yy[0], ml), e1 = exp_fast_shift(yy[1], ml), s = e0 + e1;
            const float lse = m + log_fast(s);
            lossv[h] += (lse * spl - (p0 * yy[0] + p1 * yy[1])) * wgt[h];
            const float gs = A.grad_scale * wgt[h] * inv_rows, ps = gs * spl * __frcp_rn(s);
            if (A.dy[h]) { A.dy[h][2 * row] = fmaf(e0, ps, -gs * p0); A.dy[h][2 * row + 1] = fmaf(e1, ps, -gs * p1); }
        }
    }
    softce_block_finish(lossv, A);
}

// LPR lanes per row (32 / LPR rows per warp for small K, like cgpl_pgls_kernel), NV float2 items per lane
template <int LPR, int NV, bool FAST>
__global__ void __launch_bounds__(kRowBlock, (NV <= 5 ? 3 : 1)) masked_softce_kernel(const SoftCeArgs A) {
    constexpr int RPW = 32 / LPR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    float lossv[3] = {0.f, 0.f, 0.f};
    const int k = A.k;
    const float inv_rows = 1.0f / (float)A.rows;
    // grid-stride over row groups: the grid is capped so that the final (serial, deterministic) reduction stays short
    const int wpb = blockDim.x >> 5;
    for (int row0 = (blockIdx.x * wpb + warp) * RPW; row0 < A.rows; row0 += gridDim.x * wpb * RPW) {
        const int row = row0 + lane / LPR;
        const bool ok = row < A.rows;
        const int r = ok ? row : A.rows - 1;      // lanes past the end keep running on a clamped row (full-warp shuffles)
        // all four rows of this sample in flight together
        float pl[2 * NV], y[3][2 * NV];
        if constexpr (FAST) {
            load_row_fast<LPR, NV>(A.pl, A.ld_pl, r, k >> 1, sub, pl);
#pragma unroll
            for (int h = 0; h < 3; ++h) load_row_fast<LPR, NV>(static_cast<const float*>(A.y[h]), A.ld_y, r, k >> 1, sub, y[h]);
        } else {
            load_row<LPR, NV>(A.pl, STIL_F32, A.ld_pl, r, k, sub, A.vec_pl, pl);
#pragma unroll
            for (int h = 0; h < 3; ++h) load_row<LPR, NV>(A.y[h], A.logit_dtype, A.ld_y, r, k, sub, A.vec_y, y[h]);
        }
        const float m1 = (ok && A.mask1[r]) ? 1.f : 0.f;
        const float c1 = A.case1[r] ? 1.f : 0.f, c2i = A.case2_i[r] ? 1.f : 0.f;
        const float c2t = A.case2_t[r] ? 1.f : 0.f, c3 = A.case3[r] ? 1.f : 0.f;
        const float mr = A.mask_random[r] ? 1.f : 0.f;
        const float wgt[3] = {m1 * c1, m1 * (c1 + c2t + c3 * mr), m1 * (c1 + c2i + c3 * (1.f - mr))};
        float spl = 0.f;
#pragma unroll
        for (int it = 0; it < NV; ++it)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (2 * (sub + LPR * it) + h >= k) pl[2 * it + h] = 0.f;   // padding was -inf
                spl += pl[2 * it + h];
            }
        spl = group_sum<LPR>(spl);
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            // one warp = one row (LPR == 32): a zero weight skips the softmax (warp-uniform branch); several rows per warp:
            // no branch, a zero weight zeroes the loss term and the gradient
            float dy[2 * NV];
            if (LPR < 32 || wgt[h] != 0.f) {
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 2 * NV; ++j) m = fmaxf(m, y[h][j]);
                m = group_max<LPR>(m);
                float s = 0.f, py = 0.f;
                const float ml = m * 1.4426950408889634f;
#pragma unroll
                for (int j = 0; j < 2 * NV; ++j) {
                    if (pl[j] != 0.f) py += pl[j] * y[h][j];
                    y[h][j] = exp_fast_shift(y[h][j], ml);        // e = exp(y - max) replaces the logit; exp(-inf) = 0 for the padding
                    s += y[h][j];
                }
                s = group_sum<LPR>(s);
                py = group_sum<LPR>(py);
                const float lse = m + log_fast(s);
                if (sub == 0 && wgt[h] != 0.f) lossv[h] += (lse * spl - py) * wgt[h];   // -sum_k pl_k log_softmax(y)_k, weighted
                const float gs = A.grad_scale * wgt[h] * inv_rows;
                const float ps = gs * spl * __frcp_rn(s);         // softmax = e / s
#pragma unroll
                for (int j = 0; j < 2 * NV; ++j) dy[j] = fmaf(y[h][j], ps, -gs * pl[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 2 * NV; ++j) dy[j] = 0.f;
            }
            if (A.dy[h] && ok) {
                if constexpr (FAST) store_row_fast<LPR, NV>(A.dy[h], A.ld_g, row, k >> 1, sub, dy);
                else store_row<LPR, NV>(A.dy[h], A.ld_g, row, k, sub, A.vec_g, dy);
            }
        }
    }
    softce_block_finish(lossv, A);
}

// =====================================================================================
// distribution alignment (STiLModel.py:171-180)
// =====================================================================================
// mean over rows of each column; one thread per column inside a 32-column slab, rows split over 8 groups and
// combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) da_batch_mean_kernel(const float* __restrict__ probs, long long ld, int rows, int k,
                                                            float* __restrict__ mean) {
    __shared__ float part[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int grp = threadIdx.x >> 5;
    float s = 0.f;
    if (col < k)
        for (int r = grp; r < rows; r += 8) s += probs[(long long)r * ld + col];
    part[grp][threadIdx.x & 31] = s;
    __syncthreads();
    if (grp == 0 && col < k) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += part[g][threadIdx.x];
        mean[col] = t / (float)rows;                                   // probs.mean(0), :173
    }
}
// column sums of a [rows, k] f32 matrix (bias gradient of a Linear layer): same fixed-order reduction as the batch mean
__global__ void __launch_bounds__(256) col_sum_kernel(const float* __restrict__ x, long long ld, int rows, int k, float* __restrict__ out) {
    __shared__ float part[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int grp = threadIdx.x >> 5;
    float s = 0.f;
    if (col < k)
        for (int r = grp; r < rows; r += 8) s += x[(long long)r * ld + col];
    part[grp][threadIdx.x & 31] = s;
    __syncthreads();
    if (grp == 0 && col < k) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += part[g][threadIdx.x];
        out[col] = t;
    }
}
// queue[ptr] = batch_mean (already averaged over ranks); ptr = (ptr+1) % len; qmean = queue.mean(0) over ALL rows
__global__ void __launch_bounds__(256) da_update_kernel(const float* __restrict__ batch_mean, float* da_queue, int da_len,
                                                        int k, long long* da_ptr, float* qmean) {
    const int ptr = (int)(*da_ptr);
    for (int c = threadIdx.x; c < k; c += blockDim.x) da_queue[(long long)ptr * k + c] = batch_mean[c];   // :176
    __syncthreads();
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < da_len; ++r) s += da_queue[(long long)r * k + c];
        qmean[c] = s / (float)da_len;                                                                   // :178
    }
    __syncthreads();
    if (threadIdx.x == 0) *da_ptr = (ptr + 1) % da_len;                                                  // :177
}
// probs / qmean, rows renormalised (:178-179); one warp per row
__global__ void __launch_bounds__(kRowBlock) da_apply_kernel(const float* __restrict__ probs, long long ld, int rows, int k,
                                                             const float* __restrict__ qmean, float* out, long long ld_out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int c = lane; c < k; c += 32) s += __fdiv_rn(probs[(long long)row * ld + c], qmean[c]);
    s = warp_sum(s);
    for (int c = lane; c < k; c += 32)
        out[(long long)row * ld_out + c] = __fdiv_rn(__fdiv_rn(probs[(long long)row * ld + c], qmean[c]), s);
}

// torch.softmax(y, dim=1) of the teacher logits feeding distribution alignment (STiLModel.py:277): one warp per row, the
// row stays in registers for k <= 1024 (re-read otherwise); IEEE exp and division
__global__ void __launch_bounds__(kRowBlock) softmax_rows_kernel(const void* __restrict__ y, int dtype, long long ld, int rows,
                                                                 int k, float* __restrict__ out, long long ld_out) {
    pdl_wait();
    pdl_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float m = -INFINITY;
    for (int c = lane; c < k; c += 32) m = fmaxf(m, ld_as_float(y, dtype, (long long)row * ld + c));
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < k; c += 32) s += expf(ld_as_float(y, dtype, (long long)row * ld + c) - m);
    s = warp_sum(s);
    for (int c = lane; c < k; c += 32)
        out[(long long)row * ld_out + c] = __fdiv_rn(expf(ld_as_float(y, dtype, (long long)row * ld + c) - m), s);
}

// =====================================================================================
// SimMatch bank block on materialised logits (simmatch_model.py:268-286): one block per unlabelled row
// =====================================================================================
//   T = softmax(zt/tt), F_j = p[y_j], T' = T∘F / sum(T∘F), A_c = sum_{j: y_j=c} T_j, p' = cs*p + (1-cs)*A,
//   S = softmax(zs/st), loss_in = -sum_j T'_j log S_j = lse_s - (sum_c p_c Q_c)/den with Q_c = sum_{j in c} T_j zs_j/st
//   and den = sum_c p_c A_c;   G_j = d loss_in / d zs_j (before the 1/st of the logits) = S_j - T'_j
constexpr int kSimBlock = 256;
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = is_max ? -INFINITY : 0.f;
    for (int i = 0; i < kSimBlock / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

__global__ void __launch_bounds__(kSimBlock) simmatch_rows_kernel(const float* __restrict__ zt, const float* __restrict__ zs,
                                                                  long long ldz, const long long* __restrict__ labels,
                                                                  int k_bank, const float* __restrict__ p_orig, int C,
                                                                  float inv_tt, float inv_st, float c_smooth, float* p_out,
                                                                  float* loss_in, __nv_bfloat16* gop, long long ld_g,
                                                                  int g_nseg) {
    extern __shared__ float sm[];   // p[C] | A[C] | Q[C] | red[8]
    float* sp = sm;
    float* sA = sm + C;
    float* sQ = sm + 2 * C;
    float* red = sm + 3 * C;
    const int row = blockIdx.x;
    const float* rt = zt + (long long)row * ldz;
    const float* rs = zs + (long long)row * ldz;
    for (int c = threadIdx.x; c < C; c += kSimBlock) {
        sp[c] = p_orig[(long long)row * C + c];
        sA[c] = 0.f;
        sQ[c] = 0.f;
    }
    float mt = -INFINITY, ms = -INFINITY;
    for (int j = threadIdx.x; j < k_bank; j += kSimBlock) {
        mt = fmaxf(mt, rt[j] * inv_tt);
        ms = fmaxf(ms, rs[j] * inv_st);
    }
    mt = block_reduce(mt, red, true);
    ms = block_reduce(ms, red, true);
    float st_ = 0.f, ss_ = 0.f;
    for (int j = threadIdx.x; j < k_bank; j += kSimBlock) {
        st_ += expf(rt[j] * inv_tt - mt);
        ss_ += expf(rs[j] * inv_st - ms);
    }
    st_ = block_reduce(st_, red, false);
    ss_ = block_reduce(ss_, red, false);
    const float lse_t = mt + logf(st_), lse_s = ms + logf(ss_);
    __syncthreads();
    for (int j = threadIdx.x; j < k_bank; j += kSimBlock) {
        const float T = expf(rt[j] * inv_tt - lse_t);
        const int y = (int)labels[j];
        atomicAdd(&sA[y], T);                              // :276-279 scatter_add (order not reproducible, like the reference)
        atomicAdd(&sQ[y], T * (rs[j] * inv_st));
    }
    __syncthreads();
    float den = 0.f, num = 0.f;
    for (int c = threadIdx.x; c < C; c += kSimBlock) {
        den += sp[c] * sA[c];
        num += sp[c] * sQ[c];
        p_out[(long long)row * C + c] = c_smooth < 1.f ? sp[c] * c_smooth + sA[c] * (1.f - c_smooth) : sp[c];   // :280
    }
    den = block_reduce(den, red, false);
    num = block_reduce(num, red, false);
    if (threadIdx.x == 0) loss_in[row] = lse_s - num / den;                                                       // :286
    if (gop) {
        const float inv_den = 1.f / den;
        __nv_bfloat16* gh = gop + (long long)row * g_nseg * ld_g;
        for (int j = threadIdx.x; j < k_bank; j += kSimBlock) {
            const float T = expf(rt[j] * inv_tt - lse_t);
            const float S = expf(rs[j] * inv_st - lse_s);
            const float g = (S - T * sp[(int)labels[j]] * inv_den) * inv_st;
            const __nv_bfloat16 h = __float2bfloat16_rn(g);
            gh[j] = h;
            if (g_nseg > 1) gh[ld_g + j] = __float2bfloat16_rn(g - __bfloat162float(h));
        }
    }
}

// ---- column-sharded bank: the same block with a FIXED shift instead of the row max, so that the
// statistics of the shards are additive.  Features and bank columns are unit vectors (simmatch_model.py:68-69, :251-252),
// hence z <= 1 and e = exp((z - 1)/T) lies in [e^{-2/T}, 1]: no overflow, and for the reference temperatures (0.1) no
// underflow that matters (e^{-20} = 2e-9 against a row sum >= 1 entry of order 1... K_b entries of order e^{-10}).
//   stats[row] = [ sum_j e_t | sum_j e_s | sum_j e_t p[y_j] z_s/st | A_c = sum_{j in c} e_t  (c < C) ]      one pass over j
// c1 = log2(e)/T: e = 2^(z*c1 - c1), one FFMA + one clamp + one ex2.approx (relative error 2^-22, identical in the
// statistics and the gradient pass)
__device__ __forceinline__ float shifted_exp(float z, float c1) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fminf(fmaf(z, c1, -c1), 86.f)));
    return y;
}
// The per-class sums A_c are accumulated in 36-bit FIXED POINT, as two native 32-bit shared-memory atomics per column
// (hi = bits 18.., lo = bits 0..17 of e * 2^36).  A float atomicAdd on shared memory is a compare-and-swap loop
// (ATOMS.CAST.SPIN): with 286 classes and 32 random labels per warp instruction it retried so often that the kernel ran at
// 25 % of the HBM rate with half of its stall samples on the loop (profiles/r2_ncu_bank_before.txt; replicating the accumulators 4-8x per block to spread the
// remaining bank conflicts was tried and bought nothing: the pass is latency-bound).  Integer sums are
// also order-independent: the statistics of a shard are bit-reproducible.  e <= 1 (+ bf16 rounding of unit vectors), at
// most kSimFxCols columns per block: hi <= 2^18 * 8192 * 1.x < 2^32, lo < 2^18 * 8192 = 2^31.
constexpr float kSimFxScale = 68719476736.f;   // 2^36: absolute resolution 1.5e-11 per column
constexpr int kSimFxLoBits = 18;
constexpr int kSimFxCols = 8192;
// both logit rows and the int64 labels can be read 16 bytes at a time from any column that is a multiple of 4
__device__ __forceinline__ bool sim_vec4(const float* zt, const float* zs, long long ldz, const long long* labels) {
    return (ldz & 3) == 0 && ((reinterpret_cast<uintptr_t>(zt) | reinterpret_cast<uintptr_t>(zs) | reinterpret_cast<uintptr_t>(labels)) & 15) == 0;
}

__global__ void __launch_bounds__(kSimBlock) simmatch_shard_stats_kernel(const float* __restrict__ zt, const float* __restrict__ zs,
                                                                         long long ldz, const long long* __restrict__ labels,
                                                                         int k_shard, const float* __restrict__ p_all, int C,
                                                                         float inv_tt, float inv_st, float* __restrict__ stats,
                                                                         long long chunk_stride) {
    extern __shared__ float sm[];   // p[C] | A hi[C] | A lo[C] | red[8]
    float* sp = sm;
    unsigned int* sHi = reinterpret_cast<unsigned int*>(sm + C);
    unsigned int* sLo = sHi + C;
    float* red = sm + 3 * C;
    const int row = blockIdx.x;
    // a row is cut into gridDim.y column chunks (one block each) so that a few hundred rows still fill the chip; chunk c
    // writes its partial statistics to stats + c * chunk_stride, simmatch_shard_reduce_kernel adds the chunks in order
    const int cw = ((k_shard + gridDim.y - 1) / gridDim.y + 3) & ~3;
    const int j0 = min(k_shard, (int)blockIdx.y * cw), j1 = min(k_shard, j0 + cw);
    stats += (long long)blockIdx.y * chunk_stride;
    const float* rt = zt + (long long)row * ldz;
    const float* rs = zs + (long long)row * ldz;
    for (int c = threadIdx.x; c < C; c += kSimBlock) {
        sp[c] = p_all[(long long)row * C + c];
        sHi[c] = 0u;
        sLo[c] = 0u;
    }
    __syncthreads();
    const float ct = inv_tt * 1.4426950408889634f, cs = inv_st * 1.4426950408889634f;
    float st_ = 0.f, ss_ = 0.f, num = 0.f;
    auto column = [&](float z_t, float z_s, int y) {
        const float et = shifted_exp(z_t, ct);
        st_ += et;
        ss_ += shifted_exp(z_s, cs);
        num = fmaf(et * sp[y], z_s, num);           // the 1/st of z_s/st is applied once, after the reduction
        const unsigned long long v = __float2ull_rn(fminf(et, 1.5f) * kSimFxScale);
        atomicAdd(&sHi[y], (unsigned int)(v >> kSimFxLoBits));
        atomicAdd(&sLo[y], (unsigned int)v & ((1u << kSimFxLoBits) - 1u));
    };
    // four columns per thread per trip (16-byte loads of both logit rows and of the int64 labels, two trips in flight);
    // chunk starts are multiples of 4 and rows are 16-byte aligned (ldz % 4 == 0)
    int jv = j0;
    if (sim_vec4(zt, zs, ldz, labels)) {
        const int j4 = j0 + ((j1 - j0) & ~3);
        // software pipeline: the next trip's 64 bytes are requested before this trip's columns are processed (the shared
        // atomics keep the compiler from hoisting the loads itself; ncu had half of the stall samples on the first use
        // of loaded data at 39 % of the HBM rate)
        int j = j0 + 4 * threadIdx.x;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        longlong2 l0 = make_longlong2(0, 0), l1 = l0;
        if (j < j4) {
            a = *reinterpret_cast<const float4*>(rt + j);
            b = *reinterpret_cast<const float4*>(rs + j);
            l0 = __ldg(reinterpret_cast<const longlong2*>(labels + j));
            l1 = __ldg(reinterpret_cast<const longlong2*>(labels + j + 2));
        }
        for (; j < j4; j += 4 * kSimBlock) {
            const int jn = j + 4 * kSimBlock;
            float4 an = a, bn = b;
            longlong2 l0n = l0, l1n = l1;
            if (jn < j4) {
                an = *reinterpret_cast<const float4*>(rt + jn);
                bn = *reinterpret_cast<const float4*>(rs + jn);
                l0n = __ldg(reinterpret_cast<const longlong2*>(labels + jn));
                l1n = __ldg(reinterpret_cast<const longlong2*>(labels + jn + 2));
            }
            column(a.x, b.x, (int)l0.x);
            column(a.y, b.y, (int)l0.y);
            column(a.z, b.z, (int)l1.x);
            column(a.w, b.w, (int)l1.y);
            a = an; b = bn; l0 = l0n; l1 = l1n;
        }
        jv = j4;
    }
    for (int j = jv + threadIdx.x; j < j1; j += kSimBlock) column(rt[j], rs[j], (int)labels[j]);
    num *= inv_st;
    st_ = block_reduce(st_, red, false);
    ss_ = block_reduce(ss_, red, false);
    num = block_reduce(num, red, false);
    float* out = stats + (long long)row * (3 + C);
    if (threadIdx.x == 0) { out[0] = st_; out[1] = ss_; out[2] = num; }
    for (int c = threadIdx.x; c < C; c += kSimBlock)
        out[3 + c] = __ull2float_rn(((unsigned long long)sHi[c] << kSimFxLoBits) + sLo[c]) * (1.f / kSimFxScale);
}

// stats[i] = sum over chunks, in chunk order (deterministic)
__global__ void __launch_bounds__(256) simmatch_shard_reduce_kernel(const float* __restrict__ parts, long long chunk_stride, int nchunk,
                                                                    long long n, float* __restrict__ stats) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int c = 0; c < nchunk; ++c) s += parts[(long long)c * chunk_stride + i];
    stats[i] = s;
}

// totals (summed over the shards) -> prob_ku (:280), loss_in (:286) and the two normalisers the gradient needs; warp per row
__global__ void __launch_bounds__(kRowBlock) simmatch_shard_finish_kernel(const float* __restrict__ stats, const float* __restrict__ p_all,
                                                                          int rows, int C, float inv_st, float c_smooth,
                                                                          float* __restrict__ p_out, float* __restrict__ loss_in,
                                                                          float* __restrict__ norms) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* st = stats + (long long)row * (3 + C);
    const float sum_t = st[0], sum_s = st[1], num = st[2];
    const float inv_sum_t = 1.f / sum_t;
    float den = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float p = p_all[(long long)row * C + c], a = st[3 + c];
        den += p * a;
        if (p_out) p_out[(long long)row * C + c] = c_smooth < 1.f ? p * c_smooth + (a * inv_sum_t) * (1.f - c_smooth) : p;
    }
    den = warp_sum(den);
    if (lane == 0) {
        // log S_ij = z_ij/st - (1/st + log sum_s);  loss_in = lse_s - sum_j T'_ij z_ij/st
        if (loss_in) loss_in[row] = inv_st + logf(sum_s) - num / den;
        norms[2 * row] = 1.f / sum_s;
        norms[2 * row + 1] = 1.f / den;
    }
}

// G_ij = (S_ij - T'_ij)/st for the columns of this shard, as a bf16 hi/lo operand of the dX GEMM
__global__ void __launch_bounds__(kSimBlock) simmatch_shard_grad_kernel(const float* __restrict__ zt, const float* __restrict__ zs,
                                                                        long long ldz, const long long* __restrict__ labels,
                                                                        int k_shard, const float* __restrict__ p_all, int C,
                                                                        float inv_tt, float inv_st, const float* __restrict__ norms,
                                                                        __nv_bfloat16* gop, long long ld_g, int g_nseg,
                                                                        float4* __restrict__ zero_out, long long zero_n4) {
    extern __shared__ float sm[];   // p[C]
    // the dX product that follows ADDS its partial tiles into its output: clear it from here (one memset node and one
    // launch gap less per sweep)
    for (long long i = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * kSimBlock + threadIdx.x; i < zero_n4;
         i += (long long)gridDim.x * gridDim.y * kSimBlock)
        zero_out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int row = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += kSimBlock) sm[c] = p_all[(long long)row * C + c];
    __syncthreads();
    const float* rt = zt + (long long)row * ldz;
    const float* rs = zs + (long long)row * ldz;
    const float inv_sum_s = norms[2 * row], inv_den = norms[2 * row + 1];
    __nv_bfloat16* gh = gop + (long long)row * g_nseg * ld_g;
    const int cw = ((k_shard + gridDim.y - 1) / gridDim.y + 3) & ~3;
    const int j0 = min(k_shard, (int)blockIdx.y * cw), j1 = min(k_shard, j0 + cw);
    const float ct = inv_tt * 1.4426950408889634f, cs = inv_st * 1.4426950408889634f;
    const float ws = inv_sum_s * inv_st, wt = inv_den * inv_st;
    auto column = [&](float z_t, float z_s, int y) { return shifted_exp(z_s, cs) * ws - shifted_exp(z_t, ct) * sm[y] * wt; };
    int jv = j0;
    if (sim_vec4(zt, zs, ldz, labels) && (ld_g & 3) == 0 && (reinterpret_cast<uintptr_t>(gop) & 7) == 0) {
        const int j4 = j0 + ((j1 - j0) & ~3);
#pragma unroll 2
        for (int j = j0 + 4 * threadIdx.x; j < j4; j += 4 * kSimBlock) {
            const float4 a = *reinterpret_cast<const float4*>(rt + j), b = *reinterpret_cast<const float4*>(rs + j);
            const longlong2 l0 = __ldg(reinterpret_cast<const longlong2*>(labels + j));
            const longlong2 l1 = __ldg(reinterpret_cast<const longlong2*>(labels + j + 2));
            const float g[4] = {column(a.x, b.x, (int)l0.x), column(a.y, b.y, (int)l0.y), column(a.z, b.z, (int)l1.x),
                                column(a.w, b.w, (int)l1.y)};
            __align__(8) __nv_bfloat16 h[4];
            __align__(8) __nv_bfloat16 l[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                h[u] = __float2bfloat16_rn(g[u]);
                l[u] = __float2bfloat16_rn(g[u] - __bfloat162float(h[u]));
            }
            *reinterpret_cast<uint2*>(gh + j) = *reinterpret_cast<const uint2*>(h);
            if (g_nseg > 1) *reinterpret_cast<uint2*>(gh + ld_g + j) = *reinterpret_cast<const uint2*>(l);
        }
        jv = j4;
    }
    for (int j = jv + threadIdx.x; j < j1; j += kSimBlock) {
        const float g = column(rt[j], rs[j], (int)labels[j]);
        const __nv_bfloat16 h = __float2bfloat16_rn(g);
        gh[j] = h;
        if (g_nseg > 1) gh[ld_g + j] = __float2bfloat16_rn(g - __bfloat162float(h));
    }
}

__global__ void zero_u32_kernel(unsigned int* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
}

template <int LPR, int NV>
int launch_cgpl_t(const CgplArgs& A, cudaStream_t stream) {
    const int threads = LPR > 32 ? LPR : row_block_threads(ceil_div(A.rows, LPR > 32 ? 1 : 32 / LPR));
    const int rows_per_block = LPR > 32 ? 1 : (threads / 32) * (32 / (LPR > 32 ? 32 : LPR));
    const int blocks = (int)ceil_div(A.rows, rows_per_block);
    static const bool once = (prefer_max_shared(cgpl_pgls_kernel<LPR, NV, false>), prefer_max_shared(cgpl_pgls_kernel<LPR, NV, true>), true);
    (void)once;
    // fp32 rows, even k, every row 8-byte aligned: the one-predicate float2 path
    const bool fast = A.logit_dtype == STIL_F32 && (A.k & 1) == 0 && A.vec_y && A.vec_t && A.vec_pl && (!A.prediction || A.vec_pred) &&
                      (!A.pred_in || A.vec_pin);
    if (fast) STIL_CUDA(launch_pdl(cgpl_pgls_kernel<LPR, NV, true>, dim3(blocks), dim3(threads), 0, stream, A));
    else STIL_CUDA(launch_pdl(cgpl_pgls_kernel<LPR, NV, false>, dim3(blocks), dim3(threads), 0, stream, A));
    return STIL_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------- launchers
int launch_labelled_cls(const int64_t* y_l, int64_t b_l, float th, int32_t* cls, uint8_t* conf, cudaStream_t stream);
void prep_add(PrepLaunch& L, const PrepJob& j) {
    PrepJob& d = L.job[L.njobs];
    d = j;
    d.block_begin = L.total_blocks;
    L.total_blocks += (int)ceil_div(j.rows, kPrepRows);
    L.njobs++;
}

int launch_prep(const PrepLaunch& L, cudaStream_t stream) {
    if (L.total_blocks == 0) {
        if (L.n_zero > 0) return launch_zero_u32(L.zero_words, L.n_zero, stream);
        return STIL_OK;
    }
    static const bool once = (prefer_max_shared(prep_kernel), true);
    (void)once;
    STIL_CUDA(launch_pdl(prep_kernel, dim3(L.total_blocks), dim3(kRowBlock), 0, stream, L));
    return STIL_OK;
}

int launch_zero_u32(unsigned int* p, int n, cudaStream_t stream) {
    zero_u32_kernel<<<(int)ceil_div(n, 128), 128, 0, stream>>>(p, n);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

// eight rows per block once there is a row per scheduler anyway: fewer blocks = a shorter ticket reduction
inline int finish_threads(int total_rows) { return total_rows >= 512 ? kRowBlock : 64; }
int64_t finish_blocks(int total_rows) { return ceil_div(total_rows, finish_threads(total_rows) / 32); }

int launch_finish(const FinishLaunch& L, cudaStream_t stream) {
    if (L.total_rows == 0) return STIL_OK;
    static const bool once = (prefer_max_shared(finish_kernel), true);
    (void)once;
    STIL_CUDA(launch_pdl(finish_kernel, dim3((unsigned)finish_blocks(L.total_rows)),
                         dim3((unsigned)finish_threads(L.total_rows)), 0, stream, L));
    return STIL_OK;
}

int launch_grad_finish(const GradFinishLaunch& L, cudaStream_t stream) {
    if (L.total_rows == 0) return STIL_OK;
    const int threads = row_block_threads(L.total_rows);
    STIL_CUDA(launch_pdl(grad_finish_kernel, dim3((unsigned)ceil_div(L.total_rows, threads / 32)), dim3((unsigned)threads), 0,
                         stream, L));
    return STIL_OK;
}

int launch_cgpl_pgls(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                     const float* teacher_logits, int64_t ld_t, int64_t rows, int64_t k, float temperature,
                     float rate_pseudo, float th1, int past_start_epoch, const float* prediction_in, int64_t ld_pin,
                     float* pseudo_label, int64_t ld_pl,
                     float* prediction, int64_t ld_pred, float* max_prob, int64_t* max_idx, uint8_t* mask1,
                     uint8_t* case1, uint8_t* case2_i, uint8_t* case2_t, uint8_t* case3, int64_t* top1,
                     int32_t* cls, uint8_t* conf, const int64_t* y_l, int64_t b_l, int32_t* cls_l, uint8_t* conf_l,
                     cudaStream_t stream) {
    STIL_REQUIRE(k >= 1 && k <= 1024, STIL_E_SHAPE, "cgpl_pgls supports 1 <= k <= 1024 classes (got %lld)", (long long)k);
    STIL_REQUIRE(rows >= 0 && rows < (1LL << 31), STIL_E_SHAPE, "rows out of range");
    if (rows == 0) return launch_labelled_cls(y_l, y_l ? b_l : 0, th1, cls_l, conf_l, stream);
    CgplArgs A;
    A.y_l = reinterpret_cast<const long long*>(y_l); A.b_l = y_l ? (int)b_l : 0; A.cls_l = cls_l; A.conf_l = conf_l;
    A.y_m = y_m; A.y_i = y_i; A.y_t = y_t;
    A.logit_dtype = logit_dtype; A.ld_y = ld_y;
    A.tl = teacher_logits; A.ld_t = ld_t;
    A.rows = (int)rows; A.k = (int)k;
    A.temperature = temperature;
    A.inv_temperature = 1.0f / temperature;
    A.third = 1.0f / 3.0f;
    A.rate_pseudo = rate_pseudo;
    // `1 - rate_pseudo` is formed in Python double and then rounded to fp32 (STiLModel.py:295)
    A.one_minus_rate = (float)(1.0 - (double)rate_pseudo);
    A.th1 = th1;
    A.past_start = past_start_epoch;
    A.pseudo_label = pseudo_label; A.ld_pl = ld_pl;
    A.prediction = prediction; A.ld_pred = ld_pred;
    A.max_prob = max_prob; A.max_idx = reinterpret_cast<long long*>(max_idx);
    A.mask1 = mask1; A.case1 = case1; A.case2_i = case2_i; A.case2_t = case2_t; A.case3 = case3;
    A.top1 = reinterpret_cast<long long*>(top1);
    A.cls = cls; A.conf = conf;
    const int esz = logit_dtype == STIL_BF16 ? 2 : 4;
    auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; };
    A.vec_y = (ld_y % 2 == 0) && al(y_m, 2 * esz) && al(y_i, 2 * esz) && al(y_t, 2 * esz);
    A.vec_t = (ld_t % 2 == 0) && al(teacher_logits, 8);
    A.vec_pl = (ld_pl % 2 == 0) && al(pseudo_label, 8);
    A.vec_pred = prediction ? ((ld_pred % 2 == 0) && al(prediction, 8)) : 0;
    A.pred_in = prediction_in; A.ld_pin = ld_pin;
    A.vec_pin = prediction_in ? ((ld_pin % 2 == 0) && al(prediction_in, 8)) : 0;
    // small batches are bound by each row's dependency chain (5 softmaxes): spread a row over 4 warps
    // (four warps per row measured best: eight warps per row was 2.4 us slower per C2 step)
    if (rows <= 2048 && k > 128 && k <= 512) return launch_cgpl_t<128, 2>(A, stream);
    if (rows <= 2048 && k > 512) return launch_cgpl_t<128, 4>(A, stream);
    if (k <= 2) return launch_cgpl_t<1, 1>(A, stream);
    if (k <= 16) return launch_cgpl_t<4, 2>(A, stream);
    if (k <= 64) return launch_cgpl_t<8, 4>(A, stream);
    if (k <= 128) return launch_cgpl_t<32, 2>(A, stream);
    if (k <= 320) return launch_cgpl_t<32, 5>(A, stream);
    if (k <= 512) return launch_cgpl_t<32, 8>(A, stream);
    return launch_cgpl_t<32, 16>(A, stream);
}

int launch_label_argmax(const float* label, int64_t ld, int64_t rows, int64_t k, float threshold, int32_t* cls,
                        uint8_t* conf, float* max_prob, cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    STIL_REQUIRE(k >= 1, STIL_E_SHAPE, "label_argmax needs k >= 1");
    label_argmax_kernel<<<(int)ceil_div(rows, kRowBlock / 32), kRowBlock, 0, stream>>>(label, ld, (int)rows, (int)k,
                                                                                      threshold, cls, conf, max_prob);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_labelled_cls(const int64_t* y_l, int64_t b_l, float th, int32_t* cls, uint8_t* conf, cudaStream_t stream) {
    if (b_l == 0) return STIL_OK;
    labelled_cls_kernel<<<(int)ceil_div(b_l, 128), 128, 0, stream>>>(reinterpret_cast<const long long*>(y_l), (int)b_l,
                                                                     th, cls, conf);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_proto_accumulate(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const int32_t* cls,
                            const uint8_t* conf, int64_t b_l, float repeat_ratio, int64_t k, float* class_sum,
                            float* class_count, float* psum, float* pcount, cudaStream_t stream) {
    if (k == 0) return STIL_OK;
    static const bool once = (prefer_max_shared(proto_accumulate_kernel), true);
    (void)once;
    // one block per class; few classes with many rows each (cardiac, K=2) get 8 warps, many small classes 2
    const int acc_threads = (rows / (k > 0 ? k : 1)) >= 16 ? kRowBlock : 64;
    proto_accumulate_kernel<<<(int)k, acc_threads, 0, stream>>>(
        feat, dtype, (int)rows, (int)dim, ld, cls, conf, (int)b_l, repeat_ratio, (int)k, class_sum, class_count, psum,
        pcount);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_proto_add(const float* class_sum, const float* class_count, int64_t k, int64_t dim, float* psum,
                     float* pcount, cudaStream_t stream) {
    if (k == 0) return STIL_OK;
    proto_add_kernel<<<(int)ceil_div(k * dim, 256), 256, 0, stream>>>(class_sum, class_count, (int)k, (int)dim, psum,
                                                                       pcount);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_proto_add_gathered(const float* parts, int64_t world, int64_t slot, int64_t k, int64_t dim, float* class_sum,
                              float* class_count, float* psum, float* pcount, cudaStream_t stream,
                              const unsigned long long* wait_flags, const unsigned long long* wait_target,
                              const unsigned long long* loss_ll, const unsigned long long* loss_tag, float* loss_out) {
    if (k == 0) return STIL_OK;
    proto_add_gathered_kernel<<<(int)ceil_div(k * dim + k, 256), 256, 0, stream>>>(parts, (int)world, slot, (int)k, (int)dim,
                                                                                class_sum, class_count, psum, pcount,
                                                                                wait_flags, wait_target, loss_ll, loss_tag,
                                                                                loss_out);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_proto_finalize(float* prototypes, float* psum, float* pcount, int64_t k, int64_t dim,
                          int32_t* empty_classes, cudaStream_t stream) {
    STIL_CUDA(cudaMemsetAsync(empty_classes, 0, sizeof(int32_t), stream));
    if (k == 0) return STIL_OK;
    proto_finalize_kernel<<<(int)k, 128, 0, stream>>>(prototypes, psum, pcount, (int)dim, empty_classes);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_col_sum(const float* x, int64_t ld, int64_t rows, int64_t k, float* out, cudaStream_t stream) {
    if (k == 0) return STIL_OK;
    col_sum_kernel<<<(unsigned)ceil_div(k, 32), 256, 0, stream>>>(x, (long long)ld, (int)rows, (int)k, out);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}
int launch_da_batch_mean(const float* probs, int64_t ld, int64_t rows, int64_t k, float* mean, cudaStream_t stream) {
    if (k == 0) return STIL_OK;
    da_batch_mean_kernel<<<(int)ceil_div(k, 32), 256, 0, stream>>>(probs, ld, (int)rows, (int)k, mean);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}
int launch_da_apply(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* batch_mean, float* da_queue,
                    int64_t da_len, int64_t* da_ptr, float* qmean, float* out, int64_t ld_out, cudaStream_t stream) {
    da_update_kernel<<<1, 256, 0, stream>>>(batch_mean, da_queue, (int)da_len, (int)k, reinterpret_cast<long long*>(da_ptr),
                                            qmean);
    STIL_LAUNCH_CHECK();
    return launch_da_rows(probs, ld, rows, k, qmean, out, ld_out, stream);
}
int launch_softmax_rows(const void* y, int dtype, int64_t ld, int64_t rows, int64_t k, float* out, int64_t ld_out,
                        cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    const int threads = row_block_threads(rows);
    STIL_CUDA(launch_pdl(softmax_rows_kernel, dim3((unsigned)ceil_div(rows, threads / 32)), dim3((unsigned)threads), 0, stream,
                         y, dtype, (long long)ld, (int)rows, (int)k, out, (long long)ld_out));
    return STIL_OK;
}
int launch_da_rows(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* qmean, float* out, int64_t ld_out,
                   cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    const int threads = row_block_threads(rows);
    da_apply_kernel<<<(int)ceil_div(rows, threads / 32), threads, 0, stream>>>(probs, ld, (int)rows, (int)k, qmean, out,
                                                                               ld_out);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_simmatch_rows(const float* zt, const float* zs, long long ldz, const long long* labels, int rows, int k_bank,
                         const float* p_orig, int num_classes, float tt, float st, float c_smooth, float* p_out,
                         float* loss_in, __nv_bfloat16* gop, long long ld_g, int g_nseg, cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    const size_t smem = (3 * (size_t)num_classes + 8) * sizeof(float);
    STIL_REQUIRE(smem <= 48 * 1024, STIL_E_SHAPE, "simmatch: too many classes (%d)", num_classes);
    simmatch_rows_kernel<<<rows, kSimBlock, smem, stream>>>(zt, zs, ldz, labels, k_bank, p_orig, num_classes, 1.0f / tt,
                                                            1.0f / st, c_smooth, p_out, loss_in, gop, ld_g, g_nseg);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int64_t masked_softce_blocks(int64_t rows, int64_t) {
    return std::min<int64_t>(ceil_div(rows, row_block_threads(rows) / 32), 148 * 16);
}

int simmatch_shard_chunks(int64_t rows, int64_t k_shard) {
    // ~2048 blocks in flight, at least 1024 columns per block
    const int64_t fill = std::min<int64_t>(std::min<int64_t>(ceil_div(2048, std::max<int64_t>(rows, 1)), 32), k_shard / 1024);
    // the fixed-point class sums of simmatch_shard_stats_kernel hold at most kSimFxCols columns per block
    return (int)std::max<int64_t>(std::max<int64_t>(1, fill), ceil_div(k_shard, kSimFxCols));
}
int launch_simmatch_shard_stats(const float* zt, const float* zs, long long ldz, const long long* labels, int rows, int k_shard,
                                const float* p_all, int num_classes, float tt, float st, float* stats, float* chunk_scratch,
                                cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    const size_t smem = (3 * (size_t)num_classes + 8) * sizeof(float);
    STIL_REQUIRE(smem <= 48 * 1024, STIL_E_SHAPE, "simmatch: too many classes (%d)", num_classes);
    const int nchunk = simmatch_shard_chunks(rows, k_shard);
    const long long n = (long long)rows * (3 + num_classes);
    float* dst = nchunk > 1 ? chunk_scratch : stats;
    simmatch_shard_stats_kernel<<<dim3(rows, nchunk), kSimBlock, smem, stream>>>(zt, zs, ldz, labels, k_shard, p_all, num_classes,
                                                                                 1.f / tt, 1.f / st, dst, n);
    STIL_LAUNCH_CHECK();
    if (nchunk > 1) {
        simmatch_shard_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(chunk_scratch, n, nchunk, n, stats);
        STIL_LAUNCH_CHECK();
    }
    return STIL_OK;
}
int launch_simmatch_shard_finish(const float* stats, const float* p_all, int rows, int num_classes, float st, float c_smooth,
                                 float* p_out, float* loss_in, float* norms, cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    simmatch_shard_finish_kernel<<<(unsigned)ceil_div(rows, kRowBlock / 32), kRowBlock, 0, stream>>>(stats, p_all, rows, num_classes,
                                                                                                    1.f / st, c_smooth, p_out,
                                                                                                    loss_in, norms);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}
int launch_simmatch_shard_grad(const float* zt, const float* zs, long long ldz, const long long* labels, int rows, int k_shard,
                               const float* p_all, int num_classes, float tt, float st, const float* norms, __nv_bfloat16* gop,
                               long long ld_g, int g_nseg, float* zero_out, long long zero_n, cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    const size_t smem = (size_t)num_classes * sizeof(float);
    simmatch_shard_grad_kernel<<<dim3(rows, simmatch_shard_chunks(rows, k_shard)), kSimBlock, smem, stream>>>(
        zt, zs, ldz, labels, k_shard, p_all, num_classes, 1.f / tt, 1.f / st, norms, gop, ld_g, g_nseg,
        reinterpret_cast<float4*>(zero_out), zero_out ? zero_n / 4 : 0);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_masked_softce(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                         const float* pseudo_label, int64_t ld_pl, const uint8_t* mask1, const uint8_t* case1,
                         const uint8_t* case2_i, const uint8_t* case2_t, const uint8_t* case3,
                         const uint8_t* mask_random, int64_t rows, int64_t k, float* losses, float* d_y_m,
                         float* d_y_i, float* d_y_t, int64_t ld_g, float grad_scale, float* block_partials,
                         unsigned int* ticket, cudaStream_t stream) {
    if (rows == 0) {
        STIL_CUDA(cudaMemsetAsync(losses, 0, 3 * sizeof(float), stream));
        return STIL_OK;
    }
    SoftCeArgs A;
    A.y[0] = y_m; A.y[1] = y_i; A.y[2] = y_t;
    A.logit_dtype = logit_dtype; A.ld_y = ld_y;
    A.pl = pseudo_label; A.ld_pl = ld_pl;
    A.mask1 = mask1; A.case1 = case1; A.case2_i = case2_i; A.case2_t = case2_t; A.case3 = case3;
    A.mask_random = mask_random;
    A.rows = (int)rows; A.k = (int)k;
    A.dy[0] = d_y_m; A.dy[1] = d_y_i; A.dy[2] = d_y_t;
    A.ld_g = ld_g; A.grad_scale = grad_scale;
    A.block_partials = block_partials; A.ticket = ticket; A.losses = losses;
    STIL_REQUIRE(k >= 1 && k <= 1024, STIL_E_SHAPE, "masked_softce supports 1 <= k <= 1024 classes (got %lld)", (long long)k);
    const int esz = logit_dtype == STIL_BF16 ? 2 : 4;
    auto al = [](const void* p, int a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
    A.vec_y = (ld_y % 2 == 0) && al(y_m, 2 * esz) && al(y_i, 2 * esz) && al(y_t, 2 * esz);
    A.vec_pl = (ld_pl % 2 == 0) && al(pseudo_label, 8);
    A.vec_g = (ld_g % 2 == 0) && al(d_y_m, 8) && al(d_y_i, 8) && al(d_y_t, 8);
    const int threads = row_block_threads(rows);
    // fp32 rows, even k, 8-byte aligned rows, all three gradients wanted: the one-predicate float2 path
    const bool fast = logit_dtype == STIL_F32 && (k & 1) == 0 && A.vec_y && A.vec_pl && A.vec_g;
    auto go = [&](auto kernel_fast, auto kernel_gen, int lpr) {
        const int64_t rows_per_block = (threads / 32) * (32 / lpr);
        // never more blocks than masked_softce_blocks() sized the partials for
        const int blocks = (int)std::min<int64_t>(ceil_div(rows, rows_per_block), masked_softce_blocks(rows, k));
        if (fast) kernel_fast<<<blocks, threads, 0, stream>>>(A);
        else kernel_gen<<<blocks, threads, 0, stream>>>(A);
    };
#define STIL_SOFTCE(LPR, NV) go(masked_softce_kernel<LPR, NV, true>, masked_softce_kernel<LPR, NV, false>, LPR)
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool k2 = k == 2 && logit_dtype == STIL_F32 && ld_y == 2 && ld_pl == 2 && ld_g == 2 && al16(y_m) && al16(y_i) && al16(y_t) &&
                    al16(pseudo_label) && al16(d_y_m) && al16(d_y_i) && al16(d_y_t) &&
                    ((reinterpret_cast<uintptr_t>(mask1) | reinterpret_cast<uintptr_t>(case1) | reinterpret_cast<uintptr_t>(case2_i) |
                      reinterpret_cast<uintptr_t>(case2_t) | reinterpret_cast<uintptr_t>(case3) | reinterpret_cast<uintptr_t>(mask_random)) & 1) == 0;
    if (k2) {
        const int blocks = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(rows / 2, threads), 1), masked_softce_blocks(rows, k));
        masked_softce_k2_kernel<<<blocks, threads, 0, stream>>>(A);
    } else if (k <= 2) STIL_SOFTCE(1, 1);
    else if (k <= 16) STIL_SOFTCE(4, 2);
    else if (k <= 64) STIL_SOFTCE(8, 4);
    else if (k <= 128) STIL_SOFTCE(32, 2);
    else if (k <= 320) STIL_SOFTCE(32, 5);
    else if (k <= 512) STIL_SOFTCE(32, 8);
    else STIL_SOFTCE(32, 16);
#undef STIL_SOFTCE
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

}  // namespace stil
