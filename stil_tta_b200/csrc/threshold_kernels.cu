// Pseudo-label thresholds of the remaining baselines (VERDICT "missing 6"):
//   * FreeMatch self-adaptive threshold — models/MatchModel/FreeMatchFolder/freematch_model.py:128-165 (`update`, `masking`)
//   * FreeMatch fairness ("entropy") loss — FreeMatchFolder/freematch_utils.py:17-45, forward and gradient
//   * CoTraining cross pseudo labels — models/SemiMultimodal/CoTraining.py:141-146 (softmax, row max, threshold)
// All of it is [rows, C] row work plus C-sized state; bit-exact integer parts (argmax, histograms), fp32 elsewhere.
#include <cstring>

#include "common.cuh"
#include "internal.h"

namespace stil {
namespace {

constexpr int kThrBlock = 256;

// row max / first arg max of a probability row (torch.max(dim=-1)), optional histogram of the arg max (torch.bincount)
// and optional per-row selection mask: unselected rows are skipped (their probabilities are ZEROED in place when
// `zero_unselected`, so a plain column sum afterwards is the sum over the selected rows)
__global__ void __launch_bounds__(kThrBlock) row_max_kernel(float* __restrict__ probs, long long ld, int rows, int C,
                                                            const float* __restrict__ select, int zero_unselected,
                                                            float* __restrict__ max_p, long long* __restrict__ max_i,
                                                            int* __restrict__ hist) {
    const int row = blockIdx.x * (kThrBlock >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* p = probs + (long long)row * ld;
    if (select && !(select[row] != 0.f)) {
        if (zero_unselected)
            for (int c = lane; c < C; c += 32) p[c] = 0.f;
        if (lane == 0) {
            if (max_p) max_p[row] = 0.f;
            if (max_i) max_i[row] = 0;
        }
        return;
    }
    float bv = -INFINITY;
    int bi = 0;
    for (int c = lane; c < C; c += 32) {
        const float v = p[c];
        if (v > bv || (v != v && !(bv != bv))) { bv = v; bi = c; }      // first maximum; a NaN wins like in torch
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool o_nan = ov != ov, b_nan = bv != bv;
        // the other lane wins with a NaN against a number, a larger value, or the same value (or NaN) at a smaller index
        const bool take = o_nan ? (!b_nan || oi < bi) : (!b_nan && (ov > bv || (ov == bv && oi < bi)));
        if (take) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
        if (max_p) max_p[row] = bv;
        if (max_i) max_i[row] = bi;
        if (hist) atomicAdd(&hist[bi], 1);
    }
}

// stats = [ colsum(probs) (C) | hist (C, as float) | sum(max_probs) | rows ]: additive over ranks (the reference gathers the
// probabilities of all ranks first, freematch_model.py:129-130; summing these is the same thing)
__global__ void __launch_bounds__(kThrBlock) freematch_pack_stats_kernel(const float* __restrict__ colsum, const int* __restrict__ hist,
                                                                         const float* __restrict__ summax, int rows, int C,
                                                                         float* __restrict__ stats) {
    for (int c = threadIdx.x; c < C; c += kThrBlock) {
        stats[c] = colsum[c];
        stats[C + c] = (float)hist[c];
    }
    if (threadIdx.x == 0) {
        stats[2 * C] = summax[0];
        stats[2 * C + 1] = (float)rows;
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < kThrBlock / 32; ++i) r += red[i];
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = -INFINITY;
    for (int i = 0; i < kThrBlock / 32; ++i) r = fmaxf(r, red[i]);
    return r;
}

// `update` (:133-143) on the device-resident state, then the per-class thresholds time_p * p_model / max(p_model) (:162-163)
__global__ void __launch_bounds__(kThrBlock) freematch_state_kernel(const float* __restrict__ stats, int C, float m, float clip_thresh,
                                                                    float* time_p, float* p_model, float* label_hist,
                                                                    float* __restrict__ thr) {
    __shared__ float red[kThrBlock / 32];
    const float n = stats[2 * C + 1];
    const float om = (float)(1.0 - (double)m);                       // python: (1 - self.m), then cast with the tensor
    float tp = __fadd_rn(__fmul_rn(*time_p, m), __fmul_rn(om, __fdiv_rn(stats[2 * C], n)));                 // :137
    if (clip_thresh != 0.f) tp = fminf(fmaxf(tp, 0.f), 0.95f);                                              // :139-140
    float pmax = -INFINITY;
    for (int c = threadIdx.x; c < C; c += kThrBlock) {
        const float pm = __fadd_rn(__fmul_rn(p_model[c], m), __fmul_rn(om, __fdiv_rn(stats[c], n)));        // :142
        p_model[c] = pm;
        label_hist[c] = __fadd_rn(__fmul_rn(label_hist[c], m), __fmul_rn(om, __fdiv_rn(stats[C + c], n)));  // :143-144 (hist.sum() = n)
        pmax = fmaxf(pmax, pm);
    }
    pmax = block_max(pmax, red);
    for (int c = threadIdx.x; c < C; c += kThrBlock) thr[c] = __fmul_rn(tp, __fdiv_rn(p_model[c], pmax));     // :162-163
    if (threadIdx.x == 0) *time_p = tp;
}

__global__ void __launch_bounds__(kThrBlock) threshold_mask_kernel(const float* __restrict__ max_p, const long long* __restrict__ max_i,
                                                                   const float* __restrict__ thr, float thr_const, int rows,
                                                                   float* __restrict__ mask) {
    const int i = blockIdx.x * kThrBlock + threadIdx.x;
    if (i >= rows) return;
    const float t = thr ? thr[max_i[i]] : thr_const;
    mask[i] = max_p[i] >= t ? 1.f : 0.f;                            // max_probs.ge(...).to(dtype)
}

// fairness loss, scalar part (freematch_utils.py:26-45): from the selected rows' histogram and column sums
//   saved[0..C) = h_k with d loss / d prob_s[i, k] = h_k for every selected row i;  saved[C] = number of selected rows
__global__ void __launch_bounds__(kThrBlock) freematch_entropy_finish_kernel(const float* __restrict__ colsum, const int* __restrict__ hist,
                                                                             const float* __restrict__ p_model,
                                                                             const float* __restrict__ label_hist, int C,
                                                                             float* __restrict__ loss, float* __restrict__ hist_mean,
                                                                             float* __restrict__ saved) {
    __shared__ float red[kThrBlock / 32];
    float nf = 0.f;
    for (int c = threadIdx.x; c < C; c += kThrBlock) nf += (float)hist[c];
    const float n = block_sum(nf, red);
    // modulated model distribution (:31-35) and modulated mean prediction (:38-41)
    float s_pm = 0.f, s_u = 0.f;
    for (int c = threadIdx.x; c < C; c += kThrBlock) {
        float sc = __fdiv_rn(1.f, label_hist[c]);
        if (sc == INFINITY) sc = 0.f;                               // replace_inf_to_zero
        s_pm += p_model[c] * sc;
        const float hs = __fdiv_rn((float)hist[c], n);
        float ss = __fdiv_rn(1.f, hs);
        if (ss == INFINITY) ss = 0.f;
        s_u += __fdiv_rn(colsum[c], n) * ss;
    }
    s_pm = block_sum(s_pm, red);
    s_u = block_sum(s_u, red);
    float l = 0.f, gw = 0.f;
    for (int c = threadIdx.x; c < C; c += kThrBlock) {
        float sc = __fdiv_rn(1.f, label_hist[c]);
        if (sc == INFINITY) sc = 0.f;
        const float q = __fdiv_rn(p_model[c] * sc, s_pm);
        const float hs = __fdiv_rn((float)hist[c], n);
        float ss = __fdiv_rn(1.f, hs);
        if (ss == INFINITY) ss = 0.f;
        const float w = __fdiv_rn(__fdiv_rn(colsum[c], n) * ss, s_u);
        l += q * logf(w + 1e-12f);                                   // :43
        gw += __fdiv_rn(q, w + 1e-12f) * w;
    }
    l = block_sum(l, red);
    gw = block_sum(gw, red);
    for (int c = threadIdx.x; c < C; c += kThrBlock) {
        float sc = __fdiv_rn(1.f, label_hist[c]);
        if (sc == INFINITY) sc = 0.f;
        const float q = __fdiv_rn(p_model[c] * sc, s_pm);
        const float hs = __fdiv_rn((float)hist[c], n);
        float ss = __fdiv_rn(1.f, hs);
        if (ss == INFINITY) ss = 0.f;
        const float w = __fdiv_rn(__fdiv_rn(colsum[c], n) * ss, s_u);
        saved[c] = ss * (__fdiv_rn(q, w + 1e-12f) - gw) / (s_u * n);
    }
    if (threadIdx.x == 0) {
        *loss = l;                                                   // .sum(dim=1).mean() over one row
        *hist_mean = 1.f / (float)C;                                 // hist_s.mean(): a histogram that sums to one
        saved[C] = n;
    }
}

// d logits_s[i, j] = up * p_ij (h_j - sum_k h_k p_ik) for selected rows (probabilities of unselected rows are zero)
__global__ void __launch_bounds__(kThrBlock) freematch_entropy_grad_kernel(const float* __restrict__ probs, long long ld, int rows, int C,
                                                                           const float* __restrict__ h, const float* __restrict__ up,
                                                                           float* __restrict__ d_logits, long long ld_g) {
    const int row = blockIdx.x * (kThrBlock >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = probs + (long long)row * ld;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot += h[c] * p[c];
    dot = warp_sum(dot);
    const float u = up ? *up : 1.f;
    for (int c = lane; c < C; c += 32) d_logits[(long long)row * ld_g + c] = u * p[c] * (h[c] - dot);
}

}  // namespace

int64_t threshold_workspace_bytes(int64_t rows, int64_t C) {
    Workspace W(nullptr, 0);
    W.take<float>(rows * C);      // probabilities
    W.take<float>(rows);          // max_p (when the caller does not want it)
    W.take<long long>(rows);      // max_i
    W.take<int>(C);               // histogram
    W.take<float>(C + 8);         // column sums (+ sum of the row maxima)
    W.take<float>(C + 8);         // thresholds / saved gradient vector
    return W.off;
}

// softmax (optional) -> probabilities, row max / arg max, histogram, column sums -> stats [2C + 2]
int launch_freematch_stats(const float* x, int64_t ld, int64_t rows, int64_t C, int is_logits, float* stats, float* max_p,
                           int64_t* max_i, float* probs_out, int64_t ld_probs, void* workspace, int64_t workspace_bytes,
                           cudaStream_t stream) {
    STIL_REQUIRE(workspace && workspace_bytes >= threshold_workspace_bytes(rows, C), STIL_E_WORKSPACE, "freematch workspace too small");
    Workspace W(workspace, workspace_bytes);
    float* probs_ws = W.take<float>(rows * C);
    float* maxp_ws = W.take<float>(rows);
    long long* maxi_ws = W.take<long long>(rows);
    int* hist = W.take<int>(C);
    float* colsum = W.take<float>(C + 8);
    float* probs = probs_out ? probs_out : probs_ws;
    const int64_t ldp = probs_out ? ld_probs : C;
    float* mp = max_p ? max_p : maxp_ws;
    long long* mi = max_i ? reinterpret_cast<long long*>(max_i) : maxi_ws;
    int rc;
    if (is_logits) {
        if ((rc = launch_softmax_rows(x, STIL_F32, ld, rows, C, probs, ldp, stream))) return rc;
    } else {
        STIL_CUDA(cudaMemcpy2DAsync(probs, ldp * sizeof(float), x, ld * sizeof(float), C * sizeof(float), rows, cudaMemcpyDeviceToDevice,
                                    stream));
    }
    STIL_CUDA(cudaMemsetAsync(hist, 0, C * sizeof(int), stream));
    row_max_kernel<<<(unsigned)ceil_div(rows, kThrBlock / 32), kThrBlock, 0, stream>>>(probs, ldp, (int)rows, (int)C, nullptr, 0, mp, mi, hist);
    STIL_LAUNCH_CHECK();
    if ((rc = launch_col_sum(probs, ldp, rows, C, colsum, stream))) return rc;
    if ((rc = launch_col_sum(mp, 1, rows, 1, colsum + C, stream))) return rc;
    freematch_pack_stats_kernel<<<1, kThrBlock, 0, stream>>>(colsum, hist, colsum + C, (int)rows, (int)C, stats);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_freematch_update_mask(const float* stats_total, int64_t rows, int64_t C, float momentum, float clip_thresh, float* time_p,
                                 float* p_model, float* label_hist, const float* max_p, const int64_t* max_i, float* mask,
                                 void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    STIL_REQUIRE(workspace && workspace_bytes >= threshold_workspace_bytes(rows, C), STIL_E_WORKSPACE, "freematch workspace too small");
    Workspace W(workspace, workspace_bytes);
    W.take<float>(rows * C);
    float* maxp_ws = W.take<float>(rows);
    long long* maxi_ws = W.take<long long>(rows);
    W.take<int>(C);
    W.take<float>(C + 8);
    float* thr = W.take<float>(C + 8);
    freematch_state_kernel<<<1, kThrBlock, 0, stream>>>(stats_total, (int)C, momentum, clip_thresh, time_p, p_model, label_hist, thr);
    STIL_LAUNCH_CHECK();
    if (rows == 0) return STIL_OK;
    threshold_mask_kernel<<<(unsigned)ceil_div(rows, kThrBlock), kThrBlock, 0, stream>>>(
        max_p ? max_p : maxp_ws, max_i ? reinterpret_cast<const long long*>(max_i) : maxi_ws, thr, 0.f, (int)rows, mask);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

// CoTraining.py:141-146: probabilities, row maxima and the fixed-threshold mask of one classifier's teacher logits
int launch_threshold_rows(const float* logits, int64_t ld, int64_t rows, int64_t C, float threshold, float* probs, int64_t ld_probs,
                          float* max_p, int64_t* max_i, float* mask, cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    int rc;
    if ((rc = launch_softmax_rows(logits, STIL_F32, ld, rows, C, probs, ld_probs, stream))) return rc;
    row_max_kernel<<<(unsigned)ceil_div(rows, kThrBlock / 32), kThrBlock, 0, stream>>>(probs, ld_probs, (int)rows, (int)C, nullptr, 0, max_p,
                                                                                        reinterpret_cast<long long*>(max_i), nullptr);
    STIL_LAUNCH_CHECK();
    threshold_mask_kernel<<<(unsigned)ceil_div(rows, kThrBlock), kThrBlock, 0, stream>>>(max_p, nullptr, nullptr, threshold, (int)rows, mask);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_freematch_entropy_fwd(const float* mask, const float* logits_s, int64_t ld, int64_t rows, int64_t C, const float* p_model,
                                 const float* label_hist, float* loss, float* hist_mean, void* workspace, int64_t workspace_bytes,
                                 cudaStream_t stream) {
    STIL_REQUIRE(workspace && workspace_bytes >= threshold_workspace_bytes(rows, C), STIL_E_WORKSPACE, "freematch workspace too small");
    Workspace W(workspace, workspace_bytes);
    float* probs = W.take<float>(rows * C);
    W.take<float>(rows);
    W.take<long long>(rows);
    int* hist = W.take<int>(C);
    float* colsum = W.take<float>(C + 8);
    float* saved = W.take<float>(C + 8);
    int rc;
    if ((rc = launch_softmax_rows(logits_s, STIL_F32, ld, rows, C, probs, C, stream))) return rc;
    STIL_CUDA(cudaMemsetAsync(hist, 0, C * sizeof(int), stream));
    if (rows > 0) {
        row_max_kernel<<<(unsigned)ceil_div(rows, kThrBlock / 32), kThrBlock, 0, stream>>>(probs, C, (int)rows, (int)C, mask, 1, nullptr, nullptr,
                                                                                            hist);
        STIL_LAUNCH_CHECK();
    }
    if ((rc = launch_col_sum(probs, C, rows, C, colsum, stream))) return rc;
    freematch_entropy_finish_kernel<<<1, kThrBlock, 0, stream>>>(colsum, hist, p_model, label_hist, (int)C, loss, hist_mean, saved);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_freematch_entropy_bwd(int64_t rows, int64_t C, const float* grad_loss, float* d_logits, int64_t ld_g, void* workspace,
                                 int64_t workspace_bytes, cudaStream_t stream) {
    STIL_REQUIRE(workspace && workspace_bytes >= threshold_workspace_bytes(rows, C), STIL_E_WORKSPACE, "freematch workspace too small");
    if (rows == 0) return STIL_OK;
    Workspace W(workspace, workspace_bytes);
    const float* probs = W.take<float>(rows * C);
    W.take<float>(rows);
    W.take<long long>(rows);
    W.take<int>(C);
    W.take<float>(C + 8);
    const float* saved = W.take<float>(C + 8);
    freematch_entropy_grad_kernel<<<(unsigned)ceil_div(rows, kThrBlock / 32), kThrBlock, 0, stream>>>(probs, C, (int)rows, (int)C, saved, grad_loss,
                                                                                                    d_logits, ld_g);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

}  // namespace stil
