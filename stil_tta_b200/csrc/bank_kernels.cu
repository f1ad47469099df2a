// Row kernels of the memory-bank baselines' head blocks (SURVEY §8 rows a8, a9 and the queue maintenance next to
// a7-a9): bank softmax -> bf16 operand, smoothing mix + max/argmax/mask, CoMatch graph contrastive loss (forward +
// gradient), embedding-graph gradient operand, single-head masked soft/hard CE, FIFO queue writes.
// All HBM-bound; warp-shuffle / block reductions, coalesced row access.
#include <algorithm>

#include "internal.h"

namespace stil {

namespace {

constexpr int kBlk = 256;

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float blk_reduce(float v, float* red, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = is_max ? -INFINITY : 0.f;
#pragma unroll
    for (int i = 0; i < kBlk / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

__device__ __forceinline__ void store_split(__nv_bfloat16* hi, long long seg_stride, int nseg, long long j, float g) {
    const __nv_bfloat16 h = __float2bfloat16_rn(g);
    hi[j] = h;
    if (nseg > 1) {
        const float r = g - __bfloat162float(h);
        const __nv_bfloat16 l = __float2bfloat16_rn(r);
        hi[seg_stride + j] = l;
        if (nseg > 2) hi[2 * seg_stride + j] = __float2bfloat16_rn(r - __bfloat162float(l));
    }
}

// A = exp(z/T) / rowsum (comatch_model.py:291-292, MMatch.py:225-226) written as a bf16 hi/lo tensor-core operand
// [rows, nseg, ldg] (zero beyond k_q); one block per row
__global__ void __launch_bounds__(kBlk) bank_softmax_rows_kernel(const float* __restrict__ z, long long ldz, int k_q,
                                                                 float inv_t, __nv_bfloat16* gop, long long ldg, int nseg) {
    __shared__ float red[kBlk / 32];
    pdl_wait();                 // predecessor complete and visible ...
    pdl_launch_dependents();    // ... before the next kernel of the chain may start (see gemm_tc05.cu)
    const int row = blockIdx.x;
    const float* r = z + (long long)row * ldz;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < k_q; j += kBlk) m = fmaxf(m, r[j] * inv_t);
    m = blk_reduce(m, red, true);
    float s = 0.f;
    for (int j = threadIdx.x; j < k_q; j += kBlk) s += expf(r[j] * inv_t - m);
    s = blk_reduce(s, red, false);
    const float inv_s = 1.0f / s;
    __nv_bfloat16* o = gop + (long long)row * nseg * ldg;
    for (int j = threadIdx.x; j < ldg; j += kBlk)
        store_split(o, ldg, nseg, j, j < k_q ? expf(r[j] * inv_t - m) * inv_s : 0.f);
}

// out = c_keep*p + c_bank*s (two rounded products, then a rounded sum — the reference's eager order), then
// max / first-index argmax / threshold mask (MMatch.py:229-230, CoMatch.py:92-93); one warp per row
__global__ void __launch_bounds__(kBlk) smooth_mix_kernel(const float* __restrict__ p, long long ld_p,
                                                          const float* __restrict__ s, long long ld_s, int rows, int k,
                                                          float c_keep, float c_bank, float* out, long long ld_out,
                                                          float th, float* max_prob, long long* max_idx,
                                                          unsigned char* mask) {
    pdl_wait();                 // predecessor complete and visible ...
    pdl_launch_dependents();    // ... before the next kernel of the chain may start (see gemm_tc05.cu)
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int c = lane; c < k; c += 32) {
        float v = p[(long long)row * ld_p + c];
        if (s) v = __fadd_rn(__fmul_rn(c_keep, v), __fmul_rn(c_bank, s[(long long)row * ld_s + c]));
        if (out) out[(long long)row * ld_out + c] = v;
        if (v > best) { best = v; bidx = c; }      // strict: the first index of a lane's maxima
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) {
        if (max_prob) max_prob[row] = best;
        if (max_idx) max_idx[row] = bidx;
        if (mask) mask[row] = best >= th ? 1 : 0;
    }
}

// G = grad_sim * sim / T  (d/d logits of sim = exp(logits/T)) as bf16 operands: columns [0, n_self) -> gs, the rest -> gp
__global__ void __launch_bounds__(kBlk) sim_grad_kernel(const float* __restrict__ gsim, const float* __restrict__ sim,
                                                        long long ld, int n_self, int k_q, float inv_t,
                                                        __nv_bfloat16* gs, long long ldg_s, __nv_bfloat16* gp,
                                                        long long ldg_p, int nseg) {
    pdl_wait();                 // predecessor complete and visible ...
    pdl_launch_dependents();    // ... before the next kernel of the chain may start (see gemm_tc05.cu)
    const int row = blockIdx.y;
    const long long j = (long long)blockIdx.x * kBlk + threadIdx.x;
    const long long base = (long long)row * ld;
    if (j < ldg_s) {
        const float g = j < n_self ? gsim[base + j] * sim[base + j] * inv_t : 0.f;
        store_split(gs + (long long)row * nseg * ldg_s, ldg_s, nseg, j, g);
    }
    if (j < ldg_p) {
        const float g = j < k_q ? gsim[base + n_self + j] * sim[base + n_self + j] * inv_t : 0.f;
        store_split(gp + (long long)row * nseg * ldg_p, ldg_p, nseg, j, g);
    }
}

// CoMatch.py:100-110, one block per row:
//   pos = Q >= th; w = Q*pos / sum(Q*pos); p = sim*pos / sum(sim); loss_row = -sum_pos w * log(p + 1e-7)
//   d loss_row / d sim_k = ( sum_pos w_j p_j/(p_j+eps)  -  pos_k w_k/(p_k+eps) ) / sum(sim)
__global__ void __launch_bounds__(kBlk) graph_contrast_kernel(const float* __restrict__ Q, const float* __restrict__ sim,
                                                              long long ld, int cols, float th, float* d_sim,
                                                              float grad_scale, float* partials, unsigned int* ticket,
                                                              float* loss) {
    __shared__ float red[kBlk / 32];
    pdl_wait();                 // predecessor complete and visible ...
    pdl_launch_dependents();    // ... before the next kernel of the chain may start (see gemm_tc05.cu)
    const int row = blockIdx.x;
    const float* q = Q + (long long)row * ld;
    const float* s = sim + (long long)row * ld;
    float qs = 0.f, ss = 0.f;
    for (int j = threadIdx.x; j < cols; j += kBlk) {
        const float qv = q[j];
        qs += qv >= th ? qv : 0.f;
        ss += s[j];
    }
    qs = blk_reduce(qs, red, false);
    ss = blk_reduce(ss, red, false);
    const float inv_q = 1.0f / qs, inv_s = 1.0f / ss;
    float lr = 0.f, t2 = 0.f;
    for (int j = threadIdx.x; j < cols; j += kBlk) {
        const float qv = q[j];
        if (qv >= th) {
            const float w = qv * inv_q, p = s[j] * inv_s;
            lr -= w * logf(p + 1e-7f);
            t2 += w * p / (p + 1e-7f);
        }
    }
    lr = blk_reduce(lr, red, false);
    t2 = blk_reduce(t2, red, false);
    if (d_sim) {
        float* d = d_sim + (long long)row * ld;
        const float gsc = grad_scale / (float)gridDim.x;
        for (int j = threadIdx.x; j < cols; j += kBlk) {
            const float qv = q[j];
            float g = t2;
            if (qv >= th) g -= (qv * inv_q) / (s[j] * inv_s + 1e-7f);
            d[j] = g * inv_s * gsc;
        }
    }
    ticket_sum(lr, partials, ticket, loss, 1.0f / (float)gridDim.x);
}

// mean_i mask_i * CE(logits_i, target_i) with a probability target (SimMatch.py:91, CoMatch.py:96-97) or a class-index
// target (the dense one-hot of MMatch.py:231-234), forward + gradient; one warp per row
__global__ void __launch_bounds__(kBlk) weighted_softce_kernel(const void* __restrict__ y, int dtype, long long ld_y,
                                                               const float* __restrict__ tp, long long ld_t,
                                                               const long long* __restrict__ tidx,
                                                               const unsigned char* __restrict__ mask, int rows, int k,
                                                               float* d_y, long long ld_g, float grad_scale,
                                                               float* partials, unsigned int* ticket, float* loss) {
    pdl_wait();                 // predecessor complete and visible ...
    pdl_launch_dependents();    // ... before the next kernel of the chain may start (see gemm_tc05.cu)
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    float lrow = 0.f;
    if (row < rows) {
        const long long yo = (long long)row * ld_y;
        float m = -INFINITY;
        for (int c = lane; c < k; c += 32) m = fmaxf(m, ld_as_float(y, dtype, yo + c));
        m = warp_max(m);
        float se = 0.f, st = 0.f, sty = 0.f;
        const int ti = tidx ? (int)tidx[row] : -1;
        for (int c = lane; c < k; c += 32) {
            const float v = ld_as_float(y, dtype, yo + c);
            se += expf(v - m);
            const float t = tp ? tp[(long long)row * ld_t + c] : (c == ti ? 1.f : 0.f);
            st += t;
            sty += t * v;
        }
        se = warp_sum(se);
        st = warp_sum(st);
        sty = warp_sum(sty);
        const float lse = m + logf(se);
        const float w = mask ? (mask[row] ? 1.f : 0.f) : 1.f;
        lrow = w * (st * lse - sty);
        if (d_y) {
            const float gsc = w * grad_scale / (float)rows;
            for (int c = lane; c < k; c += 32) {
                const float v = ld_as_float(y, dtype, yo + c);
                const float t = tp ? tp[(long long)row * ld_t + c] : (c == ti ? 1.f : 0.f);
                d_y[(long long)row * ld_g + c] = (expf(v - lse) * st - t) * gsc;
            }
        }
    }
    const float b = block_sum(lane == 0 ? lrow : 0.f);
    ticket_sum(b, partials, ticket, loss, 1.0f / (float)rows);
}

// queue[:, ptr:ptr+n'] = z[:n'].T ; probs[:, ptr:ptr+n'] = t[:n'].T with n' = min(n, K - ptr)  (truncation at the wrap
// point, comatch_model.py:117-146 / MMatch.py:102-117).  The pointer is read on the device (no host sync).
__global__ void __launch_bounds__(kBlk) queue_enqueue_kernel(void* queue, int q_dtype, long long ld_q, float* qprobs,
                                                             long long ld_qp, int k_q, const long long* ptr,
                                                             const void* __restrict__ z, int z_dtype, long long ld_z,
                                                             int n, int dim, const float* __restrict__ t, long long ld_t,
                                                             int num_classes) {
    const int p = (int)(*ptr);
    const int nn = min(n, k_q - p);
    const int total_rows = dim + num_classes;       // rows of the two queues stacked
    // a 32x32 tile of (queue row r, sample i) per warp-iteration: threads walk samples fastest for coalesced writes
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)total_rows * nn;
         idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / nn), i = (int)(idx % nn);
        if (r < dim)
            st_from_float(queue, q_dtype, (long long)r * ld_q + p + i, ld_as_float(z, z_dtype, (long long)i * ld_z + r));
        else
            qprobs[(long long)(r - dim) * ld_qp + p + i] = t[(long long)i * ld_t + (r - dim)];
    }
}
__global__ void queue_advance_kernel(long long* ptr, int n, int k_q) {
    const int p = (int)(*ptr);
    *ptr = (p + min(n, k_q - p)) % k_q;
}

// bank[:, index] = k.T ; labels[index] = y   (simmatch_model.py:141-147)
__global__ void __launch_bounds__(kBlk) bank_update_kernel(void* bank, int b_dtype, long long ld_b, long long* labels,
                                                           const void* __restrict__ k, int k_dtype, long long ld_k,
                                                           const long long* __restrict__ y,
                                                           const long long* __restrict__ index, int n, int dim) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)dim * n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / n), i = (int)(idx % n);
        st_from_float(bank, b_dtype, (long long)r * ld_b + index[i], ld_as_float(k, k_dtype, (long long)i * ld_k + r));
        if (r == 0) labels[index[i]] = y[i];
    }
}

// CoMatch's distribution alignment keeps the last <= hist_len batch means (comatch_model.py:271-283):
// hist[count % hist_len] = batch_mean; count += 1; qmean = mean of the min(count, hist_len) valid rows
__global__ void __launch_bounds__(256) da_hist_update_kernel(const float* __restrict__ batch_mean, float* hist, int hist_len,
                                                             int k, long long* count, float* qmean) {
    const long long cnt = *count;
    const int slot = (int)(cnt % hist_len);
    for (int c = threadIdx.x; c < k; c += blockDim.x) hist[(long long)slot * k + c] = batch_mean[c];
    __syncthreads();
    const int valid = (int)min((long long)hist_len, cnt + 1);
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < valid; ++r) s += hist[(long long)r * k + c];
        qmean[c] = s / (float)valid;
    }
    __syncthreads();
    if (threadIdx.x == 0) *count = cnt + 1;
}

// ---------------------------------------------------------------------------------------------------- f-2: CLUBMean
// The reference forms a B x B x D broadcast ((y_j - mu_i)^2 averaged over j, club.py:113-118).  Algebraically
//   bound = mean_i(positive_i - negative_i) = sum_i mu_i.y_i / B - (sum_i mu_i).(sum_j y_j) / B^2
// (the ||y||^2 and ||mu||^2 terms cancel), so only column sums are needed: O(B D) work and no temporary.
// stats[4][D] = (sum_i mu, sum_i y, sum_i mu*y, sum_i (mu-y)^2): 32 columns per block, rows split over 8 groups and
// combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256) club_colstats_kernel(const void* __restrict__ mu, const void* __restrict__ y, int dtype,
                                                            long long ld, int rows, int dim, float* __restrict__ stats) {
    __shared__ float part[8][4][33];
    pdl_wait();
    pdl_launch_dependents();
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int grp = threadIdx.x >> 5;
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
    if (col < dim)
        for (int r = grp; r < rows; r += 8) {
            const float m = ld_as_float(mu, dtype, (long long)r * ld + col), v = ld_as_float(y, dtype, (long long)r * ld + col);
            a += m; b += v; c += m * v; d += (m - v) * (m - v);
        }
    part[grp][0][threadIdx.x & 31] = a; part[grp][1][threadIdx.x & 31] = b;
    part[grp][2][threadIdx.x & 31] = c; part[grp][3][threadIdx.x & 31] = d;
    __syncthreads();
    if (grp < 4 && col < dim) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += part[g][grp][threadIdx.x & 31];
        stats[(long long)grp * dim + col] = t;
    }
}
// bound (club.py:107-121) and learning loss (:125-130) from the column statistics; one block
__global__ void __launch_bounds__(256) club_losses_kernel(const float* __restrict__ stats, int rows, int dim, float* bound,
                                                          float* est) {
    __shared__ float red[kBlk / 32];
    pdl_wait();
    pdl_launch_dependents();
    float p = 0.f, ms = 0.f, q = 0.f;
    for (int c = threadIdx.x; c < dim; c += 256) {
        p += stats[2LL * dim + c];
        ms += stats[c] * stats[(long long)dim + c];
        q += stats[3LL * dim + c];
    }
    p = blk_reduce(p, red, false);
    ms = blk_reduce(ms, red, false);
    q = blk_reduce(q, red, false);
    if (threadIdx.x == 0) {
        const float b = (float)rows;
        if (bound) *bound = p / b - ms / (b * b);
        if (est) *est = q / b;
    }
}
// d bound / d mu_i = y_i/B - s/B^2, d bound / d y_j = mu_j/B - m/B^2;  d est / d mu = 2(mu - y)/B = -d est / d y
__global__ void __launch_bounds__(256) club_grad_kernel(const void* __restrict__ mu, const void* __restrict__ y, int dtype,
                                                        long long ld, int rows, int dim, const float* __restrict__ stats,
                                                        const float* g_bound, const float* g_est, float* d_mu, float* d_y,
                                                        long long ld_g) {
    pdl_wait();
    pdl_launch_dependents();
    const float gb = g_bound ? *g_bound : 0.f, ge = g_est ? *g_est : 0.f;
    const float ib = 1.0f / (float)rows;
    const long long total = (long long)rows * dim;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / dim), c = (int)(idx % dim);
        const float m = ld_as_float(mu, dtype, (long long)r * ld + c), v = ld_as_float(y, dtype, (long long)r * ld + c);
        const float e2 = ge * 2.f * (m - v) * ib;
        d_mu[(long long)r * ld_g + c] = gb * (v * ib - stats[(long long)dim + c] * ib * ib) + e2;
        d_y[(long long)r * ld_g + c] = gb * (m * ib - stats[c] * ib * ib) - e2;
    }
}

inline int warp_rows_block(int64_t rows) { return rows <= 4096 ? 64 : kBlk; }

}  // namespace

int launch_club_fwd(const void* mu, const void* y, int dtype, int64_t ld, int64_t rows, int64_t dim, float* stats, float* bound,
                    float* est, cudaStream_t stream) {
    STIL_CUDA(launch_pdl(club_colstats_kernel, dim3((unsigned)ceil_div(dim, 32)), dim3(256), 0, stream, mu, y, dtype,
                         (long long)ld, (int)rows, (int)dim, stats));
    STIL_CUDA(launch_pdl(club_losses_kernel, dim3(1), dim3(256), 0, stream, static_cast<const float*>(stats), (int)rows,
                         (int)dim, bound, est));
    return STIL_OK;
}
int launch_club_bwd(const void* mu, const void* y, int dtype, int64_t ld, int64_t rows, int64_t dim, const float* stats,
                    const float* g_bound, const float* g_est, float* d_mu, float* d_y, int64_t ld_g, cudaStream_t stream) {
    const int blocks = (int)std::min<int64_t>(ceil_div(rows * dim, 256), 148 * 8);
    STIL_CUDA(launch_pdl(club_grad_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, mu, y, dtype, (long long)ld, (int)rows,
                         (int)dim, stats, g_bound, g_est, d_mu, d_y, (long long)ld_g));
    return STIL_OK;
}

int launch_bank_softmax_rows(const float* z, int64_t ldz, int64_t rows, int64_t k_q, float temperature,
                             __nv_bfloat16* gop, int64_t ldg, int nseg, cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    STIL_CUDA(launch_pdl(bank_softmax_rows_kernel, dim3((unsigned)rows), dim3(kBlk), 0, stream, z, (long long)ldz, (int)k_q,
                         1.0f / temperature, gop, (long long)ldg, nseg));
    return STIL_OK;
}

int launch_smooth_mix(const float* p, int64_t ld_p, const float* s, int64_t ld_s, int64_t rows, int64_t k, float c_keep,
                      float c_bank, float* out, int64_t ld_out, float th, float* max_prob, int64_t* max_idx,
                      uint8_t* mask, cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    const int threads = warp_rows_block(rows);
    STIL_CUDA(launch_pdl(smooth_mix_kernel, dim3((unsigned)ceil_div(rows, threads / 32)), dim3(threads), 0, stream, p,
                         (long long)ld_p, s, (long long)ld_s, (int)rows, (int)k, c_keep, c_bank, out, (long long)ld_out, th,
                         max_prob, reinterpret_cast<long long*>(max_idx), mask));
    return STIL_OK;
}

int launch_sim_grad(const float* gsim, const float* sim, int64_t ld, int64_t rows, int64_t n_self, int64_t k_q,
                    float temperature, __nv_bfloat16* gs, int64_t ldg_s, __nv_bfloat16* gp, int64_t ldg_p, int nseg,
                    cudaStream_t stream) {
    if (rows == 0) return STIL_OK;
    const int64_t w = std::max(ldg_s, ldg_p);
    STIL_CUDA(launch_pdl(sim_grad_kernel, dim3((unsigned)ceil_div(w, kBlk), (unsigned)rows), dim3(kBlk), 0, stream, gsim, sim,
                         (long long)ld, (int)n_self, (int)k_q, 1.0f / temperature, gs, (long long)ldg_s, gp,
                         (long long)ldg_p, nseg));
    return STIL_OK;
}

int launch_graph_contrast(const float* Q, const float* sim, int64_t ld, int64_t rows, int64_t cols, float th, float* d_sim,
                          float grad_scale, float* partials, unsigned int* ticket, float* loss, cudaStream_t stream) {
    STIL_CUDA(launch_pdl(graph_contrast_kernel, dim3((unsigned)rows), dim3(kBlk), 0, stream, Q, sim, (long long)ld, (int)cols,
                         th, d_sim, grad_scale, partials, ticket, loss));
    return STIL_OK;
}

int64_t weighted_softce_blocks(int64_t rows) { return ceil_div(rows, warp_rows_block(rows) / 32); }

int launch_weighted_softce(const void* y, int dtype, int64_t ld_y, const float* tp, int64_t ld_t, const int64_t* tidx,
                           const uint8_t* mask, int64_t rows, int64_t k, float* d_y, int64_t ld_g, float grad_scale,
                           float* partials, unsigned int* ticket, float* loss, cudaStream_t stream) {
    const int threads = warp_rows_block(rows);
    STIL_CUDA(launch_pdl(weighted_softce_kernel, dim3((unsigned)weighted_softce_blocks(rows)), dim3(threads), 0, stream, y,
                         dtype, (long long)ld_y, tp, (long long)ld_t, reinterpret_cast<const long long*>(tidx), mask,
                         (int)rows, (int)k, d_y, (long long)ld_g, grad_scale, partials, ticket, loss));
    return STIL_OK;
}

int launch_queue_enqueue(void* queue, int q_dtype, int64_t ld_q, float* qprobs, int64_t ld_qp, int64_t k_q, int64_t* ptr,
                         const void* z, int z_dtype, int64_t ld_z, int64_t n, int64_t dim, const float* t, int64_t ld_t,
                         int64_t num_classes, cudaStream_t stream) {
    if (n == 0) return STIL_OK;
    const int64_t work = (dim + num_classes) * n;
    const int blocks = (int)std::min<int64_t>(ceil_div(work, kBlk), 148 * 8);
    queue_enqueue_kernel<<<blocks, kBlk, 0, stream>>>(queue, q_dtype, (long long)ld_q, qprobs, (long long)ld_qp, (int)k_q,
                                                      reinterpret_cast<const long long*>(ptr), z, z_dtype, (long long)ld_z,
                                                      (int)n, (int)dim, t, (long long)ld_t, (int)num_classes);
    STIL_LAUNCH_CHECK();
    queue_advance_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<long long*>(ptr), (int)n, (int)k_q);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_bank_update(void* bank, int b_dtype, int64_t ld_b, int64_t* labels, const void* k, int k_dtype, int64_t ld_k,
                       const int64_t* y, const int64_t* index, int64_t n, int64_t dim, cudaStream_t stream) {
    if (n == 0) return STIL_OK;
    const int blocks = (int)std::min<int64_t>(ceil_div(dim * n, kBlk), 148 * 8);
    bank_update_kernel<<<blocks, kBlk, 0, stream>>>(bank, b_dtype, (long long)ld_b, reinterpret_cast<long long*>(labels), k,
                                                    k_dtype, (long long)ld_k, reinterpret_cast<const long long*>(y),
                                                    reinterpret_cast<const long long*>(index), (int)n, (int)dim);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

// =====================================================================================================================
// f-4 — EMA teacher update (STiLModel.momentum_update_ema, STiLModel.py:154-168): the reference walks the state dict in
// Python and issues mul_ / add_ (or copy_) per tensor, every step.  Here: ONE launch over a device table of tensors cut into
// fixed-size chunks (a block per chunk).  Rounding mirrors the eager ops: ema*m, (1-m)*main and their sum are each rounded to
// the tensor's dtype (no FMA contraction), so fp32 results are bit-identical to the reference's.
// =====================================================================================================================
__global__ void __launch_bounds__(256) ema_update_kernel(const stil_ema_entry* __restrict__ table, const int* __restrict__ chunk_entry,
                                                         const long long* __restrict__ chunk_start, long long chunk_elems, float m,
                                                         float om) {
    const stil_ema_entry e = table[chunk_entry[blockIdx.x]];
    const long long i0 = chunk_start[blockIdx.x];
    const long long i1 = min((long long)e.numel, i0 + chunk_elems);
    if (e.kind == 1) {
        // plain copy (`num_batches_tracked`, :163-164): numel counts BYTES
        const unsigned char* src = static_cast<const unsigned char*>(e.main);
        unsigned char* dst = static_cast<unsigned char*>(e.ema);
        if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)i0) & 15) == 0) {
            long long i = i0 + 16LL * threadIdx.x;
            for (; i + 16 <= i1; i += 16LL * blockDim.x) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
            for (long long j = i0 + ((i1 - i0) / 16) * 16 + threadIdx.x; j < i1; j += blockDim.x) dst[j] = src[j];
        } else {
            for (long long j = i0 + threadIdx.x; j < i1; j += blockDim.x) dst[j] = src[j];
        }
        return;
    }
    if (e.dtype == STIL_F32) {
        float* dst = static_cast<float*>(e.ema);
        const float* src = static_cast<const float*>(e.main);
        if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && (i0 & 3) == 0) {
            long long i = i0 + 4LL * threadIdx.x;
            for (; i + 4 <= i1; i += 4LL * blockDim.x) {
                float4 a = *reinterpret_cast<const float4*>(dst + i);
                const float4 b = __ldg(reinterpret_cast<const float4*>(src + i));
                a.x = __fadd_rn(__fmul_rn(a.x, m), __fmul_rn(om, b.x));
                a.y = __fadd_rn(__fmul_rn(a.y, m), __fmul_rn(om, b.y));
                a.z = __fadd_rn(__fmul_rn(a.z, m), __fmul_rn(om, b.z));
                a.w = __fadd_rn(__fmul_rn(a.w, m), __fmul_rn(om, b.w));
                *reinterpret_cast<float4*>(dst + i) = a;
            }
            for (long long j = i0 + ((i1 - i0) / 4) * 4 + threadIdx.x; j < i1; j += blockDim.x)
                dst[j] = __fadd_rn(__fmul_rn(dst[j], m), __fmul_rn(om, src[j]));
        } else {
            for (long long j = i0 + threadIdx.x; j < i1; j += blockDim.x)
                dst[j] = __fadd_rn(__fmul_rn(dst[j], m), __fmul_rn(om, src[j]));
        }
    } else {
        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(e.ema);
        const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(e.main);
        for (long long j = i0 + threadIdx.x; j < i1; j += blockDim.x) {
            const float a = __bfloat162float(__float2bfloat16_rn(__fmul_rn(__bfloat162float(dst[j]), m)));
            const float b = __bfloat162float(__float2bfloat16_rn(__fmul_rn(om, __bfloat162float(src[j]))));
            dst[j] = __float2bfloat16_rn(__fadd_rn(a, b));
        }
    }
}

int launch_ema_update(const stil_ema_entry* table, const int32_t* chunk_entry, const int64_t* chunk_start, int64_t n_chunks,
                      int64_t chunk_elems, double momentum, cudaStream_t stream) {
    if (n_chunks == 0) return STIL_OK;
    // `(1. - self.momentum) * v_main`: the Python double is rounded to the tensor's computation type (fp32)
    const float om = (float)(1.0 - momentum);
    ema_update_kernel<<<(unsigned)n_chunks, 256, 0, stream>>>(table, chunk_entry, reinterpret_cast<const long long*>(chunk_start),
                                                              (long long)chunk_elems, (float)momentum, om);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

int launch_da_hist_update(const float* batch_mean, float* hist, int64_t hist_len, int64_t k, int64_t* count, float* qmean,
                          cudaStream_t stream) {
    da_hist_update_kernel<<<1, 256, 0, stream>>>(batch_mean, hist, (int)hist_len, (int)k, reinterpret_cast<long long*>(count),
                                                 qmean);
    STIL_LAUNCH_CHECK();
    return STIL_OK;
}

}  // namespace stil
