// Internal launcher interfaces shared by the translation units of libstil_head.so.
#pragma once
#include "common.cuh"

namespace stil {

// --------------------------------------------------------------------------- tcgen05 GEMM (gemm_tc05.cu)
// One kernel computes 128x128 tiles of  L = alpha * diag(sx) * (X · Yᵀ) * diag(sy)  on tcgen05 with the
// accumulator in TMEM, for up to kMaxGemmJobs independent problems per launch.  X and Y are bf16
// "operand tensors" [rows, nseg, D]: an fp32 matrix is carried as a sum of bf16 segments (hi, lo, lolo)
// and the product is accumulated over the listed (xseg, yseg) pairs in fp32 — fp32-accurate results
// from bf16 tensor-core passes.
enum GemmMode : int {
    GEMM_STATS = 0,  // per-row online (max, sum-exp) partial per column tile  [+ optional fp32 store]
    GEMM_STORE = 1,  // fp32 store of L
    GEMM_GRAD = 2,   // G = u_i e^{L-lse_x[i]} + v e^{L-lse_y[j]} - d_i [j == tgt_i], times cs_j, as bf16 hi/lo
    GEMM_BWD = 3     // fused backward: per 128-row block, for each column tile: recompute L, form G in shared memory, and
                     // accumulate dX += G · Y in TMEM (G never leaves the SM); see gemm_bwd_kernel
};
constexpr int kMaxGemmJobs = 4;
constexpr int kMaxSegPairs = 6;
constexpr int kTileM = 128, kTileN = 128, kTileK = 64;

struct alignas(64) GemmJob {
    CUtensorMap tmx, tmy;  // 3-D maps (D, rows, nseg), box 64 x 128 x 1, 128-B swizzle
    CUtensorMap tmg;       // GRAD: bulk-store map of gop [ld_g, M, g_nseg] (filled by launch_gemm)
    int M, N, D;           // X rows, Y rows, per-segment contraction length
    int npair;
    int xseg[kMaxSegPairs], yseg[kMaxSegPairs];
    int mode;
    int tiles_m, tiles_n, tile_begin;
    float alpha;
    const float* sx;  // [M] row scale or nullptr
    const float* sy;  // [N] column scale or nullptr
    // STATS
    float* part_max;  // [tiles_n, M]
    float* part_sum;
    // STATS, single pass over a SYMMETRIC problem (the InfoNCE's a·bᵀ: the columns' statistics are the rows' statistics of
    // b·aᵀ): besides the row partials the tile also emits COLUMN partials [4 * tiles_m, N] — the sum of e^{L - sym_shift}
    // over each 32-row quarter of the tile (warp-shuffle transpose-reduce), so b·aᵀ is never computed.  All partials of such
    // a job use the FIXED shift sym_shift >= max L (unit rows: |L| <= alpha) instead of a running maximum, which makes them
    // plain sums; the max arrays receive the constant so that every consumer's merge code stays as it is.
    float* cpart_max;
    float* cpart_sum;
    float sym_shift;
    // GRAD: G additionally carries the row scale sx[i] (one G for both dX products of the symmetric problem);
    // STORE: the row scale is 1 / sx[i] (undoes it)
    int g_row_scale;
    int sx_recip;
    // STORE / STATS(optional)
    float* out;
    long long ld_out;
    // set by launch_gemm when `out` allows it (16-byte aligned base, ld_out % 4 == 0): the fp32 tile leaves through shared
    // memory and TMA bulk tensor stores (tmg = map of out [N, M, slices], 32 x 128 boxes, 128-byte swizzle) instead of
    // per-lane 4-byte stores
    int out_tma;
    // GRAD
    const float* lse_x;      // [M]
    const float* lse_y;      // [N] or nullptr (v ignored)
    const float* u_vec;      // [M] or nullptr -> u_scalar, d_scalar
    float u_scalar, v_scalar, d_scalar;
    const int* tgt_vec;      // [M] or nullptr -> tgt = row + tgt_offset
    int tgt_offset;
    const float* gscale;     // device scalar or nullptr (1.0)
    __nv_bfloat16* gop;      // [M, g_nseg, ld_g]
    long long ld_g;
    // GRAD with the statistics merge fused in (used when lse_x == nullptr): per-column-tile (max, sum)
    // partials of GEMM_STATS for this job's rows (px_*) and, for the symmetric InfoNCE, columns (py_*)
    const float *px_max, *px_sum, *py_max, *py_sum;
    int px_tiles, py_tiles;
    // GRAD for the prototype CE: u_i = d_i = conf_i*w_coef*p/(p+1e-7), p = exp(z[i, tgt_i] - lse_i), with z the
    // logits stored by the GEMM_STATS job of the same problem
    const float* w_z;        // [M, w_ldz] (nullptr -> u/d as above)
    long long w_ldz;
    const unsigned char* w_conf;
    float w_coef;
    int g_nseg;              // 2: G as bf16 hi+lo (fp32 gradients), 1: hi only (bf16 gradients)
    // STORE: split the contraction over `ksplit` CTAs per tile; partial tiles are added to `out` with red.global
    // (out must be zero on entry).  1 = plain store.  With slice_stride > 0 (floats) slice k instead STORES its partial
    // tile to out + k*slice_stride — deterministic, nothing to clear; the consumer adds the slices in index order.
    int ksplit;
    long long slice_stride;
    // STORE with Y given MN-major ([contraction rows, N contiguous]) — the backward's dX = G · Y
    int y_mn_major;
    // STORE with X given MN-major ([contraction rows, M contiguous]): out = Xᵀ · Y, the contraction runs over the ROWS of both
    // row-major matrices (a Linear layer's dW = gᵀ · x, STiLModel.py:56-63; tmx then has inner = M, box 64 x 64)
    int x_mn_major;
    // STORE epilogue extras of the forward of a Linear layer (f-3): + col_bias[j]; fwd_norm: rows L2-normalised like
    // F.normalize (needs tiles_n == 1), 1 / max(||row||, 1e-12) written to fwd_inv_norm and the un-normalised row to fwd_raw
    const float* col_bias;
    int fwd_norm;
    float* fwd_inv_norm;
    float* fwd_raw;
    long long fwd_ld_raw;
    // STORE with the normalise-backward / cast fused in (needs tiles_n == 1):
    //   dx = fin_sx ? sx*(g - xh*(xh·g)), xh = sx*x : g      written in fin_dx_dtype
    void* fin_dx;            // nullptr -> plain fp32 store to `out`
    int fin_dx_dtype;
    long long fin_ld_dx;
    const void* fin_x;
    int fin_x_dtype;
    long long fin_ldx;
    const float* fin_sx;
    // programmatic dependent launch: the X / Y operand bytes are not written by the kernel launched immediately before
    // this one on the stream, so the TMA producer may stream them before griddepcontrol.wait (see gemm_tc05.cu)
    int early_x, early_y;
    int early_stats;   // GRAD: the px_* statistics partials are that old too: merge them before the wait
    // Peer-memory gather consumed in place (data-parallel head): wait_flags[p] (local, one u64 per peer) must reach
    // *wait_seq before rows [p*wait_rows_per_peer, ...) of the gathered Y-side data are read — by the TMA producer for
    // the Y operand tile (wait_y) and by the epilogue for sy / lse_y.  nullptr = no waiting.
    const unsigned long long* wait_flags;
    const unsigned long long* wait_seq;
    int wait_rows_per_peer;
    int wait_y;
    // GRAD: lse_x / lse_y are arrays of 8-byte LL words {f32 value, u32 tag} (flag-less peer exchange); an entry is
    // valid once its tag equals the low 32 bits of *lse_ll_tag.  nullptr = plain f32 arrays.
    const unsigned long long* lse_ll_tag;
    // plain STORE post-op on the scaled value: 0 none, 1 exp, 2 diagonal (row == column) forced to 1
    int post_op;
    // GEMM_BWD: second product dX += G · Y (segment pairs of G x Y), column tiles dealt round-robin to `nsplit` CTAs per row
    // block (nsplit > 1: per-slice partial dX like ksplit slices), shared-memory plan (bytes from the 1024-aligned base)
    int npair2;
    int gseg2[kMaxSegPairs], yseg2[kMaxSegPairs];
    int nsplit;
    int bw_nx, bw_ny;          // X / Y segments resident in shared memory
    int bw_nbuf;               // Y tile buffers (2 when they fit: next tile's recompute overlaps this tile's epilogue)
    int bw_y_off, bw_y_stride, bw_g_off, bw_misc_off;
};

struct GemmLaunch {
    GemmJob job[kMaxGemmJobs];
    int njobs;
    int total_tiles;
    // GEMM_STORE with a split contraction reduced ON CHIP: the `cluster_k` CTAs that share a tile (every job's ksplit must
    // equal cluster_k, 2..8) are launched as one thread-block cluster; the non-leaders publish their raw partial accumulator
    // in their own shared memory, the leader adds them through distributed shared memory into its TMEM accumulator and runs the
    // normal epilogue (fused normalise-backward included) — no partial tiles in global memory, no reduction kernel.  0/1 = off.
    int cluster_k;
    unsigned int trace_id;   // launch ordinal for the optional timeline trace
    unsigned long long* trace;   // nullptr unless stil_debug_trace installed a buffer
};

// Build the operand tensor map.  `base` bf16, logical [rows, nseg, inner]; strides in elements.
int make_operand_map(CUtensorMap* tm, const void* base, int64_t inner, int64_t rows, int64_t nseg,
                     int64_t row_stride, int64_t seg_stride, int box_rows = 128, int box_segs = 1);
// fp32 output map [cols, rows, slices] for bulk tensor stores: 32 x 128 x 1 boxes, 128-byte swizzle; false = not eligible
bool make_out_map(CUtensorMap* tm, float* base, int64_t cols, int64_t rows, int64_t ld, int64_t slices, int64_t slice_stride);
void gemm_job_tiles(GemmLaunch& L);  // fills tiles_m/tiles_n/tile_begin/total_tiles
int launch_gemm(const GemmLaunch& L, cudaStream_t stream);
// a7 bank sweep (bank_sweep.cu): zt = feat_ku · bank, zs = feat_qu · bank by the persistent resident-rows kernel
bool bank_logits_eligible(int dtype, int64_t rows, int64_t dim, int64_t k_shard, int64_t ldz);
int launch_bank_logits(const void* feat_ku, const void* feat_qu, int64_t rows, int64_t dim, int64_t ld, const void* bank,
                       int64_t ld_bank, int64_t k_shard, float* zt, float* zs, int64_t ldz, cudaStream_t stream);
// out[rows, dim] = row_scale * (G[rows, (hi, lo), k_shard] · bankᵀ): whole-accumulator-in-TMEM split contraction
bool bank_dx_eligible(int dtype, int64_t rows, int64_t dim, int64_t k_shard, const float* out, int64_t ld_out);
int launch_bank_dx(const __nv_bfloat16* gop, int64_t ldg, int g_nseg, int64_t rows, const void* bank, int64_t ld_bank, int64_t dim,
                   int64_t k_shard, const float* row_scale, float* out, int64_t ld_out, cudaStream_t stream, bool out_is_zero = false);
// Turn a GEMM_GRAD job + the GEMM_STORE job that would consume its G into one GEMM_BWD job (false: shapes / shared memory
// do not allow the fused kernel — launch the two separately)
bool make_bwd_job(GemmJob& out, const GemmJob& grad, const GemmJob& store, int nsplit);
int bwd_nsplit(int64_t n_cols, int64_t row_blocks_total);
int gemm_set_trace(void* buf);   // debug: device buffer of [64 launches][64 CTAs][8] u64 %globaltimer stamps, or nullptr

// (max, sum) statistics partials of the two InfoNCE sides inside an stil_infonce workspace (api.cu; used by p2p.cu)
void infonce_stat_partials(void* workspace, int64_t m, int64_t n, int64_t dim, int dtype, const float** pmax,
                           const float** psum, int* slots);

// --------------------------------------------------------------------------- row kernels (row_kernels.cu)
// Operand preparation: for every row of x (f32 or bf16) write the bf16 segments (hi[,lo[,lolo]]) into
// op [rows, nseg, dim] and/or the transposed operand op_t [dim, nseg, ld_t] (only first `nseg_t` segments),
// and the inverse L2 norm 1/max(||x||, eps) (F.normalize semantics, clip_loss.py:29-30).
struct PrepJob {
    const void* x;
    int dtype;
    int rows, dim;
    long long ld;
    int nseg;
    __nv_bfloat16* op;       // nullable
    __nv_bfloat16* op_t;     // nullable
    long long ld_t;
    int nseg_t;
    float* inv_norm;         // nullable
    int op_dim;              // per-segment row length of `op` (>= dim, zero-filled beyond dim); 0 = dim
    int block_begin;
};
constexpr int kMaxPrepJobs = 8;
struct PrepLaunch {
    PrepJob job[kMaxPrepJobs];
    int njobs;
    int total_blocks;
    unsigned int* zero_words;  // words zeroed by block 0 (reduction tickets), n_zero <= 256
    int n_zero;
};
void prep_add(PrepLaunch& L, const PrepJob& j);
int launch_prep(const PrepLaunch& L, cudaStream_t stream);

// Merge the per-column-tile (max, sum) partials of GEMM_STATS into row LSEs and loss terms.
struct FinishJob {
    int kind;  // 0 = InfoNCE side, 1 = prototype CE
    int M;     // rows
    int tiles_n;
    const float* part_max;
    const float* part_sum;
    float* lse;  // [M] out
    // kind 0: diag term d_i = alpha*sx_i*sy_{i+off}*<x_i, y_{i+off}>, rowterm = coef*(lse_i - d_i)
    // kind 1: z = alpha*<x_i, proto_{cls_i}>, p = e^{z-lse}, rowterm = -conf_i*log(p+1e-7)*coef, w_i = conf_i*coef*p/(p+1e-7)
    const void* x;
    int x_dtype;
    long long ldx;
    const void* y;
    int y_dtype;
    long long ldy;
    int y_offset;
    int dim;
    const float* sx;
    const float* sy;
    float alpha, coef;
    const int* cls;
    const unsigned char* conf;
    float* w;
    int loss_slot;  // which of out_loss[] this job adds into
    int row_begin;
};
constexpr int kMaxFinishJobs = 4;
struct FinishLaunch {
    FinishJob job[kMaxFinishJobs];
    int njobs;
    int total_rows;
    float* block_partials;   // [blocks * 2]
    unsigned int* ticket;    // zeroed once by the caller's workspace init, self-resetting
    float* out_loss;         // [2]
    // optional: loss slot 0 also goes to every rank as one 8-byte LL word {f32, u32 tag} (data-parallel head)
    unsigned long long* ll_out[8];
    int ll_world;
    const unsigned long long* ll_tag;
};
int launch_finish(const FinishLaunch& L, cudaStream_t stream);
int64_t finish_blocks(int total_rows);

// dx = sx*(g - xh*(xh·g)) with xh = sx*x  (backward of F.normalize) or dx = g when sx == nullptr.
struct GradFinishJob {
    const float* g;  // [nslices][rows, dim] f32, slices `slice_stride` floats apart, added in index order
    int nslices;     // 0/1 = a single matrix
    long long slice_stride;
    const void* x;
    int x_dtype;
    long long ldx;
    const float* sx;
    void* dx;
    int dx_dtype;
    long long ld_dx;
    int rows, dim;
    int row_begin;
};
struct GradFinishLaunch {
    GradFinishJob job[4];
    int njobs;
    int total_rows;
};
int launch_grad_finish(const GradFinishLaunch& L, cudaStream_t stream);

int launch_cgpl_pgls(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                     const float* teacher_logits, int64_t ld_t, int64_t rows, int64_t k, float temperature,
                     float rate_pseudo, float th1, int past_start_epoch, const float* prediction_in, int64_t ld_pin,
                     float* pseudo_label, int64_t ld_pl,
                     float* prediction, int64_t ld_pred, float* max_prob, int64_t* max_idx, uint8_t* mask1,
                     uint8_t* case1, uint8_t* case2_i, uint8_t* case2_t, uint8_t* case3, int64_t* top1,
                     int32_t* cls, uint8_t* conf, const int64_t* y_l, int64_t b_l, int32_t* cls_l, uint8_t* conf_l,
                     cudaStream_t stream);
int launch_label_argmax(const float* label, int64_t ld, int64_t rows, int64_t k, float threshold, int32_t* cls,
                        uint8_t* conf, float* max_prob, cudaStream_t stream);
// cls/conf of the labelled rows: cls = y_l, conf = (1 >= th) (a one-hot row has max 1, STiLModel.py:321)
int launch_labelled_cls(const int64_t* y_l, int64_t b_l, float th, int32_t* cls, uint8_t* conf, cudaStream_t stream);
int launch_proto_accumulate(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const int32_t* cls,
                            const uint8_t* conf, int64_t b_l, float repeat_ratio, int64_t k, float* class_sum,
                            float* class_count, float* psum, float* pcount, cudaStream_t stream);
int launch_proto_add(const float* class_sum, const float* class_count, int64_t k, int64_t dim, float* psum,
                     float* pcount, cudaStream_t stream);
int launch_proto_add_gathered(const float* parts, int64_t world, int64_t slot, int64_t k, int64_t dim, float* class_sum,
                              float* class_count, float* psum, float* pcount, cudaStream_t stream,
                              const unsigned long long* wait_flags = nullptr, const unsigned long long* wait_target = nullptr,
                              const unsigned long long* loss_ll = nullptr, const unsigned long long* loss_tag = nullptr,
                              float* loss_out = nullptr);
int launch_proto_finalize(float* prototypes, float* psum, float* pcount, int64_t k, int64_t dim,
                          int32_t* empty_classes, cudaStream_t stream);
int launch_masked_softce(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                         const float* pseudo_label, int64_t ld_pl, const uint8_t* mask1, const uint8_t* case1,
                         const uint8_t* case2_i, const uint8_t* case2_t, const uint8_t* case3,
                         const uint8_t* mask_random, int64_t rows, int64_t k, float* losses, float* d_y_m,
                         float* d_y_i, float* d_y_t, int64_t ld_g, float grad_scale, float* block_partials,
                         unsigned int* ticket, cudaStream_t stream);
int64_t masked_softce_blocks(int64_t rows, int64_t k);
int launch_zero_u32(unsigned int* p, int n, cudaStream_t stream);
// SimMatch bank rows (simmatch_model.py:268-286) on materialised teacher / student logits
int launch_simmatch_rows(const float* zt, const float* zs, long long ldz, const long long* labels, int rows, int k_bank,
                         const float* p_orig, int num_classes, float tt, float st, float c_smooth, float* p_out,
                         float* loss_in, __nv_bfloat16* gop, long long ld_g, int g_nseg, cudaStream_t stream);
int launch_simmatch_shard_stats(const float* zt, const float* zs, long long ldz, const long long* labels, int rows, int k_shard,
                                const float* p_all, int num_classes, float tt, float st, float* stats, float* chunk_scratch,
                                cudaStream_t stream);
int simmatch_shard_chunks(int64_t rows, int64_t k_shard);
int launch_simmatch_shard_finish(const float* stats, const float* p_all, int rows, int num_classes, float st, float c_smooth,
                                 float* p_out, float* loss_in, float* norms, cudaStream_t stream);
int launch_simmatch_shard_grad(const float* zt, const float* zs, long long ldz, const long long* labels, int rows, int k_shard,
                               const float* p_all, int num_classes, float tt, float st, const float* norms, __nv_bfloat16* gop,
                               long long ld_g, int g_nseg, float* zero_out, long long zero_n, cudaStream_t stream);
int launch_softmax_rows(const void* y, int dtype, int64_t ld, int64_t rows, int64_t k, float* out, int64_t ld_out,
                        cudaStream_t stream);
int launch_col_sum(const float* x, int64_t ld, int64_t rows, int64_t k, float* out, cudaStream_t stream);
int launch_da_batch_mean(const float* probs, int64_t ld, int64_t rows, int64_t k, float* mean, cudaStream_t stream);
int launch_da_apply(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* batch_mean, float* da_queue,
                    int64_t da_len, int64_t* da_ptr, float* qmean, float* out, int64_t ld_out, cudaStream_t stream);

// --------------------------------------------------------------------------- bank blocks a8/a9 (bank_kernels.cu)
int launch_bank_softmax_rows(const float* z, int64_t ldz, int64_t rows, int64_t k_q, float temperature,
                             __nv_bfloat16* gop, int64_t ldg, int nseg, cudaStream_t stream);
int launch_smooth_mix(const float* p, int64_t ld_p, const float* s, int64_t ld_s, int64_t rows, int64_t k, float c_keep,
                      float c_bank, float* out, int64_t ld_out, float th, float* max_prob, int64_t* max_idx,
                      uint8_t* mask, cudaStream_t stream);
int launch_sim_grad(const float* gsim, const float* sim, int64_t ld, int64_t rows, int64_t n_self, int64_t k_q,
                    float temperature, __nv_bfloat16* gs, int64_t ldg_s, __nv_bfloat16* gp, int64_t ldg_p, int nseg,
                    cudaStream_t stream);
int launch_graph_contrast(const float* Q, const float* sim, int64_t ld, int64_t rows, int64_t cols, float th, float* d_sim,
                          float grad_scale, float* partials, unsigned int* ticket, float* loss, cudaStream_t stream);
int64_t weighted_softce_blocks(int64_t rows);
int launch_weighted_softce(const void* y, int dtype, int64_t ld_y, const float* tp, int64_t ld_t, const int64_t* tidx,
                           const uint8_t* mask, int64_t rows, int64_t k, float* d_y, int64_t ld_g, float grad_scale,
                           float* partials, unsigned int* ticket, float* loss, cudaStream_t stream);
int launch_queue_enqueue(void* queue, int q_dtype, int64_t ld_q, float* qprobs, int64_t ld_qp, int64_t k_q, int64_t* ptr,
                         const void* z, int z_dtype, int64_t ld_z, int64_t n, int64_t dim, const float* t, int64_t ld_t,
                         int64_t num_classes, cudaStream_t stream);
int launch_bank_update(void* bank, int b_dtype, int64_t ld_b, int64_t* labels, const void* k, int k_dtype, int64_t ld_k,
                       const int64_t* y, const int64_t* index, int64_t n, int64_t dim, cudaStream_t stream);
int launch_ema_update(const stil_ema_entry* table, const int32_t* chunk_entry, const int64_t* chunk_start, int64_t n_chunks,
                      int64_t chunk_elems, double momentum, cudaStream_t stream);
int launch_da_hist_update(const float* batch_mean, float* hist, int64_t hist_len, int64_t k, int64_t* count, float* qmean,
                          cudaStream_t stream);
int launch_club_fwd(const void* mu, const void* y, int dtype, int64_t ld, int64_t rows, int64_t dim, float* stats, float* bound,
                    float* est, cudaStream_t stream);
int launch_club_bwd(const void* mu, const void* y, int dtype, int64_t ld, int64_t rows, int64_t dim, const float* stats,
                    const float* g_bound, const float* g_est, float* d_mu, float* d_y, int64_t ld_g, cudaStream_t stream);
// FreeMatch / CoTraining thresholds (threshold_kernels.cu)
int64_t threshold_workspace_bytes(int64_t rows, int64_t C);
int launch_freematch_stats(const float* x, int64_t ld, int64_t rows, int64_t C, int is_logits, float* stats, float* max_p,
                           int64_t* max_i, float* probs_out, int64_t ld_probs, void* workspace, int64_t workspace_bytes,
                           cudaStream_t stream);
int launch_freematch_update_mask(const float* stats_total, int64_t rows, int64_t C, float momentum, float clip_thresh, float* time_p,
                                 float* p_model, float* label_hist, const float* max_p, const int64_t* max_i, float* mask,
                                 void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int launch_threshold_rows(const float* logits, int64_t ld, int64_t rows, int64_t C, float threshold, float* probs, int64_t ld_probs,
                          float* max_p, int64_t* max_i, float* mask, cudaStream_t stream);
int launch_freematch_entropy_fwd(const float* mask, const float* logits_s, int64_t ld, int64_t rows, int64_t C, const float* p_model,
                                 const float* label_hist, float* loss, float* hist_mean, void* workspace, int64_t workspace_bytes,
                                 cudaStream_t stream);
int launch_freematch_entropy_bwd(int64_t rows, int64_t C, const float* grad_loss, float* d_logits, int64_t ld_g, void* workspace,
                                 int64_t workspace_bytes, cudaStream_t stream);
// probs / qmean with rows renormalised (the second half of launch_da_apply)
int launch_da_rows(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* qmean, float* out, int64_t ld_out,
                   cudaStream_t stream);

}  // namespace stil
