/* stil_head.h — C ABI of the B200-native STiL semi-supervised head (libstil_head.so).
 *
 * The reference (kgutjahr/STiL-TTA) is pure Python/PyTorch and has NO FFI for this path: the head
 * sits behind two nn.Module objects and inline tensor code.  The entry points
 * below are what a binding for that path binds; each cites the reference interface it replaces.
 * INTEGRATION.md shows the ctypes stub and the three-line change in the reference trainers.
 *
 * Conventions
 *   - every pointer is a raw DEVICE address unless stated; matrices are row-major with an explicit
 *     leading dimension `ld*` counted in elements; sizes are int64_t; no torch types anywhere.
 *   - `dtype` is a stil_dtype_t.  Embeddings/logits may be STIL_F32 or STIL_BF16 (bf16 values are
 *     consumed exactly; all arithmetic accumulates in fp32).  Embedding rows must be 16-byte
 *     aligned (dim % 8 == 0 for bf16, % 4 for f32, base pointers 16-B aligned).
 *   - every function returns 0 (STIL_OK) or a negative stil_status_t, never throws, never aborts,
 *     never synchronises the device; stil_last_error() gives the message (thread-local).
 *   - workspaces of stil_head_step must be zero-filled once before their first use (reduction tickets live there)
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are CUDA-graph
 *     capturable.  The library allocates no device memory: scratch comes from the caller through
 *     `workspace` (size from the matching *_workspace_bytes query; 256-B aligned).
 *   - kernels exist for sm_100a only; there is no CPU or generic-GPU fallback.
 */
#ifndef STIL_HEAD_H_
#define STIL_HEAD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STIL_VERSION 100 /* major*10000 + minor*100 + patch */

#if defined(__GNUC__)
#define STIL_API __attribute__((visibility("default")))
#else
#define STIL_API
#endif

typedef enum { STIL_F32 = 0, STIL_BF16 = 1 } stil_dtype_t;

typedef enum {
    STIL_OK = 0,
    STIL_E_SHAPE = -1, /* size out of the supported range */
    STIL_E_DTYPE = -2,
    STIL_E_ALIGN = -3, /* pointer / leading dimension alignment */
    STIL_E_ARCH = -4,  /* device is not sm_100 */
    STIL_E_CUDA = -5,  /* a CUDA runtime/driver call failed */
    STIL_E_ARG = -6,   /* null pointer, bad scalar (e.g. lambda_0 outside [0,1]) */
    STIL_E_WORKSPACE = -7
} stil_status_t;

STIL_API int stil_version(void);
/* sizeof(stil_head_step_args) (which = 0) / stil_p2p_channel (1) / stil_ema_entry (2): lets a binding check its struct mirror */
STIL_API int64_t stil_abi_struct_bytes(int which);
STIL_API const char* stil_last_error(void);
/* Debug aid: install (or clear with NULL) a device buffer of [64 launches][64 CTAs][8] uint64 into which the GEMM
 * kernel's CTAs store %globaltimer stamps of their phases (scripts/gemm_timeline.py).  Not for production use. */
STIL_API int stil_debug_trace(void* buffer);
/* Debug / measurement aid: 0 launches every kernel with plain stream serialisation instead of programmatic dependent launch
 * (a profiler's per-kernel duration then excludes the dependent's wait for its predecessor); 1 restores the default. */
STIL_API int stil_debug_pdl(int enable);
/* 0 if the current device can run the kernels (compute capability 10.x), STIL_E_ARCH otherwise. */
STIL_API int stil_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * a1 — DCC cross-modal InfoNCE.  Replaces CLIPLoss.forward, utils/clip_loss.py:27-40.
 *   a_loc, b_loc : this rank's rows [m, dim]     (out0 / out1 of the reference)
 *   a_all, b_all : all ranks' rows  [n, dim]     (== a_loc / b_loc and n == m when not distributed)
 *   row_offset   : global index of local row 0   (0 when not distributed)
 * fwd writes lse_row[m] (log-sum-exp over all n columns of the local rows of a·bT/T), lse_col[m]
 * (same for the local rows of b·aT/T, i.e. the local COLUMNS of the logits) and
 *   loss_sum[0] = sum_local (lambda0*(lse_row_i - d_i) + (1-lambda0)*(lse_col_i - d_i)) / n
 * which IS the reference loss when n == m, and sums to it over ranks otherwise.
 * `logits` (optional, may be NULL) receives the local rows of the [n-column] logits in fp32
 * (the reference returns them for validation top-k, STiLModel.py:437-438).
 * bwd needs the LSE vectors of ALL n rows/columns (all-gathered by the caller when distributed),
 * a device scalar grad_loss (NULL = 1.0) and writes d_a, d_b [m, dim] in `grad_dtype`.
 * Returns STIL_E_ARG if lambda0 is outside [0,1] (reference raises ValueError, clip_loss.py:22-23). */
STIL_API int64_t stil_infonce_workspace_bytes(int64_t m, int64_t n, int64_t dim, int dtype);
STIL_API int stil_infonce_fwd(const void* a_loc, const void* b_loc, const void* a_all, const void* b_all, int dtype,
                     int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset, float temperature,
                     float lambda0, float* loss_sum, float* lse_row, float* lse_col, float* logits,
                     int64_t ld_logits, void* workspace, int64_t workspace_bytes, void* stream);
STIL_API int stil_infonce_bwd(const void* a_loc, const void* b_loc, const void* a_all, const void* b_all, int dtype,
                     int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset, float temperature,
                     float lambda0, const float* lse_row_all, const float* lse_col_all, const float* grad_loss,
                     void* d_a, void* d_b, int grad_dtype, int64_t ld_grad, void* workspace,
                     int64_t workspace_bytes, void* stream);
/* Same as stil_infonce_bwd when it directly follows stil_infonce_fwd with the same inputs on the same, untouched,
 * workspace (the data-parallel head: forward, LSE exchange, backward): the inverse norms / operand split left there
 * by the forward are reused instead of recomputed. */
STIL_API int stil_infonce_bwd_after_fwd(const void* a_all, const void* b_all, int dtype, int64_t m, int64_t n, int64_t dim,
                                        int64_t ld, int64_t row_offset, float temperature, float lambda0,
                                        const float* lse_row_all, const float* lse_col_all, const float* grad_loss,
                                        void* d_a, void* d_b, int grad_dtype, int64_t ld_grad, void* workspace,
                                        int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3 (first half) — raw prototype similarities out[r,k] = feat[r,:]·prototypes[k,:] in fp32.
 * Replaces `feat_m_ue @ prototypes.t()`, STiLModel.py:293 (also :350).  prototypes is fp32 [k, dim]. */
STIL_API int64_t stil_proto_logits_workspace_bytes(int64_t rows, int64_t k, int64_t dim, int dtype);
STIL_API int stil_proto_logits(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* prototypes,
                      int64_t k, float* out, int64_t ld_out, void* workspace, int64_t workspace_bytes,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * a2 + a3 — CGPL consensus pseudo-labels and PGLS smoothing/threshold for the unlabelled rows.
 * Replaces the inline block STiLModel.py:262-298 (sharpen_predictions :195-196 with T=1).
 *   y_m,y_i,y_t      teacher logits [rows, k] (logit_dtype, leading dim ld_y)
 *   teacher_logits   fp32 [rows, k] from stil_proto_logits (divided by `temperature` here, :294)
 *   prediction_in    optional fp32 [rows, k] (leading dim ld_pin): the `prediction` of :276-277 when the caller has
 *                    distribution-aligned softmax(y_m) (DA == True, stil_da_apply); NULL = softmax(y_m), :279.
 *                    It replaces the thresholded distribution only — cases and pseudo label come from the logits.
 * Outputs (any may be NULL except pseudo_label/max_idx/mask1):
 *   pseudo_label [rows,k] f32 (:295), prediction [rows,k] f32 (:296), max_prob f32, max_idx i64 (:297),
 *   mask1 u8 (:298), case1/case2_i/case2_t/case3 u8 (:264-267), top1 i64 [3, rows] (:263, order m,i,t),
 *   cls i32 [rows] (= max_idx) and conf u8 (= mask1 when `past_start_epoch`, else 0 — the gate of
 *   :317-320 zeroes `prediction`, whose max is then 0 with argmax 0) for the prototype kernels.
 * Index/mask outputs follow torch semantics: first index among equal maxima, >= on fp32. */
STIL_API int stil_cgpl_pgls(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                   const float* teacher_logits, int64_t ld_t, int64_t rows, int64_t k, float temperature,
                   float rate_pseudo, float th1, int past_start_epoch, const float* prediction_in, int64_t ld_pin,
                   float* pseudo_label, int64_t ld_pl,
                   float* prediction, int64_t ld_pred, float* max_prob, int64_t* max_idx, uint8_t* mask1,
                   uint8_t* case1, uint8_t* case2_i, uint8_t* case2_t, uint8_t* case3, int64_t* top1,
                   int32_t* cls, uint8_t* conf, void* stream);

/* Row max / argmax / threshold of a dense soft label: (max_prob, max_id) = label.max(1); conf = max>=th.
 * Replaces utils/prototype_loss.py:31-32 and STiLModel.py:204-205. */
STIL_API int stil_label_argmax(const float* label, int64_t ld, int64_t rows, int64_t k, float threshold, int32_t* cls,
                      uint8_t* conf, float* max_prob, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4 — PGLS prototype loss.  Replaces PrototypeLoss.forward, utils/prototype_loss.py:24-40, with the
 * label already reduced to (cls, conf) by stil_label_argmax / stil_cgpl_pgls.
 *   loss[0] = -(1/rows) * sum_i conf_i * log(softmax(feat·protoT/T)[i, cls_i] + 1e-7)
 * fwd also writes lse[rows] and the backward coefficient w[rows] = conf_i/rows * p/(p+1e-7). */
STIL_API int64_t stil_proto_ce_workspace_bytes(int64_t rows, int64_t k, int64_t dim, int dtype);
STIL_API int stil_proto_ce_fwd(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* prototypes,
                      int64_t k, const int32_t* cls, const uint8_t* conf, float temperature, float* loss,
                      float* lse, float* w, void* workspace, int64_t workspace_bytes, void* stream);
STIL_API int stil_proto_ce_bwd(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* prototypes,
                      int64_t k, const int32_t* cls, const float* lse, const float* w, float temperature,
                      const float* grad_loss, void* d_feat, int grad_dtype, int64_t ld_grad, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5 — prototype bank partial sums.  Replaces cal_prototypes_separate (STiLModel.py:216-226, which
 * calls cal_prototypes :199-214 on rows [:b_l] and [b_l:]) and, when psum/pcount are given, the
 * accumulate of :380-381.  class_sum [k, dim] f32, class_count [k] f32 (fractional: labelled rows
 * weigh 1/repeat_ratio).  Pass b_l = 0, repeat_ratio = 1 for plain cal_prototypes.  Deterministic
 * (rows are added in index order; no atomics). */
STIL_API int stil_proto_accumulate(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const int32_t* cls,
                          const uint8_t* conf, int64_t b_l, float repeat_ratio, int64_t k, float* class_sum,
                          float* class_count, float* psum, float* pcount, void* stream);
/* prototypes_sum += class_sum; prototypes_count_sum += class_count  (STiLModel.py:380-381), for the
 * distributed path where the partials are all-reduced between the two calls (:377-379). */
STIL_API int stil_proto_add(const float* class_sum, const float* class_count, int64_t k, int64_t dim, float* psum,
                   float* pcount, void* stream);
/* Epoch end: prototypes = psum / pcount; psum = pcount = 0; *empty_classes = #{k: pcount_k < 1}
 * (device int; the reference asserts it is 0 on the host, STiLModel.py:408-415). */
STIL_API int stil_proto_finalize(float* prototypes, float* psum, float* pcount, int64_t k, int64_t dim,
                        int32_t* empty_classes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a6 — distribution alignment.  Replaces STiLModel.distribution_alignment (STiLModel.py:171-180; same code in
 * simmatch_model.py:151-163, MMatch.py:136-148), split at its all-reduce:
 *   stil_da_batch_mean : mean[k] = probs.mean(0)                                   (:173)
 *   (caller: all-reduce mean over ranks and divide by the world size, :174-176 — a no-op on one rank)
 *   stil_da_apply      : DA_queue[ptr] = mean; ptr = (ptr+1) % da_len; out = probs / DA_queue.mean(0), rows
 *                        renormalised (:176-179).  da_ptr is the reference's int64 [1] buffer ON THE DEVICE (no
 *                        host sync, unlike `int(self.DA_ptr)`); qmean_scratch is k floats of scratch. */
/* out = torch.softmax(logits, dim=1) in fp32 — the argument of distribution_alignment at STiLModel.py:277. */
STIL_API int stil_softmax_rows(const void* logits, int dtype, int64_t ld, int64_t rows, int64_t k, float* out, int64_t ld_out,
                               void* stream);
STIL_API int stil_da_batch_mean(const float* probs, int64_t ld, int64_t rows, int64_t k, float* mean, void* stream);
STIL_API int stil_da_apply(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* batch_mean, float* da_queue,
                  int64_t da_len, int64_t* da_ptr, float* qmean_scratch, float* out, int64_t ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a7 — SimMatch memory-bank block.  Replaces models/MatchModel/simmatch_model.py:268-286 (start_unlabel branch).
 *   bank    [dim, k_bank] in the REFERENCE layout (unit columns, :68-69), same dtype as the features, read in place
 *   labels  [k_bank] int64 (:70);  prob_ku_orig [rows, num_classes] f32 (after the optional DA, :264-266)
 * fwd writes prob_ku [rows, num_classes] (:280) and loss_in [rows] (:286, per row — the caller takes .mean(),
 * SimMatch.py:92) and leaves dLoss/dLogits (bf16) in the workspace; bwd turns grad_loss_in [rows] into
 * d_feat_qu [rows, dim] (only the student feature gets a gradient).  grad_loss_in == NULL means a unit upstream
 * gradient: d_feat_qu[i, :] = d loss_in[i] / d feat_qu[i, :].  The reference overwrites the bank right after this block
 * and before loss.backward() (simmatch_model.py:291, protected there by bank.clone(), :237), so a binding should run
 * bwd with NULL inside its forward and scale the rows by grad_loss_in later — the bank is then never read again.  The workspace must be the same, untouched,
 * buffer for the forward and its backward (it holds the [rows, k_bank] teacher/student logits, fp32). */
STIL_API int64_t stil_simmatch_workspace_bytes(int64_t rows, int64_t k_bank, int64_t dim, int dtype);
STIL_API int stil_simmatch_fwd(const void* feat_ku, const void* feat_qu, int dtype, int64_t rows, int64_t dim, int64_t ld,
                      const void* bank, int64_t ld_bank, const int64_t* labels, int64_t k_bank,
                      const float* prob_ku_orig, int64_t num_classes, float tt, float st, float c_smooth, float* prob_ku,
                      float* loss_in, int grad_dtype, void* workspace, int64_t workspace_bytes, void* stream);
STIL_API int stil_simmatch_bwd(const void* feat_qu, int dtype, int64_t rows, int64_t dim, const void* bank, int64_t ld_bank,
                      int64_t k_bank, const float* grad_loss_in, void* d_feat_qu, int grad_dtype, int64_t ld_grad,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* a7, bank column-sharded over the ranks of one node (BASELINE config C5: 65536 x 512 over 8 GPUs).  Every rank
 * holds bank[:, shard] (reference layout, [dim, k_shard]) with its labels and sweeps it for the GATHERED rows of all ranks:
 *   stil_simmatch_shard_stats : teacher / student logits against the shard, then per row, with the FIXED shift
 *       e = exp((z - 1)/T) (unit features and bank columns, simmatch_model.py:68-69: z <= 1), the additive statistics
 *       stats[row, 0:3+C] = [ sum e_t | sum e_s | sum e_t p[y_j] z_s/st | A_c = sum_{j in class c} e_t ]
 *   (caller: SUM stats over ranks — one all-reduce of [rows, 3+C] floats)
 *   stil_simmatch_shard_finish: totals -> prob_ku (:280), loss_in (:286) and norms[row] = {1/sum e_s, 1/den} for the gradient
 *   stil_simmatch_shard_grad  : G = (S - T')/st on the shard's columns (bf16 hi+lo) from the logits left in the workspace
 *       by _stats, and d_feat_partial [rows, dim] f32 = G · bank_shard^T — the caller reduce-scatters the partials so
 *       that every rank ends up with d loss_in[i] / d feat_qu[i, :] of its own rows.
 * Workspace: stil_simmatch_workspace_bytes(rows, k_shard, dim, dtype), the same untouched buffer for _stats and _grad. */
STIL_API int stil_simmatch_shard_stats(const void* feat_ku, const void* feat_qu, int dtype, int64_t rows, int64_t dim, int64_t ld,
                                       const void* bank, int64_t ld_bank, const int64_t* labels, int64_t k_shard,
                                       const float* prob_ku_orig, int64_t num_classes, float tt, float st, float* stats,
                                       void* workspace, int64_t workspace_bytes, void* stream);
STIL_API int stil_simmatch_shard_finish(const float* stats_total, const float* prob_ku_orig, int64_t rows, int64_t num_classes,
                                        float st, float c_smooth, float* prob_ku, float* loss_in, float* norms, void* stream);
STIL_API int stil_simmatch_shard_grad(int dtype, int64_t rows, int64_t dim, const void* bank, int64_t ld_bank,
                                      const int64_t* labels, int64_t k_shard, const float* prob_ku_orig, int64_t num_classes,
                                      float tt, float st, const float* norms, float* d_feat_partial, int64_t ld_grad,
                                      void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8 / a9 — memory-bank smoothing of a pseudo-label distribution.  Replaces comatch_model.py:288-293 (c_keep = alpha,
 * c_bank = 1 - alpha) and MMatch.py:222-227 (the literals 0.9 / 0.1), followed by the max / argmax / threshold of
 * MMatch.py:229-230 / CoMatch.py:92-93:
 *   A = rownorm(exp(feat · queue_feat / T));   out = c_keep * probs + c_bank * (A · queue_probsᵀ)
 *   max_prob, max_idx = max(out, dim=1);  mask = max_prob >= th            (each output pointer may be NULL)
 *   feat [rows, dim]; queue_feat [dim, k_q] in the REFERENCE layout, same dtype as feat, read in place;
 *   queue_probs [num_classes, k_q] f32.  queue_feat == NULL: no bank yet (epoch gate) — out = probs. */
STIL_API int64_t stil_bank_smooth_workspace_bytes(int64_t rows, int64_t k_q, int64_t dim, int64_t num_classes, int dtype);
STIL_API int stil_bank_smooth(const float* probs, int64_t ld_p, int64_t rows, int64_t num_classes, const void* feat, int dtype,
                              int64_t dim, int64_t ld_f, const void* queue_feat, int64_t ld_q, const float* queue_probs,
                              int64_t ld_qp, int64_t k_q, float temperature, float c_keep, float c_bank, float* out,
                              int64_t ld_out, float th, float* max_prob, int64_t* max_idx, uint8_t* mask, void* workspace,
                              int64_t workspace_bytes, void* stream);

/* a8 — CoMatch pseudo-label graph and embedding graph.  Replaces comatch_model.py:298-312:
 *   Q   = [probs · probsᵀ with diagonal 1 | probs · probs_u]                  [rows, rows + k_q] f32
 *   sim = exp([feat_s0 · feat_s1ᵀ | feat_s0 · queue_s] / T)                    [rows, rows + k_q] f32
 *   probs [rows, C] f32; probs_u [C, k_q] f32; queue_s [dim, k_q] (reference layouts), same dtype as the features.
 * stil_comatch_sim_bwd turns grad_sim into d_feat_s0 (the only differentiable input, :309-311); it has its own
 * workspace. */
STIL_API int64_t stil_comatch_graphs_workspace_bytes(int64_t rows, int64_t k_q, int64_t dim, int64_t num_classes, int dtype);
STIL_API int stil_comatch_graphs_fwd(const float* probs, int64_t ld_p, int64_t rows, int64_t num_classes,
                                     const float* probs_u, int64_t ld_pu, const void* feat_s0, const void* feat_s1, int dtype,
                                     int64_t dim, int64_t ld_f, const void* queue_s, int64_t ld_q, int64_t k_q,
                                     float temperature, float* Q, float* sim, int64_t ld_out, void* workspace,
                                     int64_t workspace_bytes, void* stream);
STIL_API int64_t stil_comatch_sim_bwd_workspace_bytes(int64_t rows, int64_t k_q, int64_t dim, int dtype);
STIL_API int stil_comatch_sim_bwd(const float* grad_sim, const float* sim, int64_t ld, int64_t rows, int64_t k_q,
                                  const void* feat_s1, int dtype, int64_t dim, int64_t ld_f, const void* queue_s,
                                  int64_t ld_q, float temperature, void* d_feat_s0, int grad_dtype, int64_t ld_grad,
                                  void* workspace, int64_t workspace_bytes, void* stream);

/* a8 consumer — CoMatch graph contrastive loss.  Replaces CoMatch.py:100-110:
 *   pos = Q >= contrast_th;  w = Q*pos / rowsum;  p = sim*pos / rowsum(sim);  loss = mean_i -sum_j w log(p + 1e-7) pos
 *   d_sim (NULL to skip) = d loss / d sim * grad_scale, [rows, ld].
 * stil_row_loss_workspace_bytes sizes the scratch of this and of stil_weighted_softce. */
STIL_API int64_t stil_row_loss_workspace_bytes(int64_t rows);
STIL_API int stil_graph_contrast_loss(const float* Q, const float* sim, int64_t ld, int64_t rows, int64_t cols,
                                      float contrast_th, float* loss, float* d_sim, float grad_scale, void* workspace,
                                      int64_t workspace_bytes, void* stream);

/* f-1 for the single-head consumers of a7-a9: loss = mean_i mask_i * CE(logits_i, target_i), forward + gradient.
 * Exactly one of target_probs [rows, k] f32 (SimMatch.py:91, CoMatch.py:96-97) and target_idx [rows] int64 (the dense
 * one-hot hard label of MMatch.py:231-234) is given; mask may be NULL (all rows); d_logits may be NULL. */
STIL_API int stil_weighted_softce(const void* logits, int logit_dtype, int64_t ld_y, const float* target_probs, int64_t ld_t,
                                  const int64_t* target_idx, const uint8_t* mask, int64_t rows, int64_t k, float* loss,
                                  float* d_logits, int64_t ld_g, float grad_scale, void* workspace, int64_t workspace_bytes,
                                  void* stream);

/* Queue / bank maintenance next to a7-a9 (SURVEY f-4).
 *   stil_queue_enqueue : queue_feat[:, ptr:ptr+n'] = z[:n'].T; queue_probs[:, ptr:ptr+n'] = t[:n'].T with
 *                        n' = min(n, k_q - ptr); ptr = (ptr + n') % k_q     (comatch_model.py:117-146, MMatch.py:102-117).
 *                        ptr is the reference's int64 [1] buffer ON THE DEVICE (no host sync, unlike `int(self.queue_ptr)`).
 *   stil_bank_update   : bank[:, index] = k.T; labels[index] = y             (simmatch_model.py:141-147)
 *   stil_da_apply_hist : CoMatch's list-based distribution alignment (comatch_model.py:271-285) as a ring of the last
 *                        hist_len batch means: hist[count % hist_len] = batch_mean; count += 1;
 *                        out = rownorm(probs / mean(valid rows of hist)); count is an int64 [1] on the device. */
STIL_API int stil_queue_enqueue(void* queue_feat, int q_dtype, int64_t ld_q, float* queue_probs, int64_t ld_qp, int64_t k_q,
                                int64_t* ptr, const void* z, int z_dtype, int64_t ld_z, int64_t n, int64_t dim, const float* t,
                                int64_t ld_t, int64_t num_classes, void* stream);
STIL_API int stil_bank_update(void* bank, int b_dtype, int64_t ld_bank, int64_t* labels, const void* k, int k_dtype,
                              int64_t ld_k, const int64_t* y, const int64_t* index, int64_t n, int64_t dim, void* stream);
STIL_API int stil_da_apply_hist(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* batch_mean, float* hist,
                                int64_t hist_len, int64_t* count, float* qmean_scratch, float* out, int64_t ld_out,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * f-3 — the Linear layers either side of the head: projector_imaging / projector_tabular = nn.Linear(512, 128) followed
 * by F.normalize (STiLModel.py:56-63, project_3features :182-192) and the classifier Linears that produce y_hat_m/i/t
 * (models/Disentangle/utils/STiLModel_backbone.py:66-68, 153-155).
 *   fwd: y = x · weightᵀ + bias  (weight [out_dim, in_dim] f32, bias [out_dim] f32 or NULL), fp32-accurate on the tensor
 *        cores (3 x bf16 split); with `normalize` (out_dim <= 128) the rows are L2-normalised in the GEMM epilogue like
 *        F.normalize (eps 1e-12): y receives the unit rows, y_raw [rows, out_dim] the biased product and inv_norm [rows]
 *        1 / max(||row||, eps) — both needed by bwd.
 *   bwd: d_y [rows, out_dim] f32 -> d_x [rows, in_dim] (grad_dtype; NULL to skip), d_weight [out_dim, in_dim] f32,
 *        d_bias [out_dim] f32 (each NULL to skip); y_raw / inv_norm from fwd when it normalised, NULL otherwise. */
STIL_API int64_t stil_linear_workspace_bytes(int64_t rows, int64_t in_dim, int64_t out_dim, int dtype);
STIL_API int stil_linear_fwd(const void* x, int dtype, int64_t rows, int64_t in_dim, int64_t ld_x, const float* weight,
                             const float* bias, int64_t out_dim, int normalize, float* y, int64_t ld_y, float* y_raw,
                             float* inv_norm, void* workspace, int64_t workspace_bytes, void* stream);
STIL_API int stil_linear_bwd(const void* x, int dtype, int64_t rows, int64_t in_dim, int64_t ld_x, const float* weight,
                             int64_t out_dim, const float* y_raw, const float* inv_norm, const float* d_y, int64_t ld_dy,
                             void* d_x, int grad_dtype, int64_t ld_dx, float* d_weight, float* d_bias, void* workspace,
                             int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f-4 — EMA teacher update.  Replaces STiLModel.momentum_update_ema (STiLModel.py:154-168), a per-tensor Python loop of
 * mul_ / add_ / copy_ over the whole state dict every step, by ONE launch over a device-resident table:
 *   kind 0: ema = ema * momentum + (1 - momentum) * main   (dtype STIL_F32 / STIL_BF16; each product and the sum rounded like
 *           the eager ops, so fp32 results are bit-identical to the reference)
 *   kind 1: ema = main, byte copy (`num_batches_tracked`, :163-164); numel counts BYTES
 * `momentum` is the Python double of the reference: m = (float)momentum and (float)(1.0 - momentum) are what the eager ops use.
 * `table` is a DEVICE array of n_entries stil_ema_entry; the tensors are cut into chunks of chunk_elems elements (bytes for
 * kind 1): chunk c covers elements [chunk_start[c], chunk_start[c] + chunk_elems) of entry chunk_entry[c] (both DEVICE arrays,
 * built once per model by the binding).  One block per chunk. */
typedef struct stil_ema_entry {
    void* ema;
    const void* main;
    int64_t numel;
    int32_t dtype; /* stil_dtype_t (kind 0) */
    int32_t kind;
} stil_ema_entry;
STIL_API int stil_ema_update(const stil_ema_entry* table, int64_t n_entries, const int32_t* chunk_entry,
                             const int64_t* chunk_start, int64_t n_chunks, int64_t chunk_elems, double momentum, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f-2 — CLUBMean mutual-information bound and its learning loss, from mu = p_mu(x_samples) on.  Replaces the tensor code
 * of models/Disentangle/utils/club.py:107-121 (forward) and :125-130 (learning_loss), called STiLModel.py:327-330.
 * The reference's B x B x D broadcast collapses to column sums (SURVEY f-2):
 *   bound = sum_i mu_i.y_i / B - (sum_i mu_i).(sum_j y_j) / B^2;   est = sum_i ||mu_i - y_i||^2 / B
 * colstats is 4*dim floats (kept for the backward).  bwd: g_bound / g_est are DEVICE scalars (NULL = 0);
 * d_mu, d_y [rows, ld_grad] f32 receive g_bound * d bound + g_est * d est. */
STIL_API int stil_club_fwd(const void* mu, const void* y, int dtype, int64_t rows, int64_t dim, int64_t ld, float* colstats,
                           float* bound, float* est, void* stream);
STIL_API int stil_club_bwd(const void* mu, const void* y, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* colstats,
                           const float* g_bound, const float* g_est, float* d_mu, float* d_y, int64_t ld_grad, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pseudo-label thresholds of the remaining baselines.
 *
 * FreeMatch self-adaptive threshold — replaces FreeMatchModel.update + .masking,
 * models/MatchModel/FreeMatchFolder/freematch_model.py:128-165.  Two calls so that a data-parallel caller can add the
 * statistics of all ranks in between (the reference all-gathers the probabilities, :129-130; the sums are the same thing):
 *   stil_freematch_stats        probabilities (softmax when is_logits, :153-157), row max / first arg max (:132,:161),
 *                               stats [2*C + 2] f32 = [ column sums | arg-max histogram | sum of row maxima | rows ]
 *   stil_freematch_update_mask  EMA update of the DEVICE-resident state time_p [1], p_model [C], label_hist [C] (:137-144,
 *                               clip_thresh != 0 clips time_p to [0, 0.95]) and mask [rows] f32 0/1 =
 *                               max_probs >= time_p * p_model[max_idx] / max(p_model) (:162-164)
 * max_probs / max_idx / probs_out may be NULL in stil_freematch_stats (kept in the workspace; pass the SAME workspace and
 * NULL again to stil_freematch_update_mask).
 *
 * stil_threshold_rows — CoTraining cross pseudo labels, models/SemiMultimodal/CoTraining.py:141-146: probs = softmax(logits),
 * max_probs / max_idx (may be NULL), mask = max_probs >= threshold (f32 0/1).
 *
 * FreeMatch fairness loss — replaces entropy_loss, FreeMatchFolder/freematch_utils.py:17-45 (rows with mask != 0 are
 * selected): loss [1], hist_mean [1] (= hist_s.mean()); bwd writes d loss / d logits_s * grad_loss[0] (NULL = 1) for every
 * row (zero for unselected rows) from what fwd left in the workspace. */
STIL_API int64_t stil_threshold_workspace_bytes(int64_t rows, int64_t num_classes);
STIL_API int stil_freematch_stats(const float* probs_or_logits, int64_t ld, int64_t rows, int64_t num_classes, int is_logits, float* stats,
                                  float* max_probs, int64_t* max_idx, float* probs_out, int64_t ld_probs, void* workspace,
                                  int64_t workspace_bytes, void* stream);
STIL_API int stil_freematch_update_mask(const float* stats_total, int64_t rows, int64_t num_classes, float momentum, float clip_thresh,
                                        float* time_p, float* p_model, float* label_hist, const float* max_probs, const int64_t* max_idx,
                                        float* mask, void* workspace, int64_t workspace_bytes, void* stream);
STIL_API int stil_threshold_rows(const float* logits, int64_t ld, int64_t rows, int64_t num_classes, float threshold, float* probs,
                                 int64_t ld_probs, float* max_probs, int64_t* max_idx, float* mask, void* stream);
STIL_API int stil_freematch_entropy_fwd(const float* mask, const float* logits_s, int64_t ld, int64_t rows, int64_t num_classes,
                                        const float* p_model, const float* label_hist, float* loss, float* hist_mean, void* workspace,
                                        int64_t workspace_bytes, void* stream);
STIL_API int stil_freematch_entropy_bwd(int64_t rows, int64_t num_classes, const float* grad_loss, float* d_logits_s, int64_t ld_grad,
                                        void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f-1 — masked soft-target CE of the three student heads on the unlabelled rows, forward and
 * gradient in one pass.  Replaces STiLModel.py:301-303.
 *   losses[3]  = (loss_m_u, loss_i_u, loss_t_u), each a mean over `rows`
 *   d_y_*      = d(loss_*)/d(y_*) * grad_scale   (NULL to skip), leading dim ld_g, fp32 */
STIL_API int64_t stil_masked_softce_workspace_bytes(int64_t rows);
STIL_API int stil_masked_softce(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                       const float* pseudo_label, int64_t ld_pl, const uint8_t* mask1, const uint8_t* case1,
                       const uint8_t* case2_i, const uint8_t* case2_t, const uint8_t* case3,
                       const uint8_t* mask_random, int64_t rows, int64_t k, float* losses, float* d_y_m,
                       float* d_y_i, float* d_y_t, int64_t ld_g, float grad_scale, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU exchange over peer-mapped memory (NVLink) for the two coupled steps of the data-parallel head (SURVEY
 * §8e; the reference's `dist.all_reduce` at STiLModel.py:377-379 and its `concat_all_gather` helpers).
 *   stil_p2p_alloc/free            one cudaMalloc'ed, zeroed buffer per rank (same size on every rank)
 *   stil_p2p_export / import/close 64-byte CUDA IPC handle of the buffer / peer mapping of another rank's buffer
 *   stil_p2p_exchange              all-gather: this rank's `nseg` segments are stored into EVERY rank's buffer at
 *       dst_offset[s] (the caller makes the offsets rank-specific), then arrival flags are exchanged and the call's
 *       kernel only finishes once every peer's segments have landed in the local buffer.  `bases` is a HOST array of
 *       the `world` buffer addresses as seen from this rank.  flags_offset: 8*8*8 bytes of u64 flags, ctrl_offset:
 *       64*8+64*4 bytes of local control words (both inside the buffer, zero-initialised); `channel` (0..7) separates
 *       independent exchanges of one step.  A peer that never arrives turns into a CUDA error (bounded spin).
 *       Callers must not overwrite a destination region a slower peer may still be reading: the head alternates
 *       between two regions on successive steps. */
STIL_API int stil_p2p_alloc(int64_t bytes, void** ptr);
STIL_API int stil_p2p_free(void* ptr);
STIL_API int stil_p2p_export(void* ptr, uint8_t* handle64);
STIL_API int stil_p2p_import(const uint8_t* handle64, void** peer_ptr);
STIL_API int stil_p2p_close(void* peer_ptr);
STIL_API int stil_p2p_exchange(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                               int channel, int nseg, const void* const* src, const int64_t* nbytes,
                               const int64_t* dst_offset, void* stream);

/* Fused compute + exchange schedule of the data-parallel head (no kernel sits waiting for a peer; consumers wait, tile by
 * tile, for exactly the arrivals they read).  All take the channel description of stil_p2p_exchange.
 *   stil_p2p_push            : stil_p2p_exchange without the final wait (store to every rank, publish the flag, retire)
 *   stil_p2p_wait            : retire once every peer's latest push on `channel` has landed (for plain consumers)
 *   stil_p2p_push_embeddings : rows of [feat_i | feat_t] (leading dimension 2*dim) and their inverse L2 norms written into
 *                              every rank's gathered matrix (byte offset ab_offset) and norm vectors (ra/rb_offset, f32
 *                              [n]) at global rows row0.. — pack + F.normalize pass + all-gather in one kernel
 *   stil_p2p_push_lse        : row LSEs of the local rows of both InfoNCE sides, merged from the statistics partials in
 *                              the stil_infonce workspace and written into every rank's gathered lse_row / lse_col as
 *                              8-byte "LL" words {f32 value, u32 tag} (ONE store each: no fence, no flag — readers spin
 *                              on the word until the tag equals the low 32 bits of channel tag_channel's counter, which
 *                              stil_p2p_push_embeddings advanced at the start of the step)
 *   stil_infonce_stats_gathered / _loss_gathered / _bwd_gathered : stil_infonce_fwd / _bwd on buffers gathered that way
 *       (bf16 only).  wait_flags = this rank's flag words of the channel (buffer + flags_offset + channel*64),
 *       wait_seq = its sequence counter (buffer + ctrl_offset + channel*8); rows_per_peer = rows each rank owns.
 *       stats: statistics GEMM only, every tile waits for the owner of its columns.  bwd: gradient GEMMs; lse_row_ll /
 *       lse_col_ll are the gathered LL-word vectors [n] x 8 bytes of stil_p2p_push_lse and ll_tag the counter word whose low
 *       32 bits tag this step — the epilogue spins on exactly the entries it reads.  loss: loss partial + LSEs of the
 *       local rows from the statistics (off the critical chain); with bases != NULL the partial is also stored as an LL
 *       word [rank] at byte loss_ll_offset of every rank's buffer (tag: low 32 bits of *ll_tag); its workspace ticket
 *       words (first 256 bytes) must start out zero. */
STIL_API int stil_p2p_push(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset, int channel,
                           int nseg, const void* const* src, const int64_t* nbytes, const int64_t* dst_offset, void* stream);
STIL_API int stil_p2p_wait(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset, int channel,
                           void* stream);
STIL_API int stil_p2p_push_embeddings(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                                      int channel, const void* feat_i, const void* feat_t, int dtype, int64_t rows,
                                      int64_t dim, int64_t row0, int64_t ab_offset, int64_t ra_offset, int64_t rb_offset,
                                      void* stream);
STIL_API int stil_p2p_push_lse(void* const* bases, int world, int rank, int64_t flags_offset, int64_t ctrl_offset,
                               int tag_channel, const void* infonce_workspace, int64_t m, int64_t n, int64_t dim, int dtype,
                               int64_t row0, int64_t lse_row_offset, int64_t lse_col_offset, void* stream);
STIL_API int stil_infonce_stats_gathered(const void* a_all, const void* b_all, const float* ra_all, const float* rb_all,
                                         int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset,
                                         float temperature, const void* wait_flags, const void* wait_seq,
                                         int64_t rows_per_peer, void* workspace, int64_t workspace_bytes, void* stream);
STIL_API int stil_infonce_loss_gathered(const void* a_all, const void* b_all, const float* ra_all, const float* rb_all,
                                        int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset,
                                        float temperature, float lambda0, float* loss_sum, float* lse_row, float* lse_col,
                                        void* const* bases, int world, int rank, int64_t loss_ll_offset, const void* ll_tag,
                                        void* workspace, int64_t workspace_bytes, void* stream);
/* stil_proto_add_gathered for partials that were PUSHED (stil_p2p_push / stil_head_step's partials_push): every block
 * first waits for all ranks' arrival counters (wait_flags[w] >= *wait_target); optionally also sums the ranks' InfoNCE
 * loss partials (LL words loss_ll[w], tag = low 32 bits of *loss_tag) into loss_out — the last kernel of the step. */
STIL_API int stil_proto_add_gathered_wait(const float* parts, int64_t world, int64_t slot_floats, int64_t k, int64_t dim,
                                          float* class_sum, float* class_count, float* psum, float* pcount,
                                          const void* wait_flags, const void* wait_target, const void* loss_ll,
                                          const void* loss_tag, float* loss_out, void* stream);
STIL_API int stil_infonce_bwd_gathered(const void* a_all, const void* b_all, const float* ra_all, const float* rb_all,
                                       int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset,
                                       float temperature, float lambda0, const void* lse_row_ll,
                                       const void* lse_col_ll, const void* ll_tag, const float* grad_loss, void* d_a, void* d_b, int grad_dtype,
                                       int64_t ld_grad, void* workspace, int64_t workspace_bytes, void* stream);

/* prototypes_sum += sum_w parts[w].class_sum; prototypes_count_sum += sum_w parts[w].class_count, with the ranks added
 * in index order (deterministic); parts is the gathered [world][k*dim + k] buffer; also writes the reduced
 * class_sum / class_count.  The distributed form of STiLModel.py:377-381. */
STIL_API int stil_proto_add_gathered(const float* parts, int64_t world, int64_t slot_floats, int64_t k, int64_t dim,
                                     float* class_sum, float* class_count, float* psum, float* pcount, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The whole per-batch head in one call (what STiLModel.training_step lines 262-303, 317-322, 339,
 * 374-381 do), with launches batched across the sub-problems.  Used by bench.py and STiLHead.step. */
/* one channel of a peer-memory buffer set (see stil_p2p_exchange) */
typedef struct stil_p2p_channel {
    void* bases[8];
    int world, rank;
    int64_t flags_offset, ctrl_offset;
    int channel;
} stil_p2p_channel;

typedef struct stil_head_step_args {
    /* sizes */
    int64_t batch, b_l, k, dim;
    int embed_dtype, logit_dtype, grad_dtype;
    /* hyper-parameters (configs/config_dvm_STiL.yaml) */
    float temperature, lambda0, th1, rate_pseudo, repeat_ratio;
    int past_start_epoch;
    /* inputs */
    const void *feat_i, *feat_t, *feat_m, *feat_m_e; /* [batch, dim], ld = dim */
    const void *y_m_ue, *y_i_ue, *y_t_ue;             /* teacher logits [b_u, k] */
    const void *y_m, *y_i, *y_t;                      /* student logits [batch, k] (may be NULL: skip f-1) */
    const int64_t* y_l;                               /* [b_l] */
    const float* prototypes;                          /* [k, dim] */
    const uint8_t* mask_random;                       /* [b_u] */
    /* outputs */
    float* losses;                                    /* [5]: itc, pt, m_u, i_u, t_u */
    void *d_feat_i, *d_feat_t, *d_feat_m;             /* [batch, dim] grad_dtype */
    float *d_y_m, *d_y_i, *d_y_t;                     /* [batch, k] f32 (rows < b_l are zero) */
    float* pseudo_label;                              /* [b_u, k] */
    float* max_prob; int64_t* max_idx; uint8_t *mask1, *case1, *case2_i, *case2_t, *case3;
    float *class_sum, *class_count;                   /* [k, dim], [k] */
    float *prototypes_sum, *prototypes_count_sum;     /* accumulated in place when non-NULL */
    float rate_uce_scale;                             /* grad_scale of the f-1 gradients */
    void* workspace; int64_t workspace_bytes; void* stream;
    /* optional instrumentation (bench.py): cudaEvent_t[11] recorded around the launches of the two chains —
     * 0|prep|1|gemm stats (prototypes)|5|cgpl_pgls|6|gemm grad (prototype CE)|7|gemm dX (prototype CE)|8 on `stream`;
     * 9|prep|10|gemm stats (InfoNCE)|2|gemm grad (InfoNCE)|3|gemm dX (InfoNCE)|4 on the internal InfoNCE stream.  Leave NULL/0 otherwise (and
     * always during graph capture). */
    void** timing_events; int n_timing_events;
    /* 1: leave the InfoNCE (losses[0], d_feat_i, d_feat_t) to the caller — the data-parallel path computes it on
     * the all-gathered global batch with stil_infonce_fwd/bwd while this call does everything row-local */
    int skip_infonce;
    /* 1: the bf16 operand form of `prototypes` is already in the workspace (stil_head_prepare_prototypes was called
     * with the same workspace after the prototypes last changed) — the step then skips that conversion */
    int prototypes_prepared;
    /* data-parallel head: non-NULL = right after the prototype partial sums exist, push the packed
     * [class_sum | class_count] (class_count must sit right behind class_sum) to byte partials_dst_offset of every
     * rank's buffer on this channel (stil_p2p_push); the consumer is stil_proto_add_gathered_wait */
    const stil_p2p_channel* partials_push;
    int64_t partials_dst_offset;
    /* optional [b_u, k] f32 (ld = k): distribution-aligned softmax(y_m_ue) of STiLModel.py:276-277 (DA == True);
     * NULL = the kernel's own softmax(y_m_ue), :279 */
    const float* prediction_in;
} stil_head_step_args;
STIL_API int64_t stil_head_step_workspace_bytes(int64_t batch, int64_t b_l, int64_t k, int64_t dim, int embed_dtype);
STIL_API int stil_head_step(const stil_head_step_args* args);
/* Convert args->prototypes to the tensor-core operand form kept in args->workspace (the prototypes only change at
 * epoch end, STiLModel.py:408-415, so this runs once per change, not once per step). */
STIL_API int stil_head_prepare_prototypes(const stil_head_step_args* args);
/* number of kernel launches one stil_head_step enqueues for these args (for bench.py's gpu_launches) */
STIL_API int stil_head_step_launches(const stil_head_step_args* args);

#ifdef __cplusplus
}
#endif
#endif /* STIL_HEAD_H_ */
