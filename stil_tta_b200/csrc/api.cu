// extern "C" entry points of libstil_head.so (see include/stil_head.h).  Host-side composition only:
// plans the caller-provided workspace, builds TMA descriptors and job tables, enqueues kernels.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "internal.h"

namespace stil {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

namespace {

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

int check_embed(const void* p, int dtype, int64_t rows, int64_t dim, int64_t ld, const char* what) {
    STIL_REQUIRE(dtype == STIL_F32 || dtype == STIL_BF16, STIL_E_DTYPE, "%s: dtype %d not in {f32, bf16}", what, dtype);
    STIL_REQUIRE(rows >= 0 && rows < (1LL << 30) && dim >= 1 && dim <= 16384, STIL_E_SHAPE,
                 "%s: unsupported shape [%lld, %lld]", what, (long long)rows, (long long)dim);
    STIL_REQUIRE(p != nullptr || rows == 0, STIL_E_ARG, "%s: null pointer", what);
    const int per16 = dtype == STIL_BF16 ? 8 : 4;
    STIL_REQUIRE(dim % per16 == 0 && ld % per16 == 0 && ld >= dim, STIL_E_ALIGN,
                 "%s: dim %lld / ld %lld must be multiples of %d elements (16-byte rows)", what, (long long)dim,
                 (long long)ld, per16);
    STIL_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, STIL_E_ALIGN, "%s: pointer not 16-byte aligned", what);
    return STIL_OK;
}

// A matrix presented to the tensor cores: bf16 inputs are used in place (1 segment), fp32 inputs are
// carried as `nseg` bf16 segments written by the prep kernel.
struct Operand {
    const __nv_bfloat16* base;  // [rows, nseg, dim]
    int nseg;
    int64_t row_stride, seg_stride;
};

inline int64_t pad32(int64_t n) { return round_up(n, 32); }
// GEMM_STATS writes one (max, sum) partial per 32-column chunk of a tile
inline int64_t stat_slots(int64_t n) { return 4 * ceil_div(n, kTileN); }

// segment pairs (x_seg, y_seg) with x_seg + y_seg <= order, most significant first.  Segment s carries
// ~2^-9s of the value, so order 2 keeps every product above ~2^-26 (fp32-accurate), order 1 above 2^-17.
int seg_pairs(int nx, int ny, int order, int* xs, int* ys) {
    int n = 0;
    for (int o = 0; o <= order; ++o)
        for (int a = 0; a <= o; ++a) {
            const int b = o - a;
            if (a < nx && b < ny && n < kMaxSegPairs) { xs[n] = a; ys[n] = b; ++n; }
        }
    return n;
}

int fill_gemm_common(GemmJob& J, const Operand& X, int64_t xrow0, int64_t M, const Operand& Y, int64_t N, int64_t D,
                     int order = 2) {
    std::memset(&J, 0, sizeof(J));
    int rc = make_operand_map(&J.tmx, X.base + xrow0 * X.row_stride, D, M, X.nseg, X.row_stride, X.seg_stride);
    if (rc) return rc;
    rc = make_operand_map(&J.tmy, Y.base, D, N, Y.nseg, Y.row_stride, Y.seg_stride);
    if (rc) return rc;
    J.M = (int)M; J.N = (int)N; J.D = (int)D;
    J.npair = seg_pairs(X.nseg, Y.nseg, order, J.xseg, J.yseg);
    J.alpha = 1.f;
    return STIL_OK;
}

// Mark operands the kernel launched immediately before this one (on the same stream, by this library) does not write:
// the TMA producer then streams them before griddepcontrol.wait (gemm_tc05.cu).  Never set for the first kernel of a
// call — its stream predecessor is the caller's.
inline void set_early(GemmLaunch& GL, bool x, bool y) {
    for (int j = 0; j < GL.njobs; ++j) {
        GL.job[j].early_x = x ? 1 : 0;
        GL.job[j].early_y = y ? 1 : 0;
    }
}

// One fused launch (gemm_bwd_kernel) instead of a GRAD launch followed by the STORE launch that consumes its G, when every
// job pair allows it; the column tiles of a row block are dealt to as many CTAs as the STORE would have had slices.
bool fuse_bwd(GemmLaunch& GB, const GemmLaunch& GG, const GemmLaunch& GS) {
    std::memset(&GB, 0, sizeof(GB));
    if (GG.njobs != GS.njobs) return false;
    for (int j = 0; j < GG.njobs; ++j)
        if (!make_bwd_job(GB.job[j], GG.job[j], GS.job[j], GS.job[j].ksplit > 1 ? GS.job[j].ksplit : 1)) return false;
    GB.njobs = GG.njobs;
    gemm_job_tiles(GB);
    return true;
}

// dX[M, ncols] = G[M, (hi,lo), n] · Y[n, ncols]: X = G (K-major), Y read in place as an MN-major operand
int fill_gemm_store_mn(GemmJob& J, const Operand& G, int64_t M, const Operand& Y, int64_t n, int64_t ncols) {
    std::memset(&J, 0, sizeof(J));
    int rc = make_operand_map(&J.tmx, G.base, n, M, G.nseg, G.row_stride, G.seg_stride, 128);
    if (rc) return rc;
    rc = make_operand_map(&J.tmy, Y.base, ncols, n, Y.nseg, Y.row_stride, Y.seg_stride, 64);
    if (rc) return rc;
    J.M = (int)M; J.N = (int)ncols; J.D = (int)n;
    J.npair = seg_pairs(G.nseg, Y.nseg, 1, J.xseg, J.yseg);
    J.alpha = 1.f;
    J.mode = GEMM_STORE;
    J.y_mn_major = 1;
    return STIL_OK;
}

int infonce_dx_ksplit(int64_t m, int64_t n, int64_t dim);

// ------------------------------------------------------------------------------------------ InfoNCE plan
struct InfoncePlan {
    int nseg;
    __nv_bfloat16 *a_op, *b_op;                 // [n, nseg, dim] (fp32 inputs only)
    float *ra, *rb;                             // inverse norms [n]
    float *pmax[2], *psum[2];                   // [tiles_n, m]
    float* block_partials;
    unsigned int* ticket;
    __nv_bfloat16* gop[2];                      // [m, 2, ldg]
    int64_t ldg;
    float* g[2];                                // [m, dim]
    int64_t bytes;
};

InfoncePlan plan_infonce(void* ws, int64_t ws_bytes, int64_t m, int64_t n, int64_t dim, int dtype, bool bwd) {
    InfoncePlan P;
    Workspace W(ws, ws_bytes);
    P.nseg = dtype == STIL_BF16 ? 1 : 3;
    P.ldg = pad32(n);
    P.ticket = W.take<unsigned int>(64);
    P.ra = W.take<float>(n);
    P.rb = W.take<float>(n);
    P.a_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(n * P.nseg * dim);
    P.b_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(n * P.nseg * dim);
    for (int s = 0; s < 2; ++s) {
        P.pmax[s] = W.take<float>(stat_slots(n) * m);
        P.psum[s] = W.take<float>(stat_slots(n) * m);
    }
    P.block_partials = W.take<float>(2 * ceil_div(3 * m, 2) + 8);   // upper bound of finish_blocks():  // 3m rows: the fused step adds the prototype rows
    // backward-only regions (the forward never touches them, the query always counts them)
    for (int s = 0; s < 2; ++s) {
        P.gop[s] = W.take<__nv_bfloat16>(m * 2 * P.ldg);
        P.g[s] = W.take<float>(m * dim * infonce_dx_ksplit(m, n, dim));   // one slice per contraction split
    }
    (void)bwd;
    P.bytes = W.off;
    return P;
}

Operand rowmajor_operand(const void* x, int dtype, int64_t dim, int64_t ld, const __nv_bfloat16* op, int nseg) {
    Operand O;
    if (dtype == STIL_BF16) {
        O.base = static_cast<const __nv_bfloat16*>(x);
        O.nseg = 1;
        O.row_stride = ld;
        O.seg_stride = dim;  // unused (single segment) but must be 16-byte granular
    } else {
        O.base = op;
        O.nseg = nseg;
        O.row_stride = (int64_t)nseg * dim;
        O.seg_stride = dim;
    }
    return O;
}
// G = dLoss/dLogits as bf16: hi+lo segments when the gradients are wanted in fp32, hi only for bf16 gradients
// (whose own rounding is 2^-9 already)
inline int grad_nseg(int grad_dtype) { return grad_dtype == STIL_BF16 ? 1 : 2; }
Operand grad_operand(const __nv_bfloat16* gop, int64_t ldg, int nseg) {
    Operand O;
    O.base = gop;
    O.nseg = nseg;
    O.row_stride = (int64_t)nseg * ldg;
    O.seg_stride = ldg;
    return O;
}

// ------------------------------------------------------------------------------------------ proto plan
struct ProtoPlan {
    int feat_nseg, proto_nseg;
    __nv_bfloat16 *feat_op, *proto_op;
    float *pmax, *psum;
    float* block_partials;
    unsigned int* ticket;
    __nv_bfloat16* gop;
    int64_t ldg;
    float* g;
    int64_t bytes;
};

// The fused backward of the prototype CE gives every column tile of a row block its own CTA (the three fp32-split
// prototype segments leave room for one Y tile buffer only, so a CTA walking several tiles would serialise TMA, recompute,
// epilogue and dX): partial dX tiles go to slices that grad_finish adds.
inline int64_t proto_bwd_slices(int64_t k) {
    const int64_t tiles = ceil_div(k, kTileN);
    return tiles <= 8 ? tiles : 1;
}

ProtoPlan plan_proto(void* ws, int64_t ws_bytes, int64_t rows, int64_t k, int64_t dim, int dtype) {
    ProtoPlan P;
    Workspace W(ws, ws_bytes);
    P.feat_nseg = dtype == STIL_BF16 ? 1 : 3;
    P.proto_nseg = 3;
    P.ldg = pad32(k);
    P.ticket = W.take<unsigned int>(64);
    P.feat_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(rows * P.feat_nseg * dim);
    P.proto_op = W.take<__nv_bfloat16>(k * P.proto_nseg * dim);
    P.pmax = W.take<float>(stat_slots(k) * rows);
    P.psum = W.take<float>(stat_slots(k) * rows);
    P.block_partials = W.take<float>(2 * ceil_div(rows, 2) + 8);
    P.gop = W.take<__nv_bfloat16>(rows * 2 * P.ldg);
    P.g = W.take<float>(rows * dim * proto_bwd_slices(k));   // one slice per column tile for the fused backward
    P.bytes = W.off;
    return P;
}

PrepJob prep_job(const void* x, int dtype, int64_t rows, int64_t dim, int64_t ld, int nseg, __nv_bfloat16* op,
                 __nv_bfloat16* op_t, int64_t ld_t, int nseg_t, float* inv_norm) {
    PrepJob j;
    std::memset(&j, 0, sizeof(j));
    j.x = x; j.dtype = dtype; j.rows = (int)rows; j.dim = (int)dim; j.ld = ld;
    j.nseg = nseg; j.op = op; j.op_t = op_t; j.ld_t = ld_t; j.nseg_t = nseg_t; j.inv_norm = inv_norm;
    return j;
}

int infonce_args_check(const void* a_all, const void* b_all, int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld,
                       int64_t row_offset, float temperature, float lambda0) {
    int rc = check_embed(a_all, dtype, n, dim, ld, "infonce a");
    if (rc) return rc;
    rc = check_embed(b_all, dtype, n, dim, ld, "infonce b");
    if (rc) return rc;
    STIL_REQUIRE(lambda0 >= 0.f && lambda0 <= 1.f, STIL_E_ARG, "lambda_0 must be a float between 0 and 1.");
    STIL_REQUIRE(temperature > 0.f, STIL_E_ARG, "temperature must be positive");
    STIL_REQUIRE(m >= 0 && row_offset >= 0 && row_offset + m <= n, STIL_E_SHAPE,
                 "local rows [%lld, %lld) outside the %lld global rows", (long long)row_offset,
                 (long long)(row_offset + m), (long long)n);
    return STIL_OK;
}

// jobs shared by the op-level entry points and stil_head_step --------------------------------------
// The local InfoNCE (every row sees every column: m == n) is computed in a SINGLE pass over a·bᵀ: each tile emits row AND
// column statistics (the columns' statistics are the rows' statistics of b·aᵀ), the backward forms dLoss/dLogits once and
// feeds both dA = G·B and dB = Gᵀ·A from it (the second with G read as an MN-major A operand).  Executed tensor work: 2 + 2
// + 2x2 products of n^2 d instead of 4 + 4 + 2x2.  The fixed shift needs |logit| <= 1/T to stay in fp32 range: T >= 1/40.
// The data-parallel head (m < n) keeps the two-sided form: each rank owns the ROWS of both sides, so nothing is reduced
// across ranks.  STIL_NCE_TWO_SIDED=1 forces the two-sided form (A/B measurements).
bool infonce_single_pass(int64_t m, int64_t n, int64_t off, float inv_t) {
    static const bool two_sided = [] { const char* e = getenv("STIL_NCE_TWO_SIDED"); return e && e[0] == '1'; }();
    return !two_sided && m == n && off == 0 && inv_t <= 40.f;
}

// returns the number of jobs (1: single pass, 2: one per side) through *njobs
int infonce_stats_jobs(GemmJob* J2, const InfoncePlan& P, const Operand& A, const Operand& B, int64_t m, int64_t n,
                       int64_t dim, int64_t off, float inv_t, float* logits, int64_t ld_logits, int* njobs) {
    *njobs = 2;
    if (infonce_single_pass(m, n, off, inv_t)) {
        int rc = fill_gemm_common(J2[0], A, 0, m, B, n, dim);
        if (rc) return rc;
        GemmJob& J = J2[0];
        J.mode = GEMM_STATS;
        J.alpha = inv_t;
        J.sx = P.ra; J.sy = P.rb;
        J.part_max = P.pmax[0]; J.part_sum = P.psum[0];
        J.cpart_max = P.pmax[1]; J.cpart_sum = P.psum[1];      // same [4 * tiles, n] layout as side 1's row partials
        J.sym_shift = inv_t;
        if (logits) { J.out = logits; J.ld_out = ld_logits; }
        *njobs = 1;
        return STIL_OK;
    }
    for (int s = 0; s < 2; ++s) {
        const Operand& X = s == 0 ? A : B;
        const Operand& Y = s == 0 ? B : A;
        int rc = fill_gemm_common(J2[s], X, off, m, Y, n, dim);
        if (rc) return rc;
        J2[s].mode = GEMM_STATS;
        J2[s].alpha = inv_t;
        J2[s].sx = (s == 0 ? P.ra : P.rb) + off;
        J2[s].sy = s == 0 ? P.rb : P.ra;
        J2[s].part_max = P.pmax[s];
        J2[s].part_sum = P.psum[s];
        if (s == 0 && logits) {
            J2[s].out = logits;
            J2[s].ld_out = ld_logits;
        }
    }
    return STIL_OK;
}

void infonce_finish_jobs(FinishJob* F2, const InfoncePlan& P, const void* a_all, const void* b_all, int dtype,
                         int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t off, float inv_t, float lambda0,
                         float* lse_row, float* lse_col, int loss_slot) {
    const int esz = dtype == STIL_BF16 ? 2 : 4;
    for (int s = 0; s < 2; ++s) {
        FinishJob& F = F2[s];
        std::memset(&F, 0, sizeof(F));
        F.kind = 0;
        F.M = (int)m;
        F.tiles_n = (int)stat_slots(n);
        F.part_max = P.pmax[s];
        F.part_sum = P.psum[s];
        F.lse = s == 0 ? lse_row : lse_col;
        const char* xa = static_cast<const char*>(s == 0 ? a_all : b_all);
        F.x = xa + off * ld * esz;
        F.x_dtype = dtype; F.ldx = ld;
        F.y = s == 0 ? b_all : a_all;
        F.y_dtype = dtype; F.ldy = ld;
        F.y_offset = (int)off;
        F.dim = (int)dim;
        F.sx = (s == 0 ? P.ra : P.rb) + off;
        F.sy = s == 0 ? P.rb : P.ra;
        F.alpha = inv_t;
        F.coef = (s == 0 ? lambda0 : 1.f - lambda0) / (float)n;
        F.loss_slot = loss_slot;
    }
}

int infonce_grad_jobs(GemmJob* J2, const InfoncePlan& P, const Operand& A, const Operand& B, int64_t m, int64_t n,
                      int64_t dim, int64_t off, float inv_t, float lambda0, const float* lse_row_all,
                      const float* lse_col_all, const float* grad_loss, int grad_dtype, int* njobs) {
    *njobs = infonce_single_pass(m, n, off, inv_t) ? 1 : 2;
    for (int s = 0; s < *njobs; ++s) {
        const Operand& X = s == 0 ? A : B;
        const Operand& Y = s == 0 ? B : A;
        int rc = fill_gemm_common(J2[s], X, off, m, Y, n, dim);
        if (rc) return rc;
        GemmJob& J = J2[s];
        J.mode = GEMM_GRAD;
        J.alpha = inv_t;
        J.sx = (s == 0 ? P.ra : P.rb) + off;
        J.sy = s == 0 ? P.rb : P.ra;
        if (lse_row_all) {
            J.lse_x = (s == 0 ? lse_row_all : lse_col_all) + off;
            J.lse_y = s == 0 ? lse_col_all : lse_row_all;
        } else {
            // single-process step: merge the GEMM_STATS partials in the kernel (rows: this side, columns: the
            // other side's rows) instead of waiting for the finish kernel
            J.px_max = P.pmax[s]; J.px_sum = P.psum[s];
            J.py_max = P.pmax[1 - s]; J.py_sum = P.psum[1 - s];
            J.px_tiles = J.py_tiles = (int)stat_slots(n);
        }
        J.u_scalar = (s == 0 ? lambda0 : 1.f - lambda0) / (float)n;
        J.v_scalar = (s == 0 ? 1.f - lambda0 : lambda0) / (float)n;
        J.d_scalar = 1.f / (float)n;
        J.tgt_offset = (int)off;
        J.gscale = grad_loss;
        J.gop = P.gop[s];
        J.ld_g = P.ldg;
        J.g_nseg = grad_nseg(grad_dtype);
        // single pass: ONE G = dLoss/dLogits * alpha * sx_i * sy_j serves both products; each undoes "its" scale in the epilogue
        J.g_row_scale = *njobs == 1 ? 1 : 0;
    }
    return STIL_OK;
}

// The dX GEMM of a global batch contracts over all n columns on only 2*ceil(m/128)*ceil(dim/128) tiles: split the
// contraction over CTAs (partial tiles are added with 16-byte reductions into P.g, which the GRAD launch cleared)
// when the loop is long and the grid small.
int infonce_dx_ksplit(int64_t m, int64_t n, int64_t dim) {
    const int64_t kblocks = ceil_div(n, kTileK), tiles = 2 * ceil_div(m, kTileM) * ceil_div(dim, kTileN);
    static const int64_t min_kblocks = [] { const char* e = getenv("STIL_DX_SPLIT_MIN_KBLOCKS"); return e ? atoll(e) : 32ll; }();
    if (kblocks < min_kblocks || tiles * 2 > 148) return 1;
    return (int)std::max<int64_t>(1, std::min<int64_t>(kblocks / 8, 148 / tiles));
}

// Cluster split-K of a dX GEMM whose tiles span `dim` (<= 128 columns, fused epilogue): `tiles` CTAs would each walk
// `kblocks` operand stages at L2 -> shared-memory bandwidth; 2 / 4 / 8 CTAs per tile share the walk and the leader adds the
// partial accumulators through distributed shared memory (gemm_tc05.cu).  1 = not worth it / not possible.
int dx_cluster_k(int64_t tiles, int64_t kblocks) {
    // STIL_DX_CLUSTER=0/1: never (long contractions fall back to slice outputs + the reduction kernel), 2/4/8: always
    static const int forced = [] { const char* e = getenv("STIL_DX_CLUSTER"); return e ? atoi(e) : -1; }();
    if (forced >= 0) return forced <= 1 ? 1 : forced;
    // Measured (profiles/r2_cluster_splitk.txt): reading a 64 KB partial tile through distributed shared memory costs ~1.5 us
    // per peer even with a coalesced (column-major) layout, and a cluster needs all its SMs free in ONE GPC at the same time.
    // On the C2 step clusters of 4-8 made the dX GEMMs 2-3x slower; on one rank's global-batch backward (512 x 4096 x 128)
    // slice outputs + the reduction kernel (43 us) beat clusters of 4 / 8 (49 us).  A PAIR wins only in the narrow band where
    // the walk is long enough to share (>= 32 stages per tile) but too short for the slice path (n = 1024: 34.6 vs 36.6 us).
    int ck = 1;
    if (tiles * 4 <= 148 && kblocks >= 32) ck = 2;
    return ck;
}

// d(x̂_i) = sum_j G'_ij y_j, then the backward of F.normalize in the epilogue when one tile spans `dim`
int infonce_store_jobs(GemmJob* J2, const InfoncePlan& P, const Operand& A, const Operand& B, const void* a_all,
                       const void* b_all, int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t off,
                       void* d_a, void* d_b, int grad_dtype, int64_t ld_grad, float inv_t, bool* fused, int* cluster_k) {
    const int esz = dtype == STIL_BF16 ? 2 : 4;
    int ksplit = infonce_dx_ksplit(m, n, dim);
    *cluster_k = 1;
    if (dim <= kTileN) {
        // one tile spans the embedding: split the contraction over a CLUSTER and keep the fused epilogue
        // ... unless the contraction is long enough for the slice path (ksplit > 1 above)
        const int ck = ksplit > 1 ? 1 : dx_cluster_k(2 * ceil_div(m, kTileM), ceil_div(n, kTileK) * grad_nseg(grad_dtype) * B.nseg);
        if (ck > 1) ksplit = *cluster_k = (int)std::min<int64_t>(ck, ceil_div(n, kTileK));
    }
    *fused = dim <= kTileN && (ksplit == 1 || *cluster_k > 1);
    const bool single = infonce_single_pass(m, n, off, inv_t);
    for (int s = 0; s < 2; ++s) {
        const Operand X = grad_operand(P.gop[single ? 0 : s], P.ldg, grad_nseg(grad_dtype));
        const Operand& Y = s == 0 ? B : A;
        int rc = fill_gemm_store_mn(J2[s], X, m, Y, n, dim);
        if (rc) return rc;
        if (single) {
            // G carries alpha * sx_i * sy_j: dA_i = (1/sx_i) sum_j G_ij b_j, dB_j = (1/sy_j) sum_i G_ij a_i — the second reads
            // the SAME G as an MN-major A operand (contraction over its rows)
            J2[s].sx = s == 0 ? P.ra : P.rb;
            J2[s].sx_recip = 1;
            if (s == 1) {
                if ((rc = make_operand_map(&J2[s].tmx, X.base, n, m, X.nseg, X.row_stride, X.seg_stride, 64))) return rc;
                J2[s].x_mn_major = 1;
            }
        }
        J2[s].ksplit = ksplit;
        if (*fused) {
            J2[s].fin_dx = s == 0 ? d_a : d_b;
            J2[s].fin_dx_dtype = grad_dtype;
            J2[s].fin_ld_dx = ld_grad;
            J2[s].fin_x = static_cast<const char*>(s == 0 ? a_all : b_all) + off * ld * esz;
            J2[s].fin_x_dtype = dtype;
            J2[s].fin_ldx = ld;
            J2[s].fin_sx = (s == 0 ? P.ra : P.rb) + off;
        } else {
            J2[s].out = P.g[s];
            J2[s].ld_out = dim;
            J2[s].slice_stride = ksplit > 1 ? m * dim : 0;
        }
    }
    return STIL_OK;
}

void infonce_gradfinish_jobs(GradFinishJob* G2, const InfoncePlan& P, const void* a_all, const void* b_all, int dtype,
                             int64_t m, int64_t n_cols, int64_t dim, int64_t ld, int64_t off, void* d_a, void* d_b, int grad_dtype,
                             int64_t ld_grad) {
    const int esz = dtype == STIL_BF16 ? 2 : 4;
    for (int s = 0; s < 2; ++s) {
        GradFinishJob& j = G2[s];
        std::memset(&j, 0, sizeof(j));
        j.g = P.g[s];
        j.x = static_cast<const char*>(s == 0 ? a_all : b_all) + off * ld * esz;
        j.x_dtype = dtype; j.ldx = ld;
        j.sx = (s == 0 ? P.ra : P.rb) + off;
        j.dx = s == 0 ? d_a : d_b;
        j.dx_dtype = grad_dtype; j.ld_dx = ld_grad;
        j.rows = (int)m; j.dim = (int)dim;
        j.row_begin = (int)(s * m);
        j.nslices = infonce_dx_ksplit(m, n_cols, dim);
        j.slice_stride = m * dim;
    }
}

// side streams for the independent branches of the step (fork/join with events; capturable)
constexpr int kSide = 4;   // 0: losses, 1: prototype sums, 2: masked CE, 3: pseudo-label + prototype-CE backward chain
struct SideStreams {
    cudaStream_t s[kSide];
    cudaEvent_t fork, fork2, nce_stats, join[kSide];
    bool ready;
};
SideStreams g_side[64];
std::mutex g_side_mutex;

int get_side_streams(SideStreams** out) {
    int dev = 0;
    STIL_CUDA(cudaGetDevice(&dev));
    STIL_REQUIRE(dev >= 0 && dev < 64, STIL_E_ARG, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_side_mutex);
    SideStreams& S = g_side[dev];
    if (!S.ready) {
        for (int i = 0; i < kSide; ++i) {
            STIL_CUDA(cudaStreamCreateWithFlags(&S.s[i], cudaStreamNonBlocking));
            STIL_CUDA(cudaEventCreateWithFlags(&S.join[i], cudaEventDisableTiming));
        }
        STIL_CUDA(cudaEventCreateWithFlags(&S.fork, cudaEventDisableTiming));
        STIL_CUDA(cudaEventCreateWithFlags(&S.fork2, cudaEventDisableTiming));
        STIL_CUDA(cudaEventCreateWithFlags(&S.nce_stats, cudaEventDisableTiming));
        S.ready = true;
    }
    *out = &S;
    return STIL_OK;
}

}  // namespace
bool g_pdl_enabled = true;
bool pdl_enabled() { return g_pdl_enabled; }
}  // namespace stil

using namespace stil;

extern "C" {

STIL_API int stil_version(void) { return STIL_VERSION; }
STIL_API int64_t stil_abi_struct_bytes(int which) {
    return which == 0 ? (int64_t)sizeof(stil_head_step_args) : which == 1 ? (int64_t)sizeof(stil_p2p_channel)
           : which == 2 ? (int64_t)sizeof(stil_ema_entry) : -1;
}
STIL_API const char* stil_last_error(void) { return stil::last_error(); }

STIL_API int stil_debug_trace(void* buffer) { return gemm_set_trace(buffer); }

STIL_API int stil_debug_pdl(int enable) {
    stil::g_pdl_enabled = enable != 0;
    return STIL_OK;
}

STIL_API int stil_check_device(void) {
    int dev = 0;
    STIL_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    STIL_CUDA(cudaGetDeviceProperties(&p, dev));
    STIL_REQUIRE(p.major == 10, STIL_E_ARCH, "device %d is sm_%d%d; libstil_head is built for sm_100a only", dev, p.major,
                 p.minor);
    return STIL_OK;
}

// =============================================================================================== a1
STIL_API int64_t stil_infonce_workspace_bytes(int64_t m, int64_t n, int64_t dim, int dtype) {
    return plan_infonce(nullptr, 0, m, n, dim, dtype, true).bytes;
}

STIL_API int stil_infonce_fwd(const void* a_loc, const void* b_loc, const void* a_all, const void* b_all, int dtype,
                     int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset, float temperature,
                     float lambda0, float* loss_sum, float* lse_row, float* lse_col, float* logits,
                     int64_t ld_logits, void* workspace, int64_t workspace_bytes, void* stream) {
    (void)a_loc; (void)b_loc;
    int rc = infonce_args_check(a_all, b_all, dtype, m, n, dim, ld, row_offset, temperature, lambda0);
    if (rc) return rc;
    STIL_REQUIRE(loss_sum && lse_row && lse_col, STIL_E_ARG, "infonce_fwd: null output");
    InfoncePlan P = plan_infonce(workspace, workspace_bytes, m, n, dim, dtype, false);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "infonce workspace too small: need %lld bytes",
                 (long long)P.bytes);
    const float inv_t = 1.0f / temperature;
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    PL.zero_words = P.ticket; PL.n_zero = 8;
    prep_add(PL, prep_job(a_all, dtype, n, dim, ld, P.nseg, P.a_op, nullptr, 0, 0, P.ra));
    prep_add(PL, prep_job(b_all, dtype, n, dim, ld, P.nseg, P.b_op, nullptr, 0, 0, P.rb));
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    if (m == 0) {
        STIL_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), S(stream)));
        return STIL_OK;
    }
    const Operand A = rowmajor_operand(a_all, dtype, dim, ld, P.a_op, P.nseg);
    const Operand B = rowmajor_operand(b_all, dtype, dim, ld, P.b_op, P.nseg);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = infonce_stats_jobs(GL.job, P, A, B, m, n, dim, row_offset, inv_t, logits, ld_logits, &GL.njobs))) return rc;
    gemm_job_tiles(GL);
    set_early(GL, dtype == STIL_BF16, dtype == STIL_BF16);   // predecessor = prep: bf16 operands are the caller's inputs
    if ((rc = launch_gemm(GL, S(stream)))) return rc;
    FinishLaunch FL;
    std::memset(&FL, 0, sizeof(FL));
    infonce_finish_jobs(FL.job, P, a_all, b_all, dtype, m, n, dim, ld, row_offset, inv_t, lambda0, lse_row, lse_col, 0);
    FL.job[0].row_begin = 0;
    FL.job[1].row_begin = (int)m;
    FL.njobs = 2;
    FL.total_rows = (int)(2 * m);
    FL.block_partials = P.block_partials;
    FL.ticket = P.ticket;
    FL.out_loss = loss_sum;  // slot 0 only
    return launch_finish(FL, S(stream));
}

}  // extern "C"
namespace {
int infonce_bwd_impl(const void* a_all, const void* b_all, int dtype,
                     int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset, float temperature,
                     float lambda0, const float* lse_row_all, const float* lse_col_all, const float* grad_loss,
                     void* d_a, void* d_b, int grad_dtype, int64_t ld_grad, void* workspace,
                     int64_t workspace_bytes, void* stream, bool after_fwd) {
    int rc = infonce_args_check(a_all, b_all, dtype, m, n, dim, ld, row_offset, temperature, lambda0);
    if (rc) return rc;
    STIL_REQUIRE(lse_row_all && lse_col_all && d_a && d_b, STIL_E_ARG, "infonce_bwd: null pointer");
    STIL_REQUIRE(grad_dtype == STIL_F32 || grad_dtype == STIL_BF16, STIL_E_DTYPE, "bad grad dtype");
    InfoncePlan P = plan_infonce(workspace, workspace_bytes, m, n, dim, dtype, true);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "infonce workspace too small: need %lld bytes",
                 (long long)P.bytes);
    if (m == 0) return STIL_OK;
    const float inv_t = 1.0f / temperature;
    if (!after_fwd) {
        // inverse norms (and the fp32 operand split): still in the workspace when the forward just ran on it
        PrepLaunch PL;
        std::memset(&PL, 0, sizeof(PL));
        prep_add(PL, prep_job(a_all, dtype, n, dim, ld, P.nseg, P.a_op, nullptr, 0, 0, P.ra));
        prep_add(PL, prep_job(b_all, dtype, n, dim, ld, P.nseg, P.b_op, nullptr, 0, 0, P.rb));
        if ((rc = launch_prep(PL, S(stream)))) return rc;
    }
    const Operand A = rowmajor_operand(a_all, dtype, dim, ld, P.a_op, P.nseg);
    const Operand B = rowmajor_operand(b_all, dtype, dim, ld, P.b_op, P.nseg);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = infonce_grad_jobs(GL.job, P, A, B, m, n, dim, row_offset, inv_t, lambda0, lse_row_all, lse_col_all,
                                grad_loss, grad_dtype, &GL.njobs)))
        return rc;
    gemm_job_tiles(GL);
    // predecessor = this call's prep: bf16 operands are the caller's inputs (after_fwd: the predecessor is the caller's)
    if (!after_fwd) set_early(GL, dtype == STIL_BF16, dtype == STIL_BF16);
    GemmLaunch GS, GB;
    std::memset(&GS, 0, sizeof(GS));
    bool fused = false;
    int dx_ck = 1;
    if ((rc = infonce_store_jobs(GS.job, P, A, B, a_all, b_all, dtype, m, n, dim, ld, row_offset, d_a, d_b, grad_dtype,
                                 ld_grad, inv_t, &fused, &dx_ck)))
        return rc;
    GS.njobs = 2;
    GS.cluster_k = dx_ck;
    gemm_job_tiles(GS);
    if (fuse_bwd(GB, GL, GS)) {
        if ((rc = launch_gemm(GB, S(stream)))) return rc;       // recompute + dLogits + dX in one kernel
    } else {
        if ((rc = launch_gemm(GL, S(stream)))) return rc;
        set_early(GS, false, true);   // predecessor = GRAD (writes X = G); Y is an input / an operand of the forward call
        if ((rc = launch_gemm(GS, S(stream)))) return rc;
    }
    if (fused) return STIL_OK;
    GradFinishLaunch GF;
    std::memset(&GF, 0, sizeof(GF));
    infonce_gradfinish_jobs(GF.job, P, a_all, b_all, dtype, m, n, dim, ld, row_offset, d_a, d_b, grad_dtype, ld_grad);
    GF.njobs = 2;
    GF.total_rows = (int)(2 * m);
    return launch_grad_finish(GF, S(stream));
}
}  // namespace
extern "C" {
STIL_API int stil_infonce_bwd(const void* a_loc, const void* b_loc, const void* a_all, const void* b_all, int dtype,
                     int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset, float temperature,
                     float lambda0, const float* lse_row_all, const float* lse_col_all, const float* grad_loss,
                     void* d_a, void* d_b, int grad_dtype, int64_t ld_grad, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    (void)a_loc; (void)b_loc;
    return infonce_bwd_impl(a_all, b_all, dtype, m, n, dim, ld, row_offset, temperature, lambda0, lse_row_all, lse_col_all,
                            grad_loss, d_a, d_b, grad_dtype, ld_grad, workspace, workspace_bytes, stream, false);
}
STIL_API int stil_infonce_bwd_after_fwd(const void* a_all, const void* b_all, int dtype, int64_t m, int64_t n, int64_t dim,
                                        int64_t ld, int64_t row_offset, float temperature, float lambda0,
                                        const float* lse_row_all, const float* lse_col_all, const float* grad_loss,
                                        void* d_a, void* d_b, int grad_dtype, int64_t ld_grad, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
    return infonce_bwd_impl(a_all, b_all, dtype, m, n, dim, ld, row_offset, temperature, lambda0, lse_row_all, lse_col_all,
                            grad_loss, d_a, d_b, grad_dtype, ld_grad, workspace, workspace_bytes, stream, true);
}

// --------------------------------------------------------------------------- global-batch InfoNCE on gathered buffers
// The data-parallel head's fused schedule (csrc/p2p.cu): the embeddings and their inverse norms are pushed into every
// rank's buffer by stil_p2p_push_embeddings, the statistics GEMM consumes them tile by tile as they arrive (arrival
// flags), stil_p2p_push_lse merges and pushes the row LSEs, and the gradient GEMM waits for the LSEs it needs.
}  // extern "C"
namespace stil {
void infonce_stat_partials(void* workspace, int64_t m, int64_t n, int64_t dim, int dtype, const float** pmax,
                           const float** psum, int* slots) {
    InfoncePlan P = plan_infonce(workspace, 1ll << 60, m, n, dim, dtype, true);
    for (int s = 0; s < 2; ++s) { pmax[s] = P.pmax[s]; psum[s] = P.psum[s]; }
    *slots = (int)stat_slots(n);
}
}  // namespace stil
namespace {
void set_wait(GemmLaunch& GL, const void* flags, const void* seq, int64_t rows_per_peer, int wait_y) {
    for (int j = 0; j < GL.njobs; ++j) {
        GL.job[j].wait_flags = static_cast<const unsigned long long*>(flags);
        GL.job[j].wait_seq = static_cast<const unsigned long long*>(seq);
        GL.job[j].wait_rows_per_peer = (int)rows_per_peer;
        GL.job[j].wait_y = wait_y;
    }
}
int gathered_check(const void* a_all, const void* b_all, const float* ra_all, const float* rb_all, int dtype, int64_t m,
                   int64_t n, int64_t dim, int64_t ld, int64_t row_offset, float temperature, float lambda0,
                   const void* wait_flags, const void* wait_seq, int64_t rows_per_peer) {
    STIL_REQUIRE(dtype == STIL_BF16, STIL_E_DTYPE, "gathered InfoNCE: bf16 embeddings only (fp32 uses stil_infonce_fwd/bwd)");
    int rc = infonce_args_check(a_all, b_all, dtype, m, n, dim, ld, row_offset, temperature, lambda0);
    if (rc) return rc;
    STIL_REQUIRE(ra_all && rb_all && ((wait_flags == nullptr) == (wait_seq == nullptr)) &&
                     (wait_flags == nullptr || rows_per_peer >= 1),
                 STIL_E_ARG, "gathered InfoNCE: bad norm / flag arguments");
    return STIL_OK;
}
}  // namespace
extern "C" {
STIL_API int stil_infonce_stats_gathered(const void* a_all, const void* b_all, const float* ra_all, const float* rb_all,
                                         int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset,
                                         float temperature, const void* wait_flags, const void* wait_seq,
                                         int64_t rows_per_peer, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = gathered_check(a_all, b_all, ra_all, rb_all, dtype, m, n, dim, ld, row_offset, temperature, 0.5f, wait_flags,
                            wait_seq, rows_per_peer);
    if (rc) return rc;
    InfoncePlan P = plan_infonce(workspace, workspace_bytes, m, n, dim, dtype, false);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "infonce workspace too small: need %lld bytes",
                 (long long)P.bytes);
    if (m == 0) return STIL_OK;
    P.ra = const_cast<float*>(ra_all);
    P.rb = const_cast<float*>(rb_all);
    const Operand A = rowmajor_operand(a_all, dtype, dim, ld, nullptr, 1);
    const Operand B = rowmajor_operand(b_all, dtype, dim, ld, nullptr, 1);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = infonce_stats_jobs(GL.job, P, A, B, m, n, dim, row_offset, 1.0f / temperature, nullptr, 0, &GL.njobs))) return rc;
    gemm_job_tiles(GL);
    set_wait(GL, wait_flags, wait_seq, rows_per_peer, 1);
    return launch_gemm(GL, S(stream));
}

/* loss partial of the local rows (and their LSEs) from the statistics left by stil_infonce_stats_gathered — off the
 * critical chain.  The workspace's first 256 bytes must have been zero when it was first used (reduction ticket). */
STIL_API int stil_infonce_loss_gathered(const void* a_all, const void* b_all, const float* ra_all, const float* rb_all,
                                        int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset,
                                        float temperature, float lambda0, float* loss_sum, float* lse_row, float* lse_col,
                                        void* const* bases, int world, int rank, int64_t loss_ll_offset, const void* ll_tag,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = gathered_check(a_all, b_all, ra_all, rb_all, dtype, m, n, dim, ld, row_offset, temperature, lambda0, nullptr,
                            nullptr, 0);
    if (rc) return rc;
    STIL_REQUIRE(loss_sum && lse_row && lse_col, STIL_E_ARG, "infonce_loss_gathered: null output");
    InfoncePlan P = plan_infonce(workspace, workspace_bytes, m, n, dim, dtype, false);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "infonce workspace too small");
    if (m == 0) return STIL_OK;
    P.ra = const_cast<float*>(ra_all);
    P.rb = const_cast<float*>(rb_all);
    FinishLaunch FL;
    std::memset(&FL, 0, sizeof(FL));
    infonce_finish_jobs(FL.job, P, a_all, b_all, dtype, m, n, dim, ld, row_offset, 1.0f / temperature, lambda0, lse_row,
                        lse_col, 0);
    FL.job[0].row_begin = 0;
    FL.job[1].row_begin = (int)m;
    FL.njobs = 2;
    FL.total_rows = (int)(2 * m);
    FL.block_partials = P.block_partials;
    FL.ticket = P.ticket;
    FL.out_loss = loss_sum;
    if (bases) {
        // the partial also goes to every rank's LL word [rank] at loss_ll_offset (summed there in rank order)
        STIL_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && ll_tag && loss_ll_offset % 8 == 0, STIL_E_ARG,
                     "infonce_loss_gathered: bad peer arguments");
        for (int p = 0; p < world; ++p)
            FL.ll_out[p] = reinterpret_cast<unsigned long long*>(static_cast<char*>(bases[p]) + loss_ll_offset) + rank;
        FL.ll_world = world;
        FL.ll_tag = static_cast<const unsigned long long*>(ll_tag);
    }
    return launch_finish(FL, S(stream));
}

STIL_API int stil_infonce_bwd_gathered(const void* a_all, const void* b_all, const float* ra_all, const float* rb_all,
                                       int dtype, int64_t m, int64_t n, int64_t dim, int64_t ld, int64_t row_offset,
                                       float temperature, float lambda0, const void* lse_row_ll,
                                       const void* lse_col_ll, const void* ll_tag, const float* grad_loss, void* d_a,
                                       void* d_b, int grad_dtype,
                                       int64_t ld_grad, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = gathered_check(a_all, b_all, ra_all, rb_all, dtype, m, n, dim, ld, row_offset, temperature, lambda0, nullptr,
                            nullptr, 0);
    if (rc) return rc;
    const float* lse_row_all = static_cast<const float*>(lse_row_ll);
    const float* lse_col_all = static_cast<const float*>(lse_col_ll);
    STIL_REQUIRE(lse_row_all && lse_col_all && ll_tag && d_a && d_b, STIL_E_ARG, "infonce_bwd_gathered: null pointer");
    STIL_REQUIRE(grad_dtype == STIL_F32 || grad_dtype == STIL_BF16, STIL_E_DTYPE, "bad grad dtype");
    InfoncePlan P = plan_infonce(workspace, workspace_bytes, m, n, dim, dtype, true);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "infonce workspace too small");
    if (m == 0) return STIL_OK;
    P.ra = const_cast<float*>(ra_all);
    P.rb = const_cast<float*>(rb_all);
    const Operand A = rowmajor_operand(a_all, dtype, dim, ld, nullptr, 1);
    const Operand B = rowmajor_operand(b_all, dtype, dim, ld, nullptr, 1);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = infonce_grad_jobs(GL.job, P, A, B, m, n, dim, row_offset, 1.0f / temperature, lambda0, lse_row_all, lse_col_all,
                                grad_loss, grad_dtype, &GL.njobs)))
        return rc;
    gemm_job_tiles(GL);
    // every embedding row landed before the statistics GEMM finished (it waited for each peer): the operands stream
    // before the wait; only the epilogue needs the peers' LSEs
    set_early(GL, true, true);
    for (int s = 0; s < 2; ++s) {
        // LL-word vectors: entry i is 8 bytes; the job builder offset the row pointer in floats, redo it in words
        GL.job[s].lse_x = (s == 0 ? lse_row_all : lse_col_all) + 2 * row_offset;
        GL.job[s].lse_ll_tag = static_cast<const unsigned long long*>(ll_tag);
    }
    GemmLaunch GS, GB;
    std::memset(&GS, 0, sizeof(GS));
    bool fused = false;
    int dx_ck = 1;
    if ((rc = infonce_store_jobs(GS.job, P, A, B, a_all, b_all, dtype, m, n, dim, ld, row_offset, d_a, d_b, grad_dtype,
                                 ld_grad, 1.0f / temperature, &fused, &dx_ck)))
        return rc;
    GS.njobs = 2;
    GS.cluster_k = dx_ck;
    gemm_job_tiles(GS);
    if (fuse_bwd(GB, GL, GS)) {
        if ((rc = launch_gemm(GB, S(stream)))) return rc;
    } else {
        if ((rc = launch_gemm(GL, S(stream)))) return rc;
        set_early(GS, false, true);
        if ((rc = launch_gemm(GS, S(stream)))) return rc;
    }
    if (fused) return STIL_OK;
    GradFinishLaunch GF;
    std::memset(&GF, 0, sizeof(GF));
    infonce_gradfinish_jobs(GF.job, P, a_all, b_all, dtype, m, n, dim, ld, row_offset, d_a, d_b, grad_dtype, ld_grad);
    GF.njobs = 2;
    GF.total_rows = (int)(2 * m);
    return launch_grad_finish(GF, S(stream));
}

// =============================================================================================== a3 logits
STIL_API int64_t stil_proto_logits_workspace_bytes(int64_t rows, int64_t k, int64_t dim, int dtype) {
    return plan_proto(nullptr, 0, rows, k, dim, dtype).bytes;
}

STIL_API int stil_proto_logits(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* prototypes,
                      int64_t k, float* out, int64_t ld_out, void* workspace, int64_t workspace_bytes,
                      void* stream) {
    int rc = check_embed(feat, dtype, rows, dim, ld, "proto_logits feat");
    if (rc) return rc;
    if ((rc = check_embed(prototypes, STIL_F32, k, dim, dim, "prototypes"))) return rc;
    STIL_REQUIRE(out && ld_out >= k, STIL_E_ARG, "proto_logits: bad output");
    ProtoPlan P = plan_proto(workspace, workspace_bytes, rows, k, dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "proto workspace too small: need %lld bytes",
                 (long long)P.bytes);
    if (rows == 0 || k == 0) return STIL_OK;
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    if (dtype != STIL_BF16) prep_add(PL, prep_job(feat, dtype, rows, dim, ld, P.feat_nseg, P.feat_op, nullptr, 0, 0, nullptr));
    prep_add(PL, prep_job(prototypes, STIL_F32, k, dim, dim, P.proto_nseg, P.proto_op, nullptr, 0, 0, nullptr));
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    const Operand X = rowmajor_operand(feat, dtype, dim, ld, P.feat_op, P.feat_nseg);
    const Operand Y = rowmajor_operand(nullptr, STIL_F32, dim, dim, P.proto_op, P.proto_nseg);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = fill_gemm_common(GL.job[0], X, 0, rows, Y, k, dim))) return rc;
    GL.job[0].mode = GEMM_STORE;
    GL.job[0].out = out;
    GL.job[0].ld_out = ld_out;
    GL.njobs = 1;
    gemm_job_tiles(GL);
    return launch_gemm(GL, S(stream));
}

// =============================================================================================== a2+a3
STIL_API int stil_cgpl_pgls(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                   const float* teacher_logits, int64_t ld_t, int64_t rows, int64_t k, float temperature,
                   float rate_pseudo, float th1, int past_start_epoch, const float* prediction_in, int64_t ld_pin,
                   float* pseudo_label, int64_t ld_pl,
                   float* prediction, int64_t ld_pred, float* max_prob, int64_t* max_idx, uint8_t* mask1,
                   uint8_t* case1, uint8_t* case2_i, uint8_t* case2_t, uint8_t* case3, int64_t* top1,
                   int32_t* cls, uint8_t* conf, void* stream) {
    STIL_REQUIRE(logit_dtype == STIL_F32 || logit_dtype == STIL_BF16, STIL_E_DTYPE, "cgpl_pgls: bad logit dtype");
    STIL_REQUIRE(!prediction_in || ld_pin >= k, STIL_E_SHAPE, "cgpl_pgls: prediction_in leading dimension smaller than k");
    STIL_REQUIRE(rows == 0 || (y_m && y_i && y_t && teacher_logits && pseudo_label && max_idx && mask1), STIL_E_ARG,
                 "cgpl_pgls: null pointer");
    STIL_REQUIRE(ld_y >= k && ld_t >= k && ld_pl >= k && (!prediction || ld_pred >= k), STIL_E_SHAPE,
                 "cgpl_pgls: leading dimension smaller than k");
    return launch_cgpl_pgls(y_m, y_i, y_t, logit_dtype, ld_y, teacher_logits, ld_t, rows, k, temperature, rate_pseudo,
                            th1, past_start_epoch, prediction_in, ld_pin, pseudo_label, ld_pl, prediction, ld_pred, max_prob,
                            max_idx, mask1, case1, case2_i, case2_t, case3, top1, cls, conf, nullptr, 0, nullptr, nullptr,
                            S(stream));
}

STIL_API int stil_label_argmax(const float* label, int64_t ld, int64_t rows, int64_t k, float threshold, int32_t* cls,
                      uint8_t* conf, float* max_prob, void* stream) {
    STIL_REQUIRE(rows == 0 || (label && cls && conf), STIL_E_ARG, "label_argmax: null pointer");
    STIL_REQUIRE(ld >= k, STIL_E_SHAPE, "label_argmax: ld < k");
    return launch_label_argmax(label, ld, rows, k, threshold, cls, conf, max_prob, S(stream));
}

// =============================================================================================== a4
STIL_API int64_t stil_proto_ce_workspace_bytes(int64_t rows, int64_t k, int64_t dim, int dtype) {
    return plan_proto(nullptr, 0, rows, k, dim, dtype).bytes;
}

namespace {
int proto_stats_job(GemmJob& J, const ProtoPlan& P, const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld,
                    int64_t k, float inv_t) {
    const Operand X = rowmajor_operand(feat, dtype, dim, ld, P.feat_op, P.feat_nseg);
    const Operand Y = rowmajor_operand(nullptr, STIL_F32, dim, dim, P.proto_op, P.proto_nseg);
    int rc = fill_gemm_common(J, X, 0, rows, Y, k, dim);
    if (rc) return rc;
    J.mode = GEMM_STATS;
    J.alpha = inv_t;
    J.part_max = P.pmax;
    J.part_sum = P.psum;
    return STIL_OK;
}
void proto_finish_job(FinishJob& F, const ProtoPlan& P, const void* feat, int dtype, int64_t rows, int64_t dim,
                      int64_t ld, const float* prototypes, int64_t k, const int32_t* cls, const uint8_t* conf,
                      float inv_t, float* lse, float* w, int loss_slot) {
    std::memset(&F, 0, sizeof(F));
    F.kind = 1;
    F.M = (int)rows;
    F.tiles_n = (int)stat_slots(k);
    F.part_max = P.pmax; F.part_sum = P.psum;
    F.lse = lse;
    F.x = feat; F.x_dtype = dtype; F.ldx = ld;
    F.y = prototypes; F.y_dtype = STIL_F32; F.ldy = dim;
    F.dim = (int)dim;
    F.alpha = inv_t;
    F.coef = 1.f / (float)rows;   // .mean() over ALL rows (utils/prototype_loss.py:39)
    F.cls = cls; F.conf = conf; F.w = w;
    F.loss_slot = loss_slot;
}
int proto_grad_job(GemmJob& J, const ProtoPlan& P, const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld,
                   int64_t k, float inv_t, const int32_t* cls, const float* lse, const float* w,
                   const float* grad_loss, const float* z, int64_t ldz, const uint8_t* conf, int grad_dtype) {
    const Operand X = rowmajor_operand(feat, dtype, dim, ld, P.feat_op, P.feat_nseg);
    const Operand Y = rowmajor_operand(nullptr, STIL_F32, dim, dim, P.proto_op, P.proto_nseg);
    int rc = fill_gemm_common(J, X, 0, rows, Y, k, dim);
    if (rc) return rc;
    J.mode = GEMM_GRAD;
    J.alpha = inv_t;
    J.tgt_vec = cls;
    J.gscale = grad_loss;
    J.gop = P.gop;
    J.ld_g = P.ldg;
    J.g_nseg = grad_nseg(grad_dtype);
    if (lse) {
        J.lse_x = lse;
        J.u_vec = w;
    } else {
        // fused: LSE from the GEMM_STATS partials, coefficient from the picked logit, both in the kernel
        J.px_max = P.pmax; J.px_sum = P.psum;
        J.px_tiles = (int)stat_slots(k);
        J.w_z = z; J.w_ldz = ldz;
        J.w_conf = conf;
        J.w_coef = 1.f / (float)rows;
    }
    return STIL_OK;
}
// the STORE half of a fused, column-split backward: partial tiles into the slices of P.g
int proto_store_job_split(GemmJob& J, const ProtoPlan& P, int64_t rows, int64_t dim, int64_t k, int grad_dtype) {
    const Operand X = grad_operand(P.gop, P.ldg, grad_nseg(grad_dtype));
    const Operand Y = rowmajor_operand(nullptr, STIL_F32, dim, dim, P.proto_op, P.proto_nseg);
    int rc = fill_gemm_store_mn(J, X, rows, Y, k, dim);
    if (rc) return rc;
    J.out = P.g;
    J.ld_out = dim;
    J.ksplit = (int)proto_bwd_slices(k);
    J.slice_stride = rows * dim;
    return STIL_OK;
}
// fused backward of the prototype CE when the shapes allow it (true: launched; d_feat still needs launch_proto_gradfinish)
int proto_bwd_fused(GemmLaunch& GG, const ProtoPlan& P, int64_t rows, int64_t dim, int64_t k, int grad_dtype,
                    cudaStream_t stream, bool* done) {
    *done = false;
    if (proto_bwd_slices(k) < 2) return STIL_OK;
    GemmLaunch GS, GB;
    std::memset(&GS, 0, sizeof(GS));
    int rc = proto_store_job_split(GS.job[0], P, rows, dim, k, grad_dtype);
    if (rc) return rc;
    GS.njobs = 1;
    gemm_job_tiles(GS);
    if (!fuse_bwd(GB, GG, GS)) return STIL_OK;
    if ((rc = launch_gemm(GB, stream))) return rc;
    *done = true;
    return STIL_OK;
}
int launch_proto_gradfinish(const ProtoPlan& P, int64_t rows, int64_t dim, int64_t k, void* d_feat, int grad_dtype,
                            int64_t ld_grad, bool sliced, cudaStream_t stream) {
    GradFinishLaunch GF;
    std::memset(&GF, 0, sizeof(GF));
    GradFinishJob& j = GF.job[0];
    j.g = P.g; j.dx = d_feat; j.dx_dtype = grad_dtype; j.ld_dx = ld_grad;
    j.rows = (int)rows; j.dim = (int)dim;
    if (sliced) { j.nslices = (int)proto_bwd_slices(k); j.slice_stride = rows * dim; }
    GF.njobs = 1;
    GF.total_rows = (int)rows;
    return launch_grad_finish(GF, stream);
}
// d_feat = G · prototypes (prototypes read in place, MN-major); cast to grad_dtype in the epilogue when possible
int proto_store_job(GemmJob& J, const ProtoPlan& P, int64_t rows, int64_t dim, int64_t k, void* d_feat, int grad_dtype,
                    int64_t ld_grad, bool* fused) {
    const Operand X = grad_operand(P.gop, P.ldg, grad_nseg(grad_dtype));
    const Operand Y = rowmajor_operand(nullptr, STIL_F32, dim, dim, P.proto_op, P.proto_nseg);
    int rc = fill_gemm_store_mn(J, X, rows, Y, k, dim);
    if (rc) return rc;
    *fused = dim <= kTileN;
    if (*fused) {
        // the contraction over the classes is split over a cluster (see dx_cluster_k); the caller copies J.ksplit into
        // GemmLaunch::cluster_k
        const int ck = dx_cluster_k(ceil_div(rows, kTileM), ceil_div(k, kTileK) * J.npair);
        J.ksplit = (int)std::min<int64_t>(ck, ceil_div(k, kTileK));
        J.fin_dx = d_feat;
        J.fin_dx_dtype = grad_dtype;
        J.fin_ld_dx = ld_grad;
    } else {
        J.out = P.g;
        J.ld_out = dim;
    }
    return STIL_OK;
}
}  // namespace

STIL_API int stil_proto_ce_fwd(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* prototypes,
                      int64_t k, const int32_t* cls, const uint8_t* conf, float temperature, float* loss,
                      float* lse, float* w, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_embed(feat, dtype, rows, dim, ld, "proto_ce feat");
    if (rc) return rc;
    if ((rc = check_embed(prototypes, STIL_F32, k, dim, dim, "prototypes"))) return rc;
    STIL_REQUIRE(loss && (rows == 0 || (cls && conf && lse && w)), STIL_E_ARG, "proto_ce_fwd: null pointer");
    STIL_REQUIRE(temperature > 0.f && k >= 1, STIL_E_ARG, "proto_ce_fwd: bad temperature / k");
    ProtoPlan P = plan_proto(workspace, workspace_bytes, rows, k, dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "proto workspace too small: need %lld bytes",
                 (long long)P.bytes);
    if (rows == 0) {
        STIL_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), S(stream)));
        return STIL_OK;
    }
    const float inv_t = 1.0f / temperature;
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    PL.zero_words = P.ticket; PL.n_zero = 8;
    if (dtype != STIL_BF16) prep_add(PL, prep_job(feat, dtype, rows, dim, ld, P.feat_nseg, P.feat_op, nullptr, 0, 0, nullptr));
    prep_add(PL, prep_job(prototypes, STIL_F32, k, dim, dim, P.proto_nseg, P.proto_op, nullptr, 0, 0, nullptr));
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = proto_stats_job(GL.job[0], P, feat, dtype, rows, dim, ld, k, inv_t))) return rc;
    GL.njobs = 1;
    gemm_job_tiles(GL);
    if ((rc = launch_gemm(GL, S(stream)))) return rc;
    FinishLaunch FL;
    std::memset(&FL, 0, sizeof(FL));
    proto_finish_job(FL.job[0], P, feat, dtype, rows, dim, ld, prototypes, k, cls, conf, inv_t, lse, w, 0);
    FL.njobs = 1;
    FL.total_rows = (int)rows;
    FL.block_partials = P.block_partials;
    FL.ticket = P.ticket;
    FL.out_loss = loss;
    return launch_finish(FL, S(stream));
}

STIL_API int stil_proto_ce_bwd(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* prototypes,
                      int64_t k, const int32_t* cls, const float* lse, const float* w, float temperature,
                      const float* grad_loss, void* d_feat, int grad_dtype, int64_t ld_grad, void* workspace,
                      int64_t workspace_bytes, void* stream) {
    int rc = check_embed(feat, dtype, rows, dim, ld, "proto_ce feat");
    if (rc) return rc;
    if ((rc = check_embed(prototypes, STIL_F32, k, dim, dim, "prototypes"))) return rc;
    STIL_REQUIRE(rows == 0 || (cls && lse && w && d_feat), STIL_E_ARG, "proto_ce_bwd: null pointer");
    STIL_REQUIRE(grad_dtype == STIL_F32 || grad_dtype == STIL_BF16, STIL_E_DTYPE, "bad grad dtype");
    ProtoPlan P = plan_proto(workspace, workspace_bytes, rows, k, dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "proto workspace too small: need %lld bytes",
                 (long long)P.bytes);
    if (rows == 0) return STIL_OK;
    const float inv_t = 1.0f / temperature;
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    if (dtype != STIL_BF16) prep_add(PL, prep_job(feat, dtype, rows, dim, ld, P.feat_nseg, P.feat_op, nullptr, 0, 0, nullptr));
    prep_add(PL, prep_job(prototypes, STIL_F32, k, dim, dim, P.proto_nseg, P.proto_op, nullptr, 0, 0, nullptr));
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = proto_grad_job(GL.job[0], P, feat, dtype, rows, dim, ld, k, inv_t, cls, lse, w, grad_loss, nullptr, 0,
                             nullptr, grad_dtype)))
        return rc;
    GL.njobs = 1;
    gemm_job_tiles(GL);
    bool done = false;
    if ((rc = proto_bwd_fused(GL, P, rows, dim, k, grad_dtype, S(stream), &done))) return rc;
    if (done) return launch_proto_gradfinish(P, rows, dim, k, d_feat, grad_dtype, ld_grad, true, S(stream));
    GemmLaunch GS;
    std::memset(&GS, 0, sizeof(GS));
    bool fused = false;
    if ((rc = proto_store_job(GS.job[0], P, rows, dim, k, d_feat, grad_dtype, ld_grad, &fused))) return rc;
    GS.njobs = 1;
    GS.cluster_k = fused ? GS.job[0].ksplit : 1;
    gemm_job_tiles(GS);
    if ((rc = launch_gemm(GL, S(stream)))) return rc;
    if ((rc = launch_gemm(GS, S(stream)))) return rc;
    if (!fused) {
        GradFinishLaunch GF;
        std::memset(&GF, 0, sizeof(GF));
        GradFinishJob& j = GF.job[0];
        j.g = P.g; j.dx = d_feat; j.dx_dtype = grad_dtype; j.ld_dx = ld_grad;
        j.rows = (int)rows; j.dim = (int)dim;
        GF.njobs = 1;
        GF.total_rows = (int)rows;
        return launch_grad_finish(GF, S(stream));
    }
    return STIL_OK;
}

// =============================================================================================== a5
STIL_API int stil_proto_accumulate(const void* feat, int dtype, int64_t rows, int64_t dim, int64_t ld, const int32_t* cls,
                          const uint8_t* conf, int64_t b_l, float repeat_ratio, int64_t k, float* class_sum,
                          float* class_count, float* psum, float* pcount, void* stream) {
    STIL_REQUIRE(dtype == STIL_F32 || dtype == STIL_BF16, STIL_E_DTYPE, "proto_accumulate: bad dtype");
    STIL_REQUIRE(class_sum && class_count && (rows == 0 || (feat && cls && conf)), STIL_E_ARG, "proto_accumulate: null pointer");
    STIL_REQUIRE(b_l >= 0 && b_l <= rows && repeat_ratio > 0.f, STIL_E_ARG, "proto_accumulate: bad b_l / repeat_ratio");
    STIL_REQUIRE((psum == nullptr) == (pcount == nullptr), STIL_E_ARG, "proto_accumulate: psum/pcount must come together");
    return launch_proto_accumulate(feat, dtype, rows, dim, ld, cls, conf, b_l, repeat_ratio, k, class_sum, class_count,
                                   psum, pcount, S(stream));
}

STIL_API int stil_proto_add(const float* class_sum, const float* class_count, int64_t k, int64_t dim, float* psum,
                   float* pcount, void* stream) {
    STIL_REQUIRE(class_sum && class_count && psum && pcount, STIL_E_ARG, "proto_add: null pointer");
    return launch_proto_add(class_sum, class_count, k, dim, psum, pcount, S(stream));
}

STIL_API int stil_proto_add_gathered_wait(const float* parts, int64_t world, int64_t slot_floats, int64_t k, int64_t dim,
                                          float* class_sum, float* class_count, float* psum, float* pcount,
                                          const void* wait_flags, const void* wait_target, const void* loss_ll,
                                          const void* loss_tag, float* loss_out, void* stream) {
    STIL_REQUIRE(parts && class_sum && class_count && psum && pcount && world >= 1 && world <= 8 && wait_flags && wait_target &&
                     ((loss_ll == nullptr) == (loss_out == nullptr)) && (loss_ll == nullptr || loss_tag != nullptr),
                 STIL_E_ARG, "proto_add_gathered_wait: bad arguments");
    return launch_proto_add_gathered(parts, world, slot_floats, k, dim, class_sum, class_count, psum, pcount, S(stream),
                                     static_cast<const unsigned long long*>(wait_flags),
                                     static_cast<const unsigned long long*>(wait_target),
                                     static_cast<const unsigned long long*>(loss_ll),
                                     static_cast<const unsigned long long*>(loss_tag), loss_out);
}
STIL_API int stil_proto_add_gathered(const float* parts, int64_t world, int64_t slot_floats, int64_t k, int64_t dim,
                                     float* class_sum, float* class_count, float* psum, float* pcount, void* stream) {
    STIL_REQUIRE(parts && class_sum && class_count && psum && pcount && world >= 1 && slot_floats >= k * dim + k, STIL_E_ARG,
                 "proto_add_gathered: bad arguments");
    return launch_proto_add_gathered(parts, world, slot_floats, k, dim, class_sum, class_count, psum, pcount, S(stream));
}

STIL_API int stil_proto_finalize(float* prototypes, float* psum, float* pcount, int64_t k, int64_t dim,
                        int32_t* empty_classes, void* stream) {
    STIL_REQUIRE(prototypes && psum && pcount && empty_classes, STIL_E_ARG, "proto_finalize: null pointer");
    return launch_proto_finalize(prototypes, psum, pcount, k, dim, empty_classes, S(stream));
}

// =============================================================================================== a6
STIL_API int stil_softmax_rows(const void* logits, int dtype, int64_t ld, int64_t rows, int64_t k, float* out, int64_t ld_out,
                               void* stream) {
    STIL_REQUIRE(dtype == STIL_F32 || dtype == STIL_BF16, STIL_E_DTYPE, "softmax_rows: bad dtype");
    STIL_REQUIRE(rows == 0 || (logits && out && k >= 1 && ld >= k && ld_out >= k), STIL_E_ARG, "softmax_rows: bad arguments");
    return launch_softmax_rows(logits, dtype, ld, rows, k, out, ld_out, S(stream));
}

STIL_API int stil_da_batch_mean(const float* probs, int64_t ld, int64_t rows, int64_t k, float* mean, void* stream) {
    STIL_REQUIRE(probs && mean && rows >= 1 && ld >= k, STIL_E_ARG, "da_batch_mean: bad arguments");
    return launch_da_batch_mean(probs, ld, rows, k, mean, S(stream));
}

STIL_API int stil_da_apply(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* batch_mean, float* da_queue,
                  int64_t da_len, int64_t* da_ptr, float* qmean_scratch, float* out, int64_t ld_out, void* stream) {
    STIL_REQUIRE(probs && batch_mean && da_queue && da_ptr && qmean_scratch && out && da_len >= 1 && ld >= k && ld_out >= k,
                 STIL_E_ARG, "da_apply: bad arguments");
    return launch_da_apply(probs, ld, rows, k, batch_mean, da_queue, da_len, da_ptr, qmean_scratch, out, ld_out, S(stream));
}

// =============================================================================================== a7
namespace {
struct SimPlan {
    int nseg;                        // feature / bank operand segments (bf16: 1)
    __nv_bfloat16 *fk_op, *fq_op, *bank_op;
    float *zt, *zs;                  // [rows, ldz]
    int64_t ldz;
    __nv_bfloat16* gop;              // [rows, g_nseg, ldg]
    int64_t ldg;
    float* g;                        // [rows, dim] fp32 accumulation target of the split-K dX GEMM
    int64_t bytes;
};
SimPlan plan_sim(void* ws, int64_t ws_bytes, int64_t rows, int64_t kb, int64_t dim, int dtype) {
    SimPlan P;
    Workspace W(ws, ws_bytes);
    P.nseg = dtype == STIL_BF16 ? 1 : 3;
    P.ldz = round_up(kb, 4);
    P.ldg = pad32(kb);
    P.fk_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(rows * 3 * dim);
    P.fq_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(rows * 3 * dim);
    P.bank_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(dim * 3 * P.ldg);
    P.zt = W.take<float>(rows * P.ldz);
    P.zs = W.take<float>(rows * P.ldz);
    P.gop = W.take<__nv_bfloat16>(rows * 2 * P.ldg);
    P.g = W.take<float>(rows * dim);
    P.bytes = W.off;
    return P;
}
// the bank in the reference layout [dim, k_bank] (simmatch_model.py:68-69): k_bank contiguous
Operand bank_operand(const void* bank, int dtype, int64_t kb, int64_t ld_bank, const __nv_bfloat16* op, int64_t ldg) {
    Operand O;
    if (dtype == STIL_BF16) {
        O.base = static_cast<const __nv_bfloat16*>(bank);
        O.nseg = 1; O.row_stride = ld_bank; O.seg_stride = pad32(kb);
    } else {
        O.base = op;
        O.nseg = 3; O.row_stride = 3 * ldg; O.seg_stride = ldg;
    }
    return O;
}
}  // namespace

STIL_API int64_t stil_simmatch_workspace_bytes(int64_t rows, int64_t k_bank, int64_t dim, int dtype) {
    return plan_sim(nullptr, 0, rows, k_bank, dim, dtype).bytes;
}

STIL_API int stil_simmatch_fwd(const void* feat_ku, const void* feat_qu, int dtype, int64_t rows, int64_t dim, int64_t ld,
                      const void* bank, int64_t ld_bank, const int64_t* labels, int64_t k_bank,
                      const float* prob_ku_orig, int64_t num_classes, float tt, float st, float c_smooth, float* prob_ku,
                      float* loss_in, int grad_dtype, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_embed(feat_ku, dtype, rows, dim, ld, "simmatch feat_ku");
    if (rc) return rc;
    if ((rc = check_embed(feat_qu, dtype, rows, dim, ld, "simmatch feat_qu"))) return rc;
    const int per16 = dtype == STIL_BF16 ? 8 : 4;
    STIL_REQUIRE(bank && k_bank >= 1 && ld_bank >= k_bank && ld_bank % per16 == 0 &&
                     (reinterpret_cast<uintptr_t>(bank) & 15) == 0,
                 STIL_E_ALIGN, "simmatch bank [dim, k_bank]: leading dimension %lld must be >= k_bank and a multiple of %d, "
                 "base 16-byte aligned", (long long)ld_bank, per16);
    STIL_REQUIRE(labels && prob_ku_orig && prob_ku && loss_in && tt > 0.f && st > 0.f && num_classes >= 1, STIL_E_ARG,
                 "simmatch_fwd: bad arguments");
    SimPlan P = plan_sim(workspace, workspace_bytes, rows, k_bank, dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "simmatch workspace too small: need %lld",
                 (long long)P.bytes);
    if (rows == 0) return STIL_OK;
    if (dtype != STIL_BF16) {
        PrepLaunch PL;
        std::memset(&PL, 0, sizeof(PL));
        prep_add(PL, prep_job(feat_ku, dtype, rows, dim, ld, 3, P.fk_op, nullptr, 0, 0, nullptr));
        prep_add(PL, prep_job(feat_qu, dtype, rows, dim, ld, 3, P.fq_op, nullptr, 0, 0, nullptr));
        if ((rc = launch_prep(PL, S(stream)))) return rc;
        // the bank's segment layout is [dim, 3, ldg]: prep writes [rows=dim, nseg, "dim"=k_bank] with row pitch 3*k_bank,
        // so it needs k_bank == ldg
        STIL_REQUIRE(k_bank % 32 == 0, STIL_E_ALIGN, "simmatch: an fp32 bank needs k_bank %% 32 == 0 (got %lld)", (long long)k_bank);
        std::memset(&PL, 0, sizeof(PL));
        prep_add(PL, prep_job(bank, dtype, dim, k_bank, ld_bank, 3, P.bank_op, nullptr, 0, 0, nullptr));
        if ((rc = launch_prep(PL, S(stream)))) return rc;
    }
    const Operand Bk = bank_operand(bank, dtype, k_bank, ld_bank, P.bank_op, P.ldg);
    // teacher and student logits: X = features (K-major), Y = bank read in place as an MN-major operand
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    for (int s = 0; s < 2; ++s) {
        const Operand X = rowmajor_operand(s == 0 ? feat_ku : feat_qu, dtype, dim, ld, s == 0 ? P.fk_op : P.fq_op, 3);
        if ((rc = fill_gemm_store_mn(GL.job[s], X, rows, Bk, dim, k_bank))) return rc;
        GL.job[s].npair = seg_pairs(X.nseg, Bk.nseg, 2, GL.job[s].xseg, GL.job[s].yseg);
        GL.job[s].out = s == 0 ? P.zt : P.zs;
        GL.job[s].ld_out = P.ldz;
    }
    GL.njobs = 2;
    gemm_job_tiles(GL);
    if (bank_logits_eligible(dtype, rows, dim, k_bank, P.ldz)) {
        // short contraction, long bank axis: the persistent resident-rows kernel (bank_sweep.cu)
        if ((rc = launch_bank_logits(feat_ku, feat_qu, rows, dim, ld, bank, ld_bank, k_bank, P.zt, P.zs, P.ldz, S(stream)))) return rc;
    } else if ((rc = launch_gemm(GL, S(stream)))) return rc;
    return launch_simmatch_rows(P.zt, P.zs, P.ldz, reinterpret_cast<const long long*>(labels), (int)rows, (int)k_bank,
                                prob_ku_orig, (int)num_classes, tt, st, c_smooth, prob_ku, loss_in, P.gop, P.ldg,
                                grad_nseg(grad_dtype), S(stream));
}

STIL_API int stil_simmatch_bwd(const void* feat_qu, int dtype, int64_t rows, int64_t dim, const void* bank, int64_t ld_bank,
                      int64_t k_bank, const float* grad_loss_in, void* d_feat_qu, int grad_dtype, int64_t ld_grad,
                      void* workspace, int64_t workspace_bytes, void* stream) {
    (void)feat_qu;
    // grad_loss_in == NULL: unit upstream gradient, i.e. the per-row Jacobian d loss_in[i] / d feat_qu[i, :] — what the
    // Python drop-in computes inside its forward so that the bank may be overwritten (simmatch_model.py:291 runs
    // _update_bank right after this block, before loss.backward()) without touching the gradient
    STIL_REQUIRE(d_feat_qu && bank, STIL_E_ARG, "simmatch_bwd: null pointer");
    STIL_REQUIRE(grad_dtype == STIL_F32 || grad_dtype == STIL_BF16, STIL_E_DTYPE, "bad grad dtype");
    SimPlan P = plan_sim(workspace, workspace_bytes, rows, k_bank, dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "simmatch workspace too small");
    if (rows == 0) return STIL_OK;
    int rc;
    if (grad_dtype == STIL_F32 && bank_dx_eligible(dtype, rows, dim, k_bank, static_cast<const float*>(d_feat_qu), ld_grad))
        // long bank axis onto a small output: the whole [128 x dim] accumulator lives in tensor memory (bank_sweep.cu)
        return launch_bank_dx(P.gop, P.ldg, grad_nseg(grad_dtype), rows, bank, ld_bank, dim, k_bank, grad_loss_in,
                              static_cast<float*>(d_feat_qu), ld_grad, S(stream));
    // d_feat_q[i,:] = grad_i * sum_j G_ij bank[:, j]; G (bf16) was left in the workspace by the forward.
    // The contraction runs over k_bank: split it across CTAs so the whole chip works on it.
    const Operand X = grad_operand(P.gop, P.ldg, grad_nseg(grad_dtype));
    const Operand Bk = bank_operand(bank, dtype, k_bank, ld_bank, P.bank_op, P.ldg);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    // bank as a K-major Y operand: rows = dim (N), contraction = k_bank contiguous
    if ((rc = fill_gemm_common(GL.job[0], X, 0, rows, Bk, dim, k_bank, 1))) return rc;
    GemmJob& J = GL.job[0];
    J.mode = GEMM_STORE;
    J.sx = grad_loss_in;
    J.out = P.g;
    J.ld_out = dim;
    const int64_t tiles = ceil_div(rows, kTileM) * ceil_div(dim, kTileN);
    const int64_t kblocks = ceil_div(k_bank, kTileK);
    J.ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(kblocks / 8, (2 * 148 + tiles - 1) / tiles));
    if (J.ksplit > 1) STIL_CUDA(cudaMemsetAsync(P.g, 0, rows * dim * sizeof(float), S(stream)));
    GL.njobs = 1;
    gemm_job_tiles(GL);
    if ((rc = launch_gemm(GL, S(stream)))) return rc;
    GradFinishLaunch GF;
    std::memset(&GF, 0, sizeof(GF));
    GradFinishJob& j = GF.job[0];
    j.g = P.g; j.dx = d_feat_qu; j.dx_dtype = grad_dtype; j.ld_dx = ld_grad;
    j.rows = (int)rows; j.dim = (int)dim;
    GF.njobs = 1;
    GF.total_rows = (int)rows;
    return launch_grad_finish(GF, S(stream));
}

// ---- a7, column-sharded: every rank sweeps ITS bank shard for the gathered rows of all ranks with a fixed
// shift, the per-row statistics are summed over ranks by the caller (all-reduce), and the gradient partials reduce-scattered.
namespace {
int sim_shard_checks(const void* bank, int dtype, int64_t k_shard, int64_t ld_bank, const char* what) {
    const int per16 = dtype == STIL_BF16 ? 8 : 4;
    STIL_REQUIRE(dtype == STIL_F32 || dtype == STIL_BF16, STIL_E_DTYPE, "%s: bad dtype", what);
    STIL_REQUIRE(bank && k_shard >= 1 && ld_bank >= k_shard && ld_bank % per16 == 0 && (reinterpret_cast<uintptr_t>(bank) & 15) == 0,
                 STIL_E_ALIGN, "%s: bank shard [dim, k_shard] needs ld %% %d == 0 and a 16-byte aligned base", what, per16);
    return STIL_OK;
}
}  // namespace

STIL_API int stil_simmatch_shard_stats(const void* feat_ku, const void* feat_qu, int dtype, int64_t rows, int64_t dim, int64_t ld,
                                       const void* bank, int64_t ld_bank, const int64_t* labels, int64_t k_shard,
                                       const float* prob_ku_orig, int64_t num_classes, float tt, float st, float* stats,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_embed(feat_ku, dtype, rows, dim, ld, "simmatch_shard feat_ku");
    if (rc) return rc;
    if ((rc = check_embed(feat_qu, dtype, rows, dim, ld, "simmatch_shard feat_qu"))) return rc;
    if ((rc = sim_shard_checks(bank, dtype, k_shard, ld_bank, "simmatch_shard_stats"))) return rc;
    STIL_REQUIRE(labels && prob_ku_orig && stats && tt > 0.f && st > 0.f && num_classes >= 1, STIL_E_ARG,
                 "simmatch_shard_stats: bad arguments");
    SimPlan P = plan_sim(workspace, workspace_bytes, rows, k_shard, dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "simmatch workspace too small: need %lld",
                 (long long)P.bytes);
    if (rows == 0) return STIL_OK;
    if (dtype != STIL_BF16) {
        PrepLaunch PL;
        std::memset(&PL, 0, sizeof(PL));
        prep_add(PL, prep_job(feat_ku, dtype, rows, dim, ld, 3, P.fk_op, nullptr, 0, 0, nullptr));
        prep_add(PL, prep_job(feat_qu, dtype, rows, dim, ld, 3, P.fq_op, nullptr, 0, 0, nullptr));
        if ((rc = launch_prep(PL, S(stream)))) return rc;
        STIL_REQUIRE(k_shard % 32 == 0, STIL_E_ALIGN, "simmatch: an fp32 bank needs k_shard %% 32 == 0 (got %lld)", (long long)k_shard);
        std::memset(&PL, 0, sizeof(PL));
        prep_add(PL, prep_job(bank, dtype, dim, k_shard, ld_bank, 3, P.bank_op, nullptr, 0, 0, nullptr));
        if ((rc = launch_prep(PL, S(stream)))) return rc;
    }
    const Operand Bk = bank_operand(bank, dtype, k_shard, ld_bank, P.bank_op, P.ldg);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    for (int s = 0; s < 2; ++s) {
        const Operand X = rowmajor_operand(s == 0 ? feat_ku : feat_qu, dtype, dim, ld, s == 0 ? P.fk_op : P.fq_op, 3);
        if ((rc = fill_gemm_store_mn(GL.job[s], X, rows, Bk, dim, k_shard))) return rc;
        GL.job[s].npair = seg_pairs(X.nseg, Bk.nseg, 2, GL.job[s].xseg, GL.job[s].yseg);
        GL.job[s].out = s == 0 ? P.zt : P.zs;
        GL.job[s].ld_out = P.ldz;
    }
    GL.njobs = 2;
    gemm_job_tiles(GL);
    if (bank_logits_eligible(dtype, rows, dim, k_shard, P.ldz)) {
        if ((rc = launch_bank_logits(feat_ku, feat_qu, rows, dim, ld, bank, ld_bank, k_shard, P.zt, P.zs, P.ldz, S(stream)))) return rc;
    } else if ((rc = launch_gemm(GL, S(stream)))) return rc;
    // per-chunk partial statistics live in the G buffer of the workspace (not written before _shard_grad)
    STIL_REQUIRE((int64_t)simmatch_shard_chunks(rows, k_shard) * rows * (3 + num_classes) * 4 <= rows * 2 * P.ldg * 2, STIL_E_SHAPE,
                 "simmatch_shard_stats: %lld classes do not fit the chunk scratch", (long long)num_classes);
    return launch_simmatch_shard_stats(P.zt, P.zs, P.ldz, reinterpret_cast<const long long*>(labels), (int)rows, (int)k_shard,
                                       prob_ku_orig, (int)num_classes, tt, st, stats, reinterpret_cast<float*>(P.gop), S(stream));
}

STIL_API int stil_simmatch_shard_finish(const float* stats_total, const float* prob_ku_orig, int64_t rows, int64_t num_classes,
                                        float st, float c_smooth, float* prob_ku, float* loss_in, float* norms, void* stream) {
    STIL_REQUIRE(stats_total && prob_ku_orig && norms && st > 0.f, STIL_E_ARG, "simmatch_shard_finish: bad arguments");
    return launch_simmatch_shard_finish(stats_total, prob_ku_orig, (int)rows, (int)num_classes, st, c_smooth, prob_ku, loss_in, norms,
                                        S(stream));
}

STIL_API int stil_simmatch_shard_grad(int dtype, int64_t rows, int64_t dim, const void* bank, int64_t ld_bank,
                                      const int64_t* labels, int64_t k_shard, const float* prob_ku_orig, int64_t num_classes,
                                      float tt, float st, const float* norms, float* d_feat_partial, int64_t ld_grad,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = sim_shard_checks(bank, dtype, k_shard, ld_bank, "simmatch_shard_grad");
    if (rc) return rc;
    STIL_REQUIRE(labels && prob_ku_orig && norms && d_feat_partial, STIL_E_ARG, "simmatch_shard_grad: null pointer");
    SimPlan P = plan_sim(workspace, workspace_bytes, rows, k_shard, dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "simmatch workspace too small");
    if (rows == 0) return STIL_OK;
    // G (bf16 hi+lo) from the logits the statistics pass left in the workspace, then dX_partial = G · bank_shardᵀ
    // when the persistent dX kernel follows (it ADDS partial tiles) and the output is one contiguous block, the G pass clears it
    const bool fold_zero = ld_grad == dim && (rows * dim) % 4 == 0 && bank_dx_eligible(dtype, rows, dim, k_shard, d_feat_partial, ld_grad);
    if ((rc = launch_simmatch_shard_grad(P.zt, P.zs, P.ldz, reinterpret_cast<const long long*>(labels), (int)rows, (int)k_shard,
                                         prob_ku_orig, (int)num_classes, tt, st, norms, P.gop, P.ldg, 2,
                                         fold_zero ? d_feat_partial : nullptr, rows * dim, S(stream))))
        return rc;
    if (fold_zero)
        return launch_bank_dx(P.gop, P.ldg, 2, rows, bank, ld_bank, dim, k_shard, nullptr, d_feat_partial, ld_grad, S(stream), true);
    return stil_simmatch_bwd(nullptr, dtype, rows, dim, bank, ld_bank, k_shard, nullptr, d_feat_partial, STIL_F32, ld_grad,
                             workspace, workspace_bytes, stream);
}

// =============================================================================== FreeMatch / CoTraining thresholds
STIL_API int64_t stil_threshold_workspace_bytes(int64_t rows, int64_t num_classes) { return threshold_workspace_bytes(rows, num_classes); }

STIL_API int stil_freematch_stats(const float* probs_or_logits, int64_t ld, int64_t rows, int64_t num_classes, int is_logits, float* stats,
                                  float* max_probs, int64_t* max_idx, float* probs_out, int64_t ld_probs, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(probs_or_logits && stats && rows >= 1 && num_classes >= 1 && ld >= num_classes && (!probs_out || ld_probs >= num_classes),
                 STIL_E_ARG, "freematch_stats: bad arguments");
    return launch_freematch_stats(probs_or_logits, ld, rows, num_classes, is_logits, stats, max_probs, max_idx, probs_out, ld_probs,
                                  workspace, workspace_bytes, S(stream));
}

STIL_API int stil_freematch_update_mask(const float* stats_total, int64_t rows, int64_t num_classes, float momentum, float clip_thresh,
                                        float* time_p, float* p_model, float* label_hist, const float* max_probs, const int64_t* max_idx,
                                        float* mask, void* workspace, int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(stats_total && time_p && p_model && label_hist && num_classes >= 1 && (rows == 0 || mask), STIL_E_ARG,
                 "freematch_update_mask: bad arguments");
    return launch_freematch_update_mask(stats_total, rows, num_classes, momentum, clip_thresh, time_p, p_model, label_hist, max_probs,
                                        max_idx, mask, workspace, workspace_bytes, S(stream));
}

STIL_API int stil_threshold_rows(const float* logits, int64_t ld, int64_t rows, int64_t num_classes, float threshold, float* probs,
                                 int64_t ld_probs, float* max_probs, int64_t* max_idx, float* mask, void* stream) {
    STIL_REQUIRE(rows == 0 || (logits && probs && max_probs && mask && num_classes >= 1 && ld >= num_classes && ld_probs >= num_classes),
                 STIL_E_ARG, "threshold_rows: bad arguments");
    return launch_threshold_rows(logits, ld, rows, num_classes, threshold, probs, ld_probs, max_probs, max_idx, mask, S(stream));
}

STIL_API int stil_freematch_entropy_fwd(const float* mask, const float* logits_s, int64_t ld, int64_t rows, int64_t num_classes,
                                        const float* p_model, const float* label_hist, float* loss, float* hist_mean, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(mask && logits_s && p_model && label_hist && loss && hist_mean && rows >= 1 && num_classes >= 1 && ld >= num_classes,
                 STIL_E_ARG, "freematch_entropy_fwd: bad arguments");
    return launch_freematch_entropy_fwd(mask, logits_s, ld, rows, num_classes, p_model, label_hist, loss, hist_mean, workspace,
                                        workspace_bytes, S(stream));
}

STIL_API int stil_freematch_entropy_bwd(int64_t rows, int64_t num_classes, const float* grad_loss, float* d_logits_s, int64_t ld_grad,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(d_logits_s && rows >= 1 && num_classes >= 1 && ld_grad >= num_classes, STIL_E_ARG, "freematch_entropy_bwd: bad arguments");
    return launch_freematch_entropy_bwd(rows, num_classes, grad_loss, d_logits_s, ld_grad, workspace, workspace_bytes, S(stream));
}

// =============================================================================================== f-1
STIL_API int64_t stil_masked_softce_workspace_bytes(int64_t rows) {
    Workspace W(nullptr, 0);
    W.take<unsigned int>(64);
    W.take<float>(3 * masked_softce_blocks(rows, 0) + 8);
    return W.off;
}

STIL_API int stil_masked_softce(const void* y_m, const void* y_i, const void* y_t, int logit_dtype, int64_t ld_y,
                       const float* pseudo_label, int64_t ld_pl, const uint8_t* mask1, const uint8_t* case1,
                       const uint8_t* case2_i, const uint8_t* case2_t, const uint8_t* case3,
                       const uint8_t* mask_random, int64_t rows, int64_t k, float* losses, float* d_y_m,
                       float* d_y_i, float* d_y_t, int64_t ld_g, float grad_scale, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(logit_dtype == STIL_F32 || logit_dtype == STIL_BF16, STIL_E_DTYPE, "masked_softce: bad logit dtype");
    STIL_REQUIRE(losses && (rows == 0 || (y_m && y_i && y_t && pseudo_label && mask1 && case1 && case2_i && case2_t &&
                                          case3 && mask_random)),
                 STIL_E_ARG, "masked_softce: null pointer");
    STIL_REQUIRE(workspace && workspace_bytes >= stil_masked_softce_workspace_bytes(rows), STIL_E_WORKSPACE,
                 "masked_softce workspace too small");
    Workspace W(workspace, workspace_bytes);
    unsigned int* ticket = W.take<unsigned int>(64);
    float* partials = W.take<float>(3 * masked_softce_blocks(rows, 0) + 8);
    STIL_CUDA(cudaMemsetAsync(ticket, 0, 16, S(stream)));
    return launch_masked_softce(y_m, y_i, y_t, logit_dtype, ld_y, pseudo_label, ld_pl, mask1, case1, case2_i, case2_t,
                                case3, mask_random, rows, k, losses, d_y_m, d_y_i, d_y_t, ld_g, grad_scale, partials,
                                ticket, S(stream));
}

// =============================================================================================== a8 / a9
namespace {
inline int64_t pad8(int64_t n) { return round_up(n, 8); }

// a matrix stored [contraction rows, n contiguous] (the reference's queue layout [dim, K_q] / [C, K_q]) presented as
// an MN-major Y operand: bf16 in place, fp32 via its segment form [rows, 3, pad8(n)]
Operand colmajor_operand(const void* x, int dtype, int64_t n, int64_t ld, const __nv_bfloat16* op) {
    Operand O;
    if (dtype == STIL_BF16) {
        O.base = static_cast<const __nv_bfloat16*>(x);
        O.nseg = 1; O.row_stride = ld; O.seg_stride = pad8(n);
    } else {
        O.base = op;
        O.nseg = 3; O.row_stride = 3 * pad8(n); O.seg_stride = pad8(n);
    }
    return O;
}
PrepJob prep_job_padded(const void* x, int dtype, int64_t rows, int64_t dim, int64_t ld, __nv_bfloat16* op, int64_t op_dim) {
    PrepJob j = prep_job(x, dtype, rows, dim, ld, 3, op, nullptr, 0, 0, nullptr);
    j.op_dim = (int)op_dim;
    return j;
}
int check_queue(const void* q, int dtype, int64_t k_q, int64_t ld_q, const char* what) {
    const int per16 = dtype == STIL_BF16 ? 8 : 4;
    STIL_REQUIRE(q && k_q >= 1 && ld_q >= k_q && ld_q % per16 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0,
                 STIL_E_ALIGN, "%s [dim, k_q]: leading dimension %lld must be >= k_q and a multiple of %d, base 16-byte aligned",
                 what, (long long)ld_q, per16);
    return STIL_OK;
}

struct SmoothPlan {
    __nv_bfloat16 *f_op, *q_op, *qp_op;   // feat [rows,3,dim] | queue [dim,3,pad8(k_q)] (fp32 only) | queue_probs [C,3,pad8(k_q)]
    float* z;                              // [rows, ldz]
    int64_t ldz;
    __nv_bfloat16* a_op;                   // [rows, 2, ldg]
    int64_t ldg;
    float* s;                              // [rows, lds] = A · queue_probsᵀ
    int64_t lds;
    int64_t bytes;
};
SmoothPlan plan_smooth(void* ws, int64_t ws_bytes, int64_t rows, int64_t k_q, int64_t dim, int64_t c, int dtype) {
    SmoothPlan P;
    Workspace W(ws, ws_bytes);
    P.ldz = round_up(k_q, 4);
    P.ldg = pad32(k_q);
    P.lds = round_up(c, 4);
    P.f_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(rows * 3 * dim);
    P.q_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(dim * 3 * pad8(k_q));
    P.qp_op = W.take<__nv_bfloat16>(c * 3 * pad8(k_q));
    P.z = W.take<float>(rows * P.ldz);
    P.a_op = W.take<__nv_bfloat16>(rows * 2 * P.ldg);
    P.s = W.take<float>(rows * P.lds);
    P.bytes = W.off;
    return P;
}
}  // namespace

STIL_API int64_t stil_bank_smooth_workspace_bytes(int64_t rows, int64_t k_q, int64_t dim, int64_t num_classes, int dtype) {
    return plan_smooth(nullptr, 0, rows, k_q, dim, num_classes, dtype).bytes;
}

STIL_API int stil_bank_smooth(const float* probs, int64_t ld_p, int64_t rows, int64_t num_classes, const void* feat, int dtype,
                              int64_t dim, int64_t ld_f, const void* queue_feat, int64_t ld_q, const float* queue_probs,
                              int64_t ld_qp, int64_t k_q, float temperature, float c_keep, float c_bank, float* out,
                              int64_t ld_out, float th, float* max_prob, int64_t* max_idx, uint8_t* mask, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(probs && num_classes >= 1 && ld_p >= num_classes && (out == nullptr || ld_out >= num_classes), STIL_E_ARG,
                 "bank_smooth: bad probs / out arguments");
    if (queue_feat == nullptr)   // before the bank is used (MMatch.py:221 `current_epoch > 0`, comatch_model.py:288)
        return launch_smooth_mix(probs, ld_p, nullptr, 0, rows, num_classes, 1.f, 0.f, out, ld_out, th, max_prob, max_idx,
                                 mask, S(stream));
    int rc = check_embed(feat, dtype, rows, dim, ld_f, "bank_smooth feat");
    if (rc) return rc;
    if ((rc = check_queue(queue_feat, dtype, k_q, ld_q, "bank_smooth queue_feat"))) return rc;
    STIL_REQUIRE(queue_probs && ld_qp >= k_q && temperature > 0.f, STIL_E_ARG, "bank_smooth: bad queue_probs / temperature");
    SmoothPlan P = plan_smooth(workspace, workspace_bytes, rows, k_q, dim, num_classes, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "bank_smooth workspace too small: need %lld",
                 (long long)P.bytes);
    if (rows == 0) return STIL_OK;
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    if (dtype != STIL_BF16) {
        prep_add(PL, prep_job(feat, dtype, rows, dim, ld_f, 3, P.f_op, nullptr, 0, 0, nullptr));
        prep_add(PL, prep_job_padded(queue_feat, dtype, dim, k_q, ld_q, P.q_op, pad8(k_q)));
    }
    prep_add(PL, prep_job_padded(queue_probs, STIL_F32, num_classes, k_q, ld_qp, P.qp_op, pad8(k_q)));
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    // z = feat · queue_feat (queue read in place in the reference layout)
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    const Operand X = rowmajor_operand(feat, dtype, dim, ld_f, P.f_op, 3);
    const Operand Qf = colmajor_operand(queue_feat, dtype, k_q, ld_q, P.q_op);
    if ((rc = fill_gemm_store_mn(GL.job[0], X, rows, Qf, dim, k_q))) return rc;
    GL.job[0].npair = seg_pairs(X.nseg, Qf.nseg, 2, GL.job[0].xseg, GL.job[0].yseg);
    GL.job[0].out = P.z;
    GL.job[0].ld_out = P.ldz;
    GL.njobs = 1;
    gemm_job_tiles(GL);
    if ((rc = launch_gemm(GL, S(stream)))) return rc;
    // A = rownorm(exp(z / T)) as a bf16 hi/lo operand, then s = A · queue_probsᵀ
    if ((rc = launch_bank_softmax_rows(P.z, P.ldz, rows, k_q, temperature, P.a_op, P.ldg, 2, S(stream)))) return rc;
    std::memset(&GL, 0, sizeof(GL));
    const Operand A = grad_operand(P.a_op, P.ldg, 2);
    Operand Qp;
    Qp.base = P.qp_op; Qp.nseg = 3; Qp.row_stride = 3 * pad8(k_q); Qp.seg_stride = pad8(k_q);
    if ((rc = fill_gemm_common(GL.job[0], A, 0, rows, Qp, num_classes, k_q, 2))) return rc;
    GL.job[0].mode = GEMM_STORE;
    GL.job[0].out = P.s;
    GL.job[0].ld_out = P.lds;
    GL.njobs = 1;
    gemm_job_tiles(GL);
    if ((rc = launch_gemm(GL, S(stream)))) return rc;
    return launch_smooth_mix(probs, ld_p, P.s, P.lds, rows, num_classes, c_keep, c_bank, out, ld_out, th, max_prob, max_idx,
                             mask, S(stream));
}

namespace {
struct GraphPlan {
    __nv_bfloat16 *p_op, *pu_op;          // probs [rows,3,pad8(C)] | probs_u [C,3,pad8(k_q)]
    __nv_bfloat16 *f0_op, *f1_op, *qs_op; // fp32 features / queue only
    int64_t bytes;
};
GraphPlan plan_graph(void* ws, int64_t ws_bytes, int64_t rows, int64_t k_q, int64_t dim, int64_t c, int dtype) {
    GraphPlan P;
    Workspace W(ws, ws_bytes);
    P.p_op = W.take<__nv_bfloat16>(rows * 3 * pad8(c));
    P.pu_op = W.take<__nv_bfloat16>(c * 3 * pad8(k_q));
    const bool f32 = dtype != STIL_BF16;
    P.f0_op = f32 ? W.take<__nv_bfloat16>(rows * 3 * dim) : nullptr;
    P.f1_op = f32 ? W.take<__nv_bfloat16>(rows * 3 * dim) : nullptr;
    P.qs_op = f32 ? W.take<__nv_bfloat16>(dim * 3 * pad8(k_q)) : nullptr;
    P.bytes = W.off;
    return P;
}
}  // namespace

STIL_API int64_t stil_comatch_graphs_workspace_bytes(int64_t rows, int64_t k_q, int64_t dim, int64_t num_classes, int dtype) {
    return plan_graph(nullptr, 0, rows, k_q, dim, num_classes, dtype).bytes;
}

STIL_API int stil_comatch_graphs_fwd(const float* probs, int64_t ld_p, int64_t rows, int64_t num_classes,
                                     const float* probs_u, int64_t ld_pu, const void* feat_s0, const void* feat_s1, int dtype,
                                     int64_t dim, int64_t ld_f, const void* queue_s, int64_t ld_q, int64_t k_q,
                                     float temperature, float* Q, float* sim, int64_t ld_out, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
    int rc = check_embed(feat_s0, dtype, rows, dim, ld_f, "comatch feat_s0");
    if (rc) return rc;
    if ((rc = check_embed(feat_s1, dtype, rows, dim, ld_f, "comatch feat_s1"))) return rc;
    if ((rc = check_queue(queue_s, dtype, k_q, ld_q, "comatch queue_s"))) return rc;
    STIL_REQUIRE(probs && probs_u && Q && sim && num_classes >= 1 && ld_p >= num_classes && ld_pu >= k_q &&
                     ld_out >= rows + k_q && temperature > 0.f,
                 STIL_E_ARG, "comatch_graphs_fwd: bad arguments");
    GraphPlan P = plan_graph(workspace, workspace_bytes, rows, k_q, dim, num_classes, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "comatch_graphs workspace too small: need %lld",
                 (long long)P.bytes);
    if (rows == 0) return STIL_OK;
    const int64_t cp = pad8(num_classes);
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    prep_add(PL, prep_job_padded(probs, STIL_F32, rows, num_classes, ld_p, P.p_op, cp));
    prep_add(PL, prep_job_padded(probs_u, STIL_F32, num_classes, k_q, ld_pu, P.pu_op, pad8(k_q)));
    if (dtype != STIL_BF16) {
        prep_add(PL, prep_job(feat_s0, dtype, rows, dim, ld_f, 3, P.f0_op, nullptr, 0, 0, nullptr));
        prep_add(PL, prep_job(feat_s1, dtype, rows, dim, ld_f, 3, P.f1_op, nullptr, 0, 0, nullptr));
        prep_add(PL, prep_job_padded(queue_s, dtype, dim, k_q, ld_q, P.qs_op, pad8(k_q)));
    }
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    Operand Pr;
    Pr.base = P.p_op; Pr.nseg = 3; Pr.row_stride = 3 * cp; Pr.seg_stride = cp;
    Operand Pu;
    Pu.base = P.pu_op; Pu.nseg = 3; Pu.row_stride = 3 * pad8(k_q); Pu.seg_stride = pad8(k_q);
    const Operand F0 = rowmajor_operand(feat_s0, dtype, dim, ld_f, P.f0_op, 3);
    const Operand F1 = rowmajor_operand(feat_s1, dtype, dim, ld_f, P.f1_op, 3);
    const Operand Qs = colmajor_operand(queue_s, dtype, k_q, ld_q, P.qs_op);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    // Q_self = probs · probsᵀ with the diagonal forced to 1 (comatch_model.py:299-300)
    if ((rc = fill_gemm_common(GL.job[0], Pr, 0, rows, Pr, rows, num_classes, 2))) return rc;
    GL.job[0].mode = GEMM_STORE; GL.job[0].out = Q; GL.job[0].ld_out = ld_out; GL.job[0].post_op = 2;
    // Q_past = probs · probs_u (:303-304), probs_u [C, k_q] read as an MN-major operand
    if ((rc = fill_gemm_store_mn(GL.job[1], Pr, rows, Pu, num_classes, k_q))) return rc;
    GL.job[1].npair = seg_pairs(3, 3, 2, GL.job[1].xseg, GL.job[1].yseg);
    GL.job[1].out = Q + rows; GL.job[1].ld_out = ld_out;
    // sim_self = exp(f_s0 · f_s1ᵀ / T) (:310)
    if ((rc = fill_gemm_common(GL.job[2], F0, 0, rows, F1, rows, dim, 2))) return rc;
    GL.job[2].mode = GEMM_STORE; GL.job[2].alpha = 1.0f / temperature; GL.job[2].out = sim; GL.job[2].ld_out = ld_out;
    GL.job[2].post_op = 1;
    // sim_past = exp(f_s0 · queue_s / T) (:311-312)
    if ((rc = fill_gemm_store_mn(GL.job[3], F0, rows, Qs, dim, k_q))) return rc;
    GL.job[3].npair = seg_pairs(F0.nseg, Qs.nseg, 2, GL.job[3].xseg, GL.job[3].yseg);
    GL.job[3].alpha = 1.0f / temperature; GL.job[3].out = sim + rows; GL.job[3].ld_out = ld_out; GL.job[3].post_op = 1;
    GL.njobs = 4;
    gemm_job_tiles(GL);
    return launch_gemm(GL, S(stream));
}

STIL_API int stil_comatch_sim_bwd(const float* grad_sim, const float* sim, int64_t ld, int64_t rows, int64_t k_q,
                                  const void* feat_s1, int dtype, int64_t dim, int64_t ld_f, const void* queue_s,
                                  int64_t ld_q, float temperature, void* d_feat_s0, int grad_dtype, int64_t ld_grad,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(grad_sim && sim && d_feat_s0 && ld >= rows + k_q && temperature > 0.f, STIL_E_ARG,
                 "comatch_sim_bwd: bad arguments");
    STIL_REQUIRE(grad_dtype == STIL_F32 || grad_dtype == STIL_BF16, STIL_E_DTYPE, "bad grad dtype");
    int rc = check_embed(feat_s1, dtype, rows, dim, ld_f, "comatch feat_s1");
    if (rc) return rc;
    if ((rc = check_queue(queue_s, dtype, k_q, ld_q, "comatch queue_s"))) return rc;
    STIL_REQUIRE(workspace != nullptr, STIL_E_WORKSPACE, "comatch_sim_bwd: null workspace");
    if (rows == 0) return STIL_OK;
    // the backward has its own workspace (stil_comatch_sim_bwd_workspace_bytes): [gs | gp | g | (fp32) f1_op, qs_op]
    Workspace W(workspace, workspace_bytes);
    const int nseg = grad_nseg(grad_dtype);
    __nv_bfloat16* gs = W.take<__nv_bfloat16>(rows * 2 * pad32(rows));
    __nv_bfloat16* gp = W.take<__nv_bfloat16>(rows * 2 * pad32(k_q));
    float* g = W.take<float>(rows * dim);
    __nv_bfloat16* f1_op = dtype != STIL_BF16 ? W.take<__nv_bfloat16>(rows * 3 * dim) : nullptr;
    __nv_bfloat16* qs_op = dtype != STIL_BF16 ? W.take<__nv_bfloat16>(dim * 3 * pad8(k_q)) : nullptr;
    STIL_REQUIRE(W.off <= workspace_bytes, STIL_E_WORKSPACE, "comatch_sim_bwd workspace too small: need %lld", (long long)W.off);
    if (dtype != STIL_BF16) {
        PrepLaunch PL;
        std::memset(&PL, 0, sizeof(PL));
        prep_add(PL, prep_job(feat_s1, dtype, rows, dim, ld_f, 3, f1_op, nullptr, 0, 0, nullptr));
        prep_add(PL, prep_job_padded(queue_s, dtype, dim, k_q, ld_q, qs_op, pad8(k_q)));
        if ((rc = launch_prep(PL, S(stream)))) return rc;
    }
    if ((rc = launch_sim_grad(grad_sim, sim, ld, rows, rows, k_q, temperature, gs, pad32(rows), gp, pad32(k_q), nseg,
                              S(stream))))
        return rc;
    STIL_CUDA(cudaMemsetAsync(g, 0, rows * dim * sizeof(float), S(stream)));
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    // d_f_s0 = G_self · f_s1 (f_s1 [rows, dim] as an MN-major operand) + G_past · queue_sᵀ (queue_s [dim, k_q] K-major);
    // both jobs add into g (split-contraction mode)
    const Operand Gs = grad_operand(gs, pad32(rows), nseg), Gp = grad_operand(gp, pad32(k_q), nseg);
    const Operand F1 = rowmajor_operand(feat_s1, dtype, dim, ld_f, f1_op, 3);
    const Operand Qs = colmajor_operand(queue_s, dtype, k_q, ld_q, qs_op);
    if ((rc = fill_gemm_store_mn(GL.job[0], Gs, rows, F1, rows, dim))) return rc;
    GL.job[0].out = g; GL.job[0].ld_out = dim;
    GL.job[0].ksplit = (int)std::max<int64_t>(2, std::min<int64_t>(ceil_div(rows, kTileK), 8));
    if ((rc = fill_gemm_common(GL.job[1], Gp, 0, rows, Qs, dim, k_q, 1))) return rc;
    GL.job[1].mode = GEMM_STORE; GL.job[1].out = g; GL.job[1].ld_out = dim;
    GL.job[1].ksplit = (int)std::max<int64_t>(2, std::min<int64_t>(ceil_div(k_q, kTileK) / 2, 16));
    GL.njobs = 2;
    gemm_job_tiles(GL);
    if ((rc = launch_gemm(GL, S(stream)))) return rc;
    GradFinishLaunch GF;
    std::memset(&GF, 0, sizeof(GF));
    GradFinishJob& j = GF.job[0];
    j.g = g; j.dx = d_feat_s0; j.dx_dtype = grad_dtype; j.ld_dx = ld_grad;
    j.rows = (int)rows; j.dim = (int)dim;
    GF.njobs = 1;
    GF.total_rows = (int)rows;
    return launch_grad_finish(GF, S(stream));
}

STIL_API int64_t stil_comatch_sim_bwd_workspace_bytes(int64_t rows, int64_t k_q, int64_t dim, int dtype) {
    Workspace W(nullptr, 0);
    W.take<__nv_bfloat16>(rows * 2 * pad32(rows));
    W.take<__nv_bfloat16>(rows * 2 * pad32(k_q));
    W.take<float>(rows * dim);
    if (dtype != STIL_BF16) {
        W.take<__nv_bfloat16>(rows * 3 * dim);
        W.take<__nv_bfloat16>(dim * 3 * pad8(k_q));
    }
    return W.off;
}

STIL_API int64_t stil_row_loss_workspace_bytes(int64_t rows) {
    Workspace W(nullptr, 0);
    W.take<unsigned int>(64);
    W.take<float>(std::max<int64_t>(rows, weighted_softce_blocks(rows)) + 8);
    return W.off;
}

STIL_API int stil_graph_contrast_loss(const float* Q, const float* sim, int64_t ld, int64_t rows, int64_t cols,
                                      float contrast_th, float* loss, float* d_sim, float grad_scale, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    STIL_REQUIRE(Q && sim && loss && rows >= 1 && cols >= 1 && ld >= cols, STIL_E_ARG, "graph_contrast_loss: bad arguments");
    STIL_REQUIRE(workspace && workspace_bytes >= stil_row_loss_workspace_bytes(rows), STIL_E_WORKSPACE,
                 "graph_contrast_loss workspace too small");
    Workspace W(workspace, workspace_bytes);
    unsigned int* ticket = W.take<unsigned int>(64);
    float* partials = W.take<float>(rows + 8);
    STIL_CUDA(cudaMemsetAsync(ticket, 0, 16, S(stream)));
    return launch_graph_contrast(Q, sim, ld, rows, cols, contrast_th, d_sim, grad_scale, partials, ticket, loss, S(stream));
}

STIL_API int stil_weighted_softce(const void* logits, int logit_dtype, int64_t ld_y, const float* target_probs, int64_t ld_t,
                                  const int64_t* target_idx, const uint8_t* mask, int64_t rows, int64_t k, float* loss,
                                  float* d_logits, int64_t ld_g, float grad_scale, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
    STIL_REQUIRE(logit_dtype == STIL_F32 || logit_dtype == STIL_BF16, STIL_E_DTYPE, "weighted_softce: bad logit dtype");
    STIL_REQUIRE(logits && loss && rows >= 1 && k >= 1 && ((target_probs != nullptr) != (target_idx != nullptr)), STIL_E_ARG,
                 "weighted_softce: needs logits, loss and exactly one of target_probs / target_idx");
    STIL_REQUIRE(workspace && workspace_bytes >= stil_row_loss_workspace_bytes(rows), STIL_E_WORKSPACE,
                 "weighted_softce workspace too small");
    Workspace W(workspace, workspace_bytes);
    unsigned int* ticket = W.take<unsigned int>(64);
    float* partials = W.take<float>(std::max<int64_t>(rows, weighted_softce_blocks(rows)) + 8);
    STIL_CUDA(cudaMemsetAsync(ticket, 0, 16, S(stream)));
    return launch_weighted_softce(logits, logit_dtype, ld_y, target_probs, ld_t, target_idx, mask, rows, k, d_logits, ld_g,
                                  grad_scale, partials, ticket, loss, S(stream));
}

STIL_API int stil_queue_enqueue(void* queue_feat, int q_dtype, int64_t ld_q, float* queue_probs, int64_t ld_qp, int64_t k_q,
                                int64_t* ptr, const void* z, int z_dtype, int64_t ld_z, int64_t n, int64_t dim, const float* t,
                                int64_t ld_t, int64_t num_classes, void* stream) {
    STIL_REQUIRE(queue_feat && queue_probs && ptr && k_q >= 1 && ld_q >= k_q && ld_qp >= k_q && n >= 0 && (n == 0 || (z && t)),
                 STIL_E_ARG, "queue_enqueue: bad arguments");
    STIL_REQUIRE((q_dtype == STIL_F32 || q_dtype == STIL_BF16) && (z_dtype == STIL_F32 || z_dtype == STIL_BF16), STIL_E_DTYPE,
                 "queue_enqueue: bad dtype");
    return launch_queue_enqueue(queue_feat, q_dtype, ld_q, queue_probs, ld_qp, k_q, ptr, z, z_dtype, ld_z, n, dim, t, ld_t,
                                num_classes, S(stream));
}

STIL_API int stil_bank_update(void* bank, int b_dtype, int64_t ld_bank, int64_t* labels, const void* k, int k_dtype,
                              int64_t ld_k, const int64_t* y, const int64_t* index, int64_t n, int64_t dim, void* stream) {
    STIL_REQUIRE(bank && labels && n >= 0 && (n == 0 || (k && y && index)), STIL_E_ARG, "bank_update: bad arguments");
    STIL_REQUIRE((b_dtype == STIL_F32 || b_dtype == STIL_BF16) && (k_dtype == STIL_F32 || k_dtype == STIL_BF16), STIL_E_DTYPE,
                 "bank_update: bad dtype");
    return launch_bank_update(bank, b_dtype, ld_bank, labels, k, k_dtype, ld_k, y, index, n, dim, S(stream));
}

STIL_API int stil_club_fwd(const void* mu, const void* y, int dtype, int64_t rows, int64_t dim, int64_t ld, float* colstats,
                           float* bound, float* est, void* stream) {
    STIL_REQUIRE(dtype == STIL_F32 || dtype == STIL_BF16, STIL_E_DTYPE, "club_fwd: bad dtype");
    STIL_REQUIRE(mu && y && colstats && rows >= 1 && dim >= 1 && ld >= dim && (bound || est), STIL_E_ARG, "club_fwd: bad arguments");
    return launch_club_fwd(mu, y, dtype, ld, rows, dim, colstats, bound, est, S(stream));
}
STIL_API int stil_club_bwd(const void* mu, const void* y, int dtype, int64_t rows, int64_t dim, int64_t ld, const float* colstats,
                           const float* g_bound, const float* g_est, float* d_mu, float* d_y, int64_t ld_grad, void* stream) {
    STIL_REQUIRE(dtype == STIL_F32 || dtype == STIL_BF16, STIL_E_DTYPE, "club_bwd: bad dtype");
    STIL_REQUIRE(mu && y && colstats && d_mu && d_y && rows >= 1 && dim >= 1 && ld >= dim && ld_grad >= dim, STIL_E_ARG,
                 "club_bwd: bad arguments");
    return launch_club_bwd(mu, y, dtype, ld, rows, dim, colstats, g_bound, g_est, d_mu, d_y, ld_grad, S(stream));
}

STIL_API int stil_da_apply_hist(const float* probs, int64_t ld, int64_t rows, int64_t k, const float* batch_mean, float* hist,
                                int64_t hist_len, int64_t* count, float* qmean_scratch, float* out, int64_t ld_out,
                                void* stream) {
    STIL_REQUIRE(probs && batch_mean && hist && count && qmean_scratch && out && hist_len >= 1, STIL_E_ARG,
                 "da_apply_hist: bad arguments");
    int rc = launch_da_hist_update(batch_mean, hist, hist_len, k, count, qmean_scratch, S(stream));
    if (rc) return rc;
    return launch_da_rows(probs, ld, rows, k, qmean_scratch, out, ld_out, S(stream));
}

// =============================================================================================== f-3
// Linear (+ bias) (+ F.normalize) feeding the head: projector_imaging / projector_tabular = nn.Linear(512, 128) followed by
// F.normalize (STiLModel.py:56-63, 182-192) and the three classifier Linears (STiLModel_backbone.py:66-68, 153-155).
namespace {
struct LinearPlan {
    __nv_bfloat16 *x_op, *w_op, *g_op;   // [rows,3,in] (fp32 x only) | [out,3,in] | [rows,3,out]
    float* g;                            // [rows, out] gradient w.r.t. the pre-normalisation output
    int64_t bytes;
};
LinearPlan plan_linear(void* ws, int64_t ws_bytes, int64_t rows, int64_t in_dim, int64_t out_dim, int dtype) {
    LinearPlan P;
    Workspace W(ws, ws_bytes);
    P.x_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(rows * 3 * in_dim);
    P.w_op = W.take<__nv_bfloat16>(out_dim * 3 * in_dim);
    P.g_op = W.take<__nv_bfloat16>(rows * 3 * round_up(out_dim, 8));   // rows padded to 16-byte granularity
    P.g = W.take<float>(rows * out_dim);
    P.bytes = W.off;
    return P;
}
int linear_checks(const void* x, int dtype, int64_t rows, int64_t in_dim, int64_t ld_x, const float* weight, int64_t out_dim,
                  const char* what) {
    int rc = check_embed(x, dtype, rows, in_dim, ld_x, what);
    if (rc) return rc;
    if ((rc = check_embed(weight, STIL_F32, out_dim, in_dim, in_dim, "linear weight"))) return rc;
    STIL_REQUIRE(out_dim >= 1 && out_dim <= 16384, STIL_E_SHAPE, "%s: out_dim %lld out of range", what, (long long)out_dim);
    return STIL_OK;
}
}  // namespace

STIL_API int64_t stil_linear_workspace_bytes(int64_t rows, int64_t in_dim, int64_t out_dim, int dtype) {
    return plan_linear(nullptr, 0, rows, in_dim, out_dim, dtype).bytes;
}

STIL_API int stil_linear_fwd(const void* x, int dtype, int64_t rows, int64_t in_dim, int64_t ld_x, const float* weight,
                             const float* bias, int64_t out_dim, int normalize, float* y, int64_t ld_y, float* y_raw,
                             float* inv_norm, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = linear_checks(x, dtype, rows, in_dim, ld_x, weight, out_dim, "linear_fwd x");
    if (rc) return rc;
    STIL_REQUIRE(y && ld_y >= out_dim, STIL_E_ARG, "linear_fwd: bad output");
    STIL_REQUIRE(!normalize || (out_dim <= kTileN && y_raw && inv_norm), STIL_E_SHAPE,
                 "linear_fwd: the fused F.normalize needs out_dim <= %d and the y_raw / inv_norm outputs", kTileN);
    LinearPlan P = plan_linear(workspace, workspace_bytes, rows, in_dim, out_dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "linear workspace too small: need %lld", (long long)P.bytes);
    if (rows == 0) return STIL_OK;
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    if (dtype != STIL_BF16) prep_add(PL, prep_job(x, dtype, rows, in_dim, ld_x, 3, P.x_op, nullptr, 0, 0, nullptr));
    prep_add(PL, prep_job(weight, STIL_F32, out_dim, in_dim, in_dim, 3, P.w_op, nullptr, 0, 0, nullptr));
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    const Operand X = rowmajor_operand(x, dtype, in_dim, ld_x, P.x_op, 3);
    const Operand Wt = rowmajor_operand(nullptr, STIL_F32, in_dim, in_dim, P.w_op, 3);
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    if ((rc = fill_gemm_common(GL.job[0], X, 0, rows, Wt, out_dim, in_dim))) return rc;      // y = x · Wᵀ, fp32-accurate pairs
    GemmJob& J = GL.job[0];
    J.mode = GEMM_STORE;
    J.out = y; J.ld_out = ld_y;
    J.col_bias = bias;
    J.fwd_norm = normalize ? 1 : 0;
    J.fwd_inv_norm = inv_norm;
    J.fwd_raw = y_raw; J.fwd_ld_raw = out_dim;
    GL.njobs = 1;
    gemm_job_tiles(GL);
    return launch_gemm(GL, S(stream));
}

STIL_API int stil_linear_bwd(const void* x, int dtype, int64_t rows, int64_t in_dim, int64_t ld_x, const float* weight,
                             int64_t out_dim, const float* y_raw, const float* inv_norm, const float* d_y, int64_t ld_dy,
                             void* d_x, int grad_dtype, int64_t ld_dx, float* d_weight, float* d_bias, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    int rc = linear_checks(x, dtype, rows, in_dim, ld_x, weight, out_dim, "linear_bwd x");
    if (rc) return rc;
    STIL_REQUIRE(d_y && ld_dy >= out_dim && (d_x || d_weight || d_bias), STIL_E_ARG, "linear_bwd: bad arguments");
    STIL_REQUIRE((inv_norm == nullptr) == (y_raw == nullptr), STIL_E_ARG, "linear_bwd: y_raw and inv_norm go together");
    STIL_REQUIRE(grad_dtype == STIL_F32 || grad_dtype == STIL_BF16, STIL_E_DTYPE, "bad grad dtype");
    LinearPlan P = plan_linear(workspace, workspace_bytes, rows, in_dim, out_dim, dtype);
    STIL_REQUIRE(workspace && P.bytes <= workspace_bytes, STIL_E_WORKSPACE, "linear workspace too small");
    if (rows == 0) {
        if (d_weight) STIL_CUDA(cudaMemsetAsync(d_weight, 0, out_dim * in_dim * sizeof(float), S(stream)));
        if (d_bias) STIL_CUDA(cudaMemsetAsync(d_bias, 0, out_dim * sizeof(float), S(stream)));
        return STIL_OK;
    }
    // 1. g = backward of F.normalize applied to d_y (or d_y itself): [rows, out] f32
    const float* g = d_y;
    int64_t ld_g = ld_dy;
    if (inv_norm) {
        GradFinishLaunch GF;
        std::memset(&GF, 0, sizeof(GF));
        GradFinishJob& j = GF.job[0];
        STIL_REQUIRE(ld_dy == out_dim, STIL_E_SHAPE, "linear_bwd: d_y must be contiguous when the output was normalised");
        j.g = d_y; j.x = y_raw; j.x_dtype = STIL_F32; j.ldx = out_dim; j.sx = inv_norm;
        j.dx = P.g; j.dx_dtype = STIL_F32; j.ld_dx = out_dim;
        j.rows = (int)rows; j.dim = (int)out_dim;
        GF.njobs = 1;
        GF.total_rows = (int)rows;
        if ((rc = launch_grad_finish(GF, S(stream)))) return rc;
        g = P.g;
        ld_g = out_dim;
    }
    if (d_bias && (rc = launch_col_sum(g, ld_g, rows, out_dim, d_bias, S(stream)))) return rc;
    // 2. operands: g as three bf16 segments [rows, 3, out]; the weight (and an fp32 x) likewise
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    const int64_t og = round_up(out_dim, 8);
    {
        PrepJob pj = prep_job(g, STIL_F32, rows, out_dim, ld_g, 3, P.g_op, nullptr, 0, 0, nullptr);
        pj.op_dim = (int)og;          // zero-filled beyond out_dim
        prep_add(PL, pj);
    }
    if (d_x) prep_add(PL, prep_job(weight, STIL_F32, out_dim, in_dim, in_dim, 3, P.w_op, nullptr, 0, 0, nullptr));
    if (d_weight && dtype != STIL_BF16) prep_add(PL, prep_job(x, dtype, rows, in_dim, ld_x, 3, P.x_op, nullptr, 0, 0, nullptr));
    if ((rc = launch_prep(PL, S(stream)))) return rc;
    Operand G;
    G.base = P.g_op; G.nseg = 3; G.row_stride = 3 * og; G.seg_stride = og;
    GemmLaunch GL;
    std::memset(&GL, 0, sizeof(GL));
    int nj = 0;
    if (d_x) {
        // d_x = g · W: contraction over out; W [out, 3, in] read as an MN-major operand (in contiguous)
        Operand Wm;
        Wm.base = P.w_op; Wm.nseg = 3; Wm.row_stride = 3 * in_dim; Wm.seg_stride = in_dim;
        GemmJob& J = GL.job[nj++];
        if ((rc = fill_gemm_store_mn(J, G, rows, Wm, out_dim, in_dim))) return rc;
        J.npair = seg_pairs(G.nseg, Wm.nseg, 2, J.xseg, J.yseg);
        J.fin_dx = d_x; J.fin_dx_dtype = grad_dtype; J.fin_ld_dx = ld_dx;       // plain cast / store epilogue
    }
    if (d_weight) {
        // d_W = gᵀ · x: contraction over the ROWS of both row-major matrices (X and Y MN-major)
        Operand Xm;
        if (dtype == STIL_BF16) { Xm.base = static_cast<const __nv_bfloat16*>(x); Xm.nseg = 1; Xm.row_stride = ld_x; Xm.seg_stride = in_dim; }
        else { Xm.base = P.x_op; Xm.nseg = 3; Xm.row_stride = 3 * in_dim; Xm.seg_stride = in_dim; }
        GemmJob& J = GL.job[nj++];
        std::memset(&J, 0, sizeof(J));
        if ((rc = make_operand_map(&J.tmx, G.base, out_dim, rows, G.nseg, G.row_stride, G.seg_stride, 64))) return rc;
        if ((rc = make_operand_map(&J.tmy, Xm.base, in_dim, rows, Xm.nseg, Xm.row_stride, Xm.seg_stride, 64))) return rc;
        J.M = (int)out_dim; J.N = (int)in_dim; J.D = (int)rows;
        J.npair = seg_pairs(G.nseg, Xm.nseg, 2, J.xseg, J.yseg);
        J.alpha = 1.f;
        J.mode = GEMM_STORE;
        J.x_mn_major = 1; J.y_mn_major = 1;
        J.out = d_weight; J.ld_out = in_dim;
    }
    GL.njobs = nj;
    gemm_job_tiles(GL);
    return launch_gemm(GL, S(stream));
}

// =============================================================================================== f-4
STIL_API int stil_ema_update(const stil_ema_entry* table, int64_t n_entries, const int32_t* chunk_entry,
                             const int64_t* chunk_start, int64_t n_chunks, int64_t chunk_elems, double momentum, void* stream) {
    STIL_REQUIRE(n_chunks == 0 || (table && chunk_entry && chunk_start && n_entries >= 1), STIL_E_ARG, "ema_update: null pointer");
    STIL_REQUIRE(chunk_elems >= 1 && chunk_elems % 16 == 0, STIL_E_ARG, "ema_update: chunk_elems must be a positive multiple of 16");
    STIL_REQUIRE(n_chunks < (1LL << 31), STIL_E_SHAPE, "ema_update: too many chunks");
    return launch_ema_update(table, chunk_entry, chunk_start, n_chunks, chunk_elems, momentum, S(stream));
}

// =============================================================================================== whole step
namespace {
struct StepPlan {
    InfoncePlan nce;
    ProtoPlan pt;        // student feat_m x prototypes
    __nv_bfloat16* teach_op;   // teacher feat_m_e operand (fp32 input only)
    float* teacher_logits;     // [b_u, ldk]
    float *teach_pmax, *teach_psum;
    float* z_pt;               // [batch, ldk] student prototype logits (for the CE coefficient)
    int64_t ldk;
    int32_t* cls;              // [batch]
    uint8_t* conf;             // [batch]
    float *lse_row, *lse_col, *lse_pt, *w_pt;
    float* ce_partials;
    unsigned int* ce_ticket;
    int64_t bytes;
};

StepPlan plan_step(void* ws, int64_t ws_bytes, int64_t batch, int64_t b_l, int64_t k, int64_t dim, int dtype) {
    StepPlan P;
    const int64_t b_u = batch - b_l;
    // sub-plans are laid out back to back inside the same workspace
    InfoncePlan n0 = plan_infonce(nullptr, 0, batch, batch, dim, dtype, true);
    ProtoPlan p0 = plan_proto(nullptr, 0, batch, k, dim, dtype);
    char* base = static_cast<char*>(ws);
    P.nce = plan_infonce(base, base ? n0.bytes : 0, batch, batch, dim, dtype, true);
    P.pt = plan_proto(base ? base + n0.bytes : nullptr, base ? p0.bytes : 0, batch, k, dim, dtype);
    Workspace W(base ? base + n0.bytes + p0.bytes : nullptr, ws_bytes - n0.bytes - p0.bytes);
    P.ldk = round_up(k, 4);
    P.teach_op = dtype == STIL_BF16 ? nullptr : W.take<__nv_bfloat16>(b_u * 3 * dim);
    P.teacher_logits = W.take<float>(b_u * P.ldk);
    P.teach_pmax = W.take<float>(stat_slots(k) * b_u);
    P.teach_psum = W.take<float>(stat_slots(k) * b_u);
    P.z_pt = W.take<float>(batch * P.ldk);
    P.cls = W.take<int32_t>(batch);
    P.conf = W.take<uint8_t>(batch);
    P.lse_row = W.take<float>(batch);
    P.lse_col = W.take<float>(batch);
    P.lse_pt = W.take<float>(batch);
    P.w_pt = W.take<float>(batch);
    P.ce_ticket = P.nce.ticket ? P.nce.ticket + 16 : nullptr;
    P.ce_partials = W.take<float>(3 * masked_softce_blocks(b_u, k) + 8);
    P.bytes = n0.bytes + p0.bytes + W.off;
    return P;
}
}  // namespace

STIL_API int64_t stil_head_step_workspace_bytes(int64_t batch, int64_t b_l, int64_t k, int64_t dim, int embed_dtype) {
    return plan_step(nullptr, 0, batch, b_l, k, dim, embed_dtype).bytes;
}

STIL_API int stil_head_prepare_prototypes(const stil_head_step_args* a) {
    STIL_REQUIRE(a != nullptr && a->prototypes && a->workspace, STIL_E_ARG, "head_prepare_prototypes: null argument");
    StepPlan P = plan_step(a->workspace, a->workspace_bytes, a->batch, a->b_l, a->k, a->dim, a->embed_dtype);
    STIL_REQUIRE(P.bytes <= a->workspace_bytes, STIL_E_WORKSPACE, "head_step workspace too small: need %lld", (long long)P.bytes);
    PrepLaunch PL;
    std::memset(&PL, 0, sizeof(PL));
    prep_add(PL, prep_job(a->prototypes, STIL_F32, a->k, a->dim, a->dim, P.pt.proto_nseg, P.pt.proto_op, nullptr, 0, 0, nullptr));
    return launch_prep(PL, S(a->stream));
}

STIL_API int stil_head_step_launches(const stil_head_step_args* a) {
    if (!a) return 0;
    // gemm(stats), cgpl_pgls, gemm(grad)+gemm(dX) for the prototype CE, finish, proto_accumulate
    int n = 6;
    if (a->embed_dtype != STIL_BF16 || !a->prototypes_prepared) n += 1;   // main-stream prep
    if (!a->skip_infonce) n += 4;          // prep, gemm(stats), gemm(grad), gemm(dX) for the InfoNCE
    if (a->dim > kTileN) n += a->skip_infonce ? 1 : 2;   // un-fused normalise-backward / cast
    if (a->y_m) n += 1;                    // masked soft CE
    return n;
}

STIL_API int stil_head_step(const stil_head_step_args* a) {
    STIL_REQUIRE(a != nullptr, STIL_E_ARG, "head_step: null args");
    const int64_t B = a->batch, B_l = a->b_l, B_u = a->batch - a->b_l, K = a->k, D = a->dim;
    const int dt = a->embed_dtype;
    cudaStream_t st = S(a->stream);
    STIL_REQUIRE(B >= 1 && B_l >= 0 && B_l <= B && K >= 1, STIL_E_SHAPE, "head_step: bad sizes");
    int rc;
    if ((rc = check_embed(a->feat_i, dt, B, D, D, "feat_i"))) return rc;
    if ((rc = check_embed(a->feat_t, dt, B, D, D, "feat_t"))) return rc;
    if ((rc = check_embed(a->feat_m, dt, B, D, D, "feat_m"))) return rc;
    if ((rc = check_embed(a->feat_m_e, dt, B, D, D, "feat_m_e"))) return rc;
    if ((rc = check_embed(a->prototypes, STIL_F32, K, D, D, "prototypes"))) return rc;
    STIL_REQUIRE(a->lambda0 >= 0.f && a->lambda0 <= 1.f, STIL_E_ARG, "lambda_0 must be a float between 0 and 1.");
    STIL_REQUIRE(a->temperature > 0.f && a->repeat_ratio > 0.f, STIL_E_ARG, "head_step: bad temperature / repeat_ratio");
    STIL_REQUIRE(a->losses && a->d_feat_i && a->d_feat_t && a->d_feat_m && a->pseudo_label && a->max_idx && a->mask1 &&
                     a->case1 && a->case2_i && a->case2_t && a->case3 && a->class_sum && a->class_count && a->y_l &&
                     a->y_m_ue && a->y_i_ue && a->y_t_ue,
                 STIL_E_ARG, "head_step: null pointer");
    STIL_REQUIRE(a->grad_dtype == STIL_F32 || a->grad_dtype == STIL_BF16, STIL_E_DTYPE, "head_step: bad grad dtype");
    StepPlan P = plan_step(a->workspace, a->workspace_bytes, B, B_l, K, D, dt);
    STIL_REQUIRE(a->workspace && P.bytes <= a->workspace_bytes, STIL_E_WORKSPACE, "head_step workspace too small: need %lld",
                 (long long)P.bytes);
    SideStreams* SS = nullptr;
    if ((rc = get_side_streams(&SS))) return rc;
    // The side streams and their fork / join events are shared by every caller on this device: the whole enqueue (a few
    // dozen microseconds of host time) is serialised, so two host threads or two heads driving the same device cannot
    // interleave each other's event records and waits.  The caller still owns the workspace: one per concurrent step.
    static std::mutex enqueue_mutex[64];
    int dev_id = 0;
    STIL_CUDA(cudaGetDevice(&dev_id));
    std::lock_guard<std::mutex> enqueue_lock(enqueue_mutex[dev_id & 63]);
    cudaStream_t s_loss = SS->s[0], s_acc = SS->s[1], s_ce = SS->s[2], s_nce = SS->s[3];
    const float inv_t = 1.0f / a->temperature;
    const int esz = dt == STIL_BF16 ? 2 : 4;
    const void* feat_m_ue = static_cast<const char*>(a->feat_m_e) + B_l * D * esz;
    const bool nce = !a->skip_infonce;
    // optional instrumentation (bench.py): event pairs around the GEMM launches and the pseudo-label kernel
    cudaEvent_t* tev = reinterpret_cast<cudaEvent_t*>(a->timing_events);
    auto mark = [&](int i, cudaStream_t s) -> int {
        if (tev && i < a->n_timing_events) STIL_CUDA(cudaEventRecord(tev[i], s));
        return STIL_OK;
    };

    // ---- fork A: the InfoNCE (forward statistics and backward) is independent of the pseudo-label chain and runs on
    //      its own stream from the start of the step
    STIL_CUDA(cudaEventRecord(SS->fork, st));
    const Operand A = rowmajor_operand(a->feat_i, dt, D, D, P.nce.a_op, P.nce.nseg);
    const Operand Bm = rowmajor_operand(a->feat_t, dt, D, D, P.nce.b_op, P.nce.nseg);
    GemmLaunch GL;

    // 1. main stream: operand preparation — fp32 -> bf16 segment split of the prototypes (skipped when the caller
    //    prepared them once with stil_head_prepare_prototypes and they have not changed) and of fp32 features
    {
        PrepLaunch PL;
        std::memset(&PL, 0, sizeof(PL));
        if (dt != STIL_BF16) {
            prep_add(PL, prep_job(a->feat_m, dt, B, D, D, P.pt.feat_nseg, P.pt.feat_op, nullptr, 0, 0, nullptr));
            prep_add(PL, prep_job(feat_m_ue, dt, B_u, D, D, 3, P.teach_op, nullptr, 0, 0, nullptr));
        }
        if (!a->prototypes_prepared)
            prep_add(PL, prep_job(a->prototypes, STIL_F32, K, D, D, P.pt.proto_nseg, P.pt.proto_op, nullptr, 0, 0, nullptr));
        if ((rc = mark(0, st))) return rc;
        if (PL.njobs > 0 && (rc = launch_prep(PL, st))) return rc;
    }

    // 2. main stream: student prototype logits (statistics + store) and teacher prototype logits (store), one launch
    std::memset(&GL, 0, sizeof(GL));
    int nj = 0;
    if ((rc = proto_stats_job(GL.job[nj], P.pt, a->feat_m, dt, B, D, D, K, inv_t))) return rc;
    GL.job[nj].out = P.z_pt;
    GL.job[nj].ld_out = P.ldk;
    ++nj;
    if (B_u > 0) {
        const Operand X = rowmajor_operand(feat_m_ue, dt, D, D, P.teach_op, 3);
        const Operand Y = rowmajor_operand(nullptr, STIL_F32, D, D, P.pt.proto_op, P.pt.proto_nseg);
        if ((rc = fill_gemm_common(GL.job[nj], X, 0, B_u, Y, K, D))) return rc;
        GL.job[nj].mode = GEMM_STATS;            // same launch as the statistics job; its partials are unused
        GL.job[nj].part_max = P.teach_pmax;
        GL.job[nj].part_sum = P.teach_psum;
        GL.job[nj].out = P.teacher_logits;
        GL.job[nj].ld_out = P.ldk;
        ++nj;
    }
    GL.njobs = nj;
    gemm_job_tiles(GL);
    if ((rc = mark(1, st))) return rc;
    if ((rc = launch_gemm(GL, st))) return rc;
    if ((rc = mark(5, st))) return rc;

    // 3. main stream: CGPL + PGLS on the unlabelled rows, (cls, conf) of every row ...
    if ((rc = launch_cgpl_pgls(a->y_m_ue, a->y_i_ue, a->y_t_ue, a->logit_dtype, K, P.teacher_logits, P.ldk, B_u, K,
                               a->temperature, a->rate_pseudo, a->th1, a->past_start_epoch, a->prediction_in, K,
                               a->pseudo_label, K, nullptr, 0, a->max_prob, a->max_idx, a->mask1, a->case1, a->case2_i, a->case2_t, a->case3,
                               nullptr, P.cls + B_l, P.conf + B_l, a->y_l, B_l, P.cls, P.conf, st)))
        return rc;
    if ((rc = mark(6, st))) return rc;
    // ---- fork B: everything that only needs the pseudo labels
    STIL_CUDA(cudaEventRecord(SS->fork2, st));
    //    ... then the prototype-CE backward on the same stream
    bool fused_pt = false;
    {
        std::memset(&GL, 0, sizeof(GL));
        if ((rc = proto_grad_job(GL.job[0], P.pt, a->feat_m, dt, B, D, D, K, inv_t, P.cls, nullptr, nullptr, nullptr,
                                 P.z_pt, P.ldk, P.conf, a->grad_dtype)))
            return rc;
        GL.njobs = 1;
        gemm_job_tiles(GL);
        set_early(GL, true, true);    // predecessor = cgpl_pgls: both operands were final two kernels ago ...
        GL.job[0].early_stats = 1;    // ... and so were the statistics of the student logits
        GemmLaunch GS;
        std::memset(&GS, 0, sizeof(GS));
        bool done = false;
        if ((rc = proto_bwd_fused(GL, P.pt, B, D, K, a->grad_dtype, st, &done))) return rc;   // recompute + dLogits + dX
        if (done) {
            if ((rc = mark(7, st))) return rc;
            if ((rc = launch_proto_gradfinish(P.pt, B, D, K, a->d_feat_m, a->grad_dtype, D, true, st))) return rc;
            if ((rc = mark(8, st))) return rc;
            fused_pt = true;     // d_feat_m is written: nothing left to finish below
        } else {
            if ((rc = proto_store_job(GS.job[0], P.pt, B, D, K, a->d_feat_m, a->grad_dtype, D, &fused_pt))) return rc;
            GS.njobs = 1;
            GS.cluster_k = fused_pt ? GS.job[0].ksplit : 1;
            gemm_job_tiles(GS);
            if ((rc = launch_gemm(GL, st))) return rc;
            if ((rc = mark(7, st))) return rc;
            set_early(GS, false, true);   // predecessor = GRAD (writes X = G); Y = prototype operand
            if ((rc = launch_gemm(GS, st))) return rc;
            if ((rc = mark(8, st))) return rc;
        }
        if (!fused_pt) {
            GradFinishLaunch GF;
            std::memset(&GF, 0, sizeof(GF));
            GradFinishJob& j = GF.job[0];
            j.g = P.pt.g; j.dx = a->d_feat_m; j.dx_dtype = a->grad_dtype; j.ld_dx = D;
            j.rows = (int)B; j.dim = (int)D; j.row_begin = 0;
            GF.njobs = 1;
            GF.total_rows = (int)B;
            if ((rc = launch_grad_finish(GF, st))) return rc;
        }
    }

    // 4. InfoNCE stream: inverse norms (+ fp32 split) -> statistics of a·bT and b·aT -> G tiles (statistics merged
    //    in-kernel) -> dX = G · Y with the normalise-backward in the epilogue
    if (nce) {
        STIL_CUDA(cudaStreamWaitEvent(s_nce, SS->fork, 0));
        PrepLaunch PL;
        std::memset(&PL, 0, sizeof(PL));
        prep_add(PL, prep_job(a->feat_i, dt, B, D, D, P.nce.nseg, P.nce.a_op, nullptr, 0, 0, P.nce.ra));
        prep_add(PL, prep_job(a->feat_t, dt, B, D, D, P.nce.nseg, P.nce.b_op, nullptr, 0, 0, P.nce.rb));
        if ((rc = mark(9, s_nce))) return rc;
        if ((rc = launch_prep(PL, s_nce))) return rc;
        std::memset(&GL, 0, sizeof(GL));
        if ((rc = infonce_stats_jobs(GL.job, P.nce, A, Bm, B, B, D, 0, inv_t, nullptr, 0, &GL.njobs))) return rc;
        gemm_job_tiles(GL);
        set_early(GL, dt == STIL_BF16, dt == STIL_BF16);   // predecessor = prep: bf16 operands are the caller's inputs
        if ((rc = mark(10, s_nce))) return rc;
        if ((rc = launch_gemm(GL, s_nce))) return rc;
        if ((rc = mark(2, s_nce))) return rc;
        STIL_CUDA(cudaEventRecord(SS->nce_stats, s_nce));
        bool fused = true;
        std::memset(&GL, 0, sizeof(GL));
        if ((rc = infonce_grad_jobs(GL.job, P.nce, A, Bm, B, B, D, 0, inv_t, a->lambda0, nullptr, nullptr, nullptr,
                                    a->grad_dtype, &GL.njobs)))
            return rc;
        gemm_job_tiles(GL);
        set_early(GL, true, true);    // predecessor = STATS: operands were final two kernels ago
        GemmLaunch GS, GB;
        std::memset(&GS, 0, sizeof(GS));
        int dx_ck = 1;
        if ((rc = infonce_store_jobs(GS.job, P.nce, A, Bm, a->feat_i, a->feat_t, dt, B, B, D, D, 0, a->d_feat_i,
                                     a->d_feat_t, a->grad_dtype, D, inv_t, &fused, &dx_ck)))
            return rc;
        GS.njobs = 2;
        GS.cluster_k = dx_ck;
        gemm_job_tiles(GS);
        if (fuse_bwd(GB, GL, GS)) {
            if ((rc = launch_gemm(GB, s_nce))) return rc;
            if ((rc = mark(3, s_nce))) return rc;
            if ((rc = mark(4, s_nce))) return rc;
        } else {
            if ((rc = launch_gemm(GL, s_nce))) return rc;
            if ((rc = mark(3, s_nce))) return rc;
            set_early(GS, false, true);   // predecessor = GRAD (writes X = G)
            if ((rc = launch_gemm(GS, s_nce))) return rc;
            if ((rc = mark(4, s_nce))) return rc;
        }
        if (!fused) {
            GradFinishLaunch GF;
            std::memset(&GF, 0, sizeof(GF));
            infonce_gradfinish_jobs(GF.job, P.nce, a->feat_i, a->feat_t, dt, B, B, D, D, 0, a->d_feat_i, a->d_feat_t,
                                    a->grad_dtype, D);
            GF.njobs = 2;
            GF.total_rows = (int)(2 * B);
            if ((rc = launch_grad_finish(GF, s_nce))) return rc;
        }
    }

    STIL_CUDA(cudaStreamWaitEvent(s_loss, SS->fork2, 0));
    if (nce) STIL_CUDA(cudaStreamWaitEvent(s_loss, SS->nce_stats, 0));
    STIL_CUDA(cudaStreamWaitEvent(s_acc, SS->fork2, 0));
    if (a->y_m) STIL_CUDA(cudaStreamWaitEvent(s_ce, SS->fork2, 0));

    // 5. side branches
    {   // losses and LSE vectors
        FinishLaunch FL;
        std::memset(&FL, 0, sizeof(FL));
        int fj = 0;
        if (nce) {
            infonce_finish_jobs(FL.job, P.nce, a->feat_i, a->feat_t, dt, B, B, D, D, 0, inv_t, a->lambda0, P.lse_row,
                                P.lse_col, 0);
            FL.job[0].row_begin = 0;
            FL.job[1].row_begin = (int)B;
            fj = 2;
        }
        proto_finish_job(FL.job[fj], P.pt, a->feat_m, dt, B, D, D, a->prototypes, K, P.cls, P.conf, inv_t, P.lse_pt,
                         P.w_pt, 1);
        FL.job[fj].row_begin = (int)(fj * B);
        FL.njobs = fj + 1;
        FL.total_rows = (int)((fj + 1) * B);
        FL.block_partials = P.nce.block_partials;
        FL.ticket = P.nce.ticket;
        FL.out_loss = a->losses;   // [0] = itc, [1] = pt
        if ((rc = launch_finish(FL, s_loss))) return rc;
    }
    // prototype partial sums (+ in-place accumulate) from the teacher features (STiLModel.py:376-381)
    if ((rc = launch_proto_accumulate(a->feat_m_e, dt, B, D, D, P.cls, P.conf, B_l, a->repeat_ratio, K, a->class_sum,
                                      a->class_count, a->prototypes_sum, a->prototypes_count_sum, s_acc)))
        return rc;
    if (a->partials_push) {
        // data-parallel head: the packed [class_sum | class_count] partial leaves for every rank as soon as it exists,
        // mid-step on this side stream, instead of after the whole row-local chain
        const stil_p2p_channel* ch = a->partials_push;
        STIL_REQUIRE(a->class_count == a->class_sum + K * D, STIL_E_ARG,
                     "head_step: partials_push needs class_count right behind class_sum (one packed partial)");
        const void* src = a->class_sum;
        const int64_t nbytes = round_up((K * D + K) * (int64_t)sizeof(float), 16), dst = a->partials_dst_offset;
        if ((rc = stil_p2p_push(ch->bases, ch->world, ch->rank, ch->flags_offset, ch->ctrl_offset, ch->channel, 1, &src,
                                &nbytes, &dst, s_acc)))
            return rc;
    }
    // masked soft-target CE of the student heads (f-1)
    if (a->y_m) {
        STIL_REQUIRE(a->y_i && a->y_t && a->mask_random, STIL_E_ARG, "head_step: student logits need y_i, y_t, mask_random");
        const int lsz = a->logit_dtype == STIL_BF16 ? 2 : 4;
        auto urow = [&](const void* p) { return static_cast<const char*>(p) + B_l * K * lsz; };
        auto grow = [&](float* p) { return p ? p + B_l * K : nullptr; };
        if ((rc = launch_masked_softce(urow(a->y_m), urow(a->y_i), urow(a->y_t), a->logit_dtype, K, a->pseudo_label, K,
                                       a->mask1, a->case1, a->case2_i, a->case2_t, a->case3, a->mask_random, B_u, K,
                                       a->losses + 2, grow(a->d_y_m), grow(a->d_y_i), grow(a->d_y_t), K,
                                       a->rate_uce_scale, P.ce_partials, P.ce_ticket, s_ce)))
            return rc;
    }

    // ---- join
    for (int i = 0; i < kSide; ++i) {
        if (SS->s[i] == s_nce && !nce) continue;   // that stream never joined this step (nothing was forked to it)
        if (SS->s[i] == s_ce && !a->y_m) continue;
        STIL_CUDA(cudaEventRecord(SS->join[i], SS->s[i]));
        STIL_CUDA(cudaStreamWaitEvent(st, SS->join[i], 0));
    }
    return STIL_OK;
}

}  // extern "C"
