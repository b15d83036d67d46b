/* elimrec_b200 - C-ABI of the B200-native EliMRec hot path (libelimrec_b200.so).
 *
 * Drop-in boundary.  The reference has no FFI for its model math (it calls torch
 * ops from Python) and three Cython->C++ entry points for evaluation / sampling; every function
 * below names the reference interface it replaces.  Host side stays Python (the reference's host
 * language) and binds these with ctypes - see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.  All `const float*`/`float*`/`int*`
 *     arguments are DEVICE pointers unless the name ends in `_host`.
 *   - every launch goes to `stream` (a cudaStream_t passed as void*); no host synchronisation, no
 *     allocation, no ownership transfer: the caller (torch) allocates inputs, outputs, workspaces.
 *   - return 0 on success, <0 on error; elimrec_last_error() returns a thread-local message.
 *   - re-entrant; no hidden global state.  Never throws across the boundary.
 *   - embedding width is fixed at ELIMREC_D = 64 (`recdim=64`, conf/EliMRec.properties:6).
 */
#ifndef ELIMREC_B200_H
#define ELIMREC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ELIMREC_D 64
#define ELIMREC_MAX_LAYERS 8
#define ELIMREC_MAX_MODS 3

typedef void* elimrec_stream_t; /* cudaStream_t */

const char* elimrec_last_error(void);
int elimrec_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * spmm - replaces torch.sparse.mm(norm_adj, all_emb) (models/EliMRec.py:244) and its
 * SparseAddmmBackward (A is symmetric for adj_type='pre', so backward = forward on the gradient).
 *
 * The normalised adjacency is kept as two CSR halves (user rows -> item cols, item rows -> user cols)
 * in "segment" form built once on the host (elimrec_b200/graph.py): seg[s] = {row, edge_begin,
 * edge_end, heavy_id}.  Rows longer than the segment length are split; their segments come first in
 * the list, padded per row to a multiple of 8 (one CTA = 8 segments of ONE row), each CTA reduces its 8
 * warps in shared memory and writes one partial to `partial[cta]`; the last-arriving CTA of the row adds
 * them in a fixed order (deterministic).  heavy[h] = {first_cta, n_ctas}; `counter[h]` must be zero
 * on entry and is left zero.  `partial` holds (n_heavy_segments / 8) * width floats.
 *
 * Y[row, 0:width] = sum_e val[e] * X[col[e], 0:width]           width in {64, 128, 256}
 * Split rows and whole rows are two launches (different register budgets); `part` lets the caller put them on
 * two streams - they write disjoint rows.
 *
 * Optional fused layer-mean epilogue (models/EliMRec.py:246-247, torch.stack + torch.mean):
 *   if mean_out != NULL this launch is the LAST propagation layer; instead of (or in addition to)
 *   storing Y it writes   mean_out[row, 0:mean_width] = (sum_k layer_k[row] + Y[row]) * mean_scale
 *   where layer_k (k < n_prev) are the previous layers' rows, `prev_width[k]` wide (64 -> broadcast to
 *   every 64-column block), summed in order k = 0..n_prev-1 exactly like the reference.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t n_prev;
    const float* prev[ELIMREC_MAX_LAYERS];
    int64_t prev_ld[ELIMREC_MAX_LAYERS];
    int32_t prev_width[ELIMREC_MAX_LAYERS];
    float* mean_out;
    int64_t mean_ld;
    int32_t mean_width;
    float mean_scale;
} elimrec_mean_epilogue_t;

int elimrec_spmm(int width, int part /* 0 = all rows, 1 = split rows only, 2 = whole rows only */, int n_seg,
                 int n_heavy_seg, const int32_t* seg, const int32_t* heavy, int32_t* counter,
                 const int32_t* col, const float* val, const float* X, int64_t ldx, float* Y /* may be NULL */,
                 int64_t ldy, float* partial, const elimrec_mean_epilogue_t* epi /* may be NULL */,
                 elimrec_stream_t stream);

/* Row-sparse variant for the LAST propagation layer of a training step.  The BPR loss reads the layer-mean output
 * only at the <= 3B sampled rows (models/EliMRec.py:115-142 indexes all_users / all_items with the batch), so
 *   forward : row_mask[row] == 0  -> the row is skipped, its output left untouched (only sampled rows are produced)
 *   backward: col_mask[col] == 0  -> the edge is dropped before the gather (d x_L is zero outside the sampled rows)
 * Masks are byte arrays over the rows / columns of this CSR half; either may be NULL.  Surviving edges keep their
 * order: rows are bit-identical to elimrec_spmm on the same inputs (width 64 + col_mask: the two half-warps of a warp
 * split the surviving edges differently, so those sums agree to fp32 rounding instead).
 * Additive epilogue (addend may be NULL): Y[row, :] += addend[row, 0:width] for rows with add_mask[row] != 0 (add_mask NULL:
 * every row) - the layer-mean gradient G that enters every layer of the backward chain, d x_{k-1} = A^T d x_k + G
 * (torch.mean over the stacked layers, models/EliMRec.py:246-247); G is only valid on the sampled rows add_mask marks. */
int elimrec_spmm_masked(int width, int part, int n_seg, int n_heavy_seg, const int32_t* seg, const int32_t* heavy,
                        int32_t* counter, const int32_t* col, const float* val, const float* X, int64_t ldx, float* Y,
                        int64_t ldy, float* partial, const elimrec_mean_epilogue_t* epi, const uint8_t* row_mask,
                        const uint8_t* col_mask, int row_density_pct /* expected % of marked rows: scheduling hint only */,
                        const float* addend, int64_t ld_add, const uint8_t* add_mask, elimrec_stream_t stream);
/* The 64-wide propagation of BOTH CSR halves in one launch (csrc/spmm64.cu): an 8-lane group owns a work item, four items
 * per warp.  Each half: Y[row, 0:64] = sum_e val[e] * X[col[e], 0:64].  Work items (elimrec_b200/graph.py build_segments64):
 * item[i] = {row, edge_begin, edge_end, split_row_id or -1}; rows longer than 64 edges are dealt evenly over several items
 * that come first in the list (n_split_item of them), write partial sums to `partial` ([n_split_item x 64] floats) and the
 * last-arriving item of split row h (counter[h], zero on entry and on exit; split_rows[h] = {first item, number of items})
 * adds them in item order - deterministic.  Whole rows follow by descending degree.
 * row_mask / col_mask as in elimrec_spmm_masked (either may be NULL); masked rows carry the bits of the dense launch.
 * `b` may be NULL.  variant: 0 = default tuning; 1..4 = other unroll / occupancy points (tools/spmm64_bench.py). */
typedef struct {
    int32_t n_item, n_split_item;
    const int32_t* item;
    const int32_t* split_rows;
    int32_t* counter;
    float* partial;
    const int32_t* col;
    const float* val;
    const float* X;
    int64_t ldx;
    float* Y;
    int64_t ldy;
    const uint8_t* row_mask;
    const uint8_t* col_mask;
    const float* addend;      /* may be NULL: Y[row] += addend[row, 0:64] where add_mask[row] != 0 (add_mask NULL: every row) */
    int64_t ld_add;
    const uint8_t* add_mask;
    /* Fused Adam (adam_param != NULL; Y may then be NULL): the finished row is d loss / d table[row] - the last hop of the
     * backward chain - and the [rows x 64] table, exp_avg and exp_avg_sq rows are updated in place instead of storing the
     * gradient (torch.optim.Adam arithmetic, main.py:49,101).  The propagation is bound by L2->SM gathers and leaves HBM
     * idle; the optimizer is pure HBM streaming - fused, its traffic rides along.  adam_old_out (may be NULL) receives the
     * parameter row as it was BEFORE the update: what tables completed later for predict() must use. */
    float* adam_param;
    float* adam_exp_avg;
    float* adam_exp_avg_sq;
    float* adam_old_out;
} elimrec_spmm64_half_t;
typedef struct {
    const double* consts_dev; /* {lr / bias_correction1, sqrt(bias_correction2)} written by elimrec_adam_tick */
    double beta1, beta2;
    float eps, weight_decay;
} elimrec_adam_consts_t;
/* width: 64 (one GPU) or 32 / 16 / 8 (column-sharded multi-GPU mode: a rank propagates 64 / world columns of every row; X, Y,
 * addend, partial and the Adam tensors are then `width` columns wide) */
int elimrec_spmm64_pair(const elimrec_spmm64_half_t* a, const elimrec_spmm64_half_t* b, const elimrec_adam_consts_t* adam /* may be NULL */,
                        int width, int variant, elimrec_stream_t stream);
/* mask[0:n_nodes] = 0; mask[rows[r]] = 1 */
int elimrec_mark_rows(int n_rows, const int32_t* rows, int64_t n_nodes, uint8_t* mask, elimrec_stream_t stream);
/* rows[0:3B] = [users | num_users + pos | num_users + neg]  (node ids of the batch, models/EliMRec.py:120-122 gathers
 * exactly these rows); mask, mask2 (each may be NULL): m[0:n_nodes] = 0 then m[rows] = 1 */
int elimrec_inst_rows(int B, const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t num_users,
                      int32_t* rows, int64_t n_nodes, uint8_t* mask, uint8_t* mask2, elimrec_stream_t stream);
/* out_mask[col] = 1 for every edge (row, col) of this CSR half with row_mask[row] != 0: the rows of the other side
 * that the marked rows gather (two-hop support of the batch; seg = the same segment list elimrec_spmm takes) */
int elimrec_mark_neighbors(int n_seg, const int32_t* seg, const int32_t* col, const uint8_t* row_mask, uint8_t* out_mask,
                           elimrec_stream_t stream);
/* dst[rows[r] - row_offset, 0:width] = 0 for row_lo <= rows[r] < row_hi */
int elimrec_zero_rows(int n_rows, const int32_t* rows, int32_t row_lo, int32_t row_hi, int32_t row_offset, float* dst,
                      int64_t dst_ld, int width, elimrec_stream_t stream);

/* rows[r] selects a destination row; dst[rows[r], 0:width] += scale * src[r, 0:src_width] (atomic).
 * src_width == width: plain; src_width == fold*width: the `fold` 64-column blocks are summed first
 * (gradient of the 64->wide broadcast).  Used to seed / add the row-sparse layer-mean gradient. */
int elimrec_scatter_add_rows(int n_rows, const int32_t* rows, int32_t row_lo, int32_t row_hi, int32_t row_offset,
                             const float* src, int64_t src_ld, int src_width, float* dst, int64_t dst_ld, int width,
                             float scale, elimrec_stream_t stream);
/* dst[r, :] = src[rows[r], :] */
int elimrec_gather_rows(int n_rows, const int32_t* rows, const float* src, int64_t src_ld, float* dst, int64_t dst_ld,
                        int width, elimrec_stream_t stream);
/* out[r, g*64 + c] = src[r, c] for g < n_rep   (layer-0 user rows feed every modality graph) */
int elimrec_broadcast_cols(int64_t n_rows, const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int n_rep,
                           elimrec_stream_t stream);
int elimrec_copy_2d(int64_t n_rows, int width, const float* src, int64_t src_ld, float* dst, int64_t dst_ld,
                    elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * column-block helpers of the non-default model variants (csrc/blocks.cu)
 * ------------------------------------------------------------------------------------------------ */
/* dst[r, g*64 + c] = scale * src[r, c], g < n_rep.  mm_fusion_mode='mean' (models/EliMRec.py:224-225):
 * W.mean(stack(reps)) == [W/G | ... | W/G].cat(reps), so the mean fusion runs on the concat kernels with a tied weight. */
int elimrec_tie_blocks(int64_t n_rows, const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int n_rep, float scale,
                       elimrec_stream_t stream);
/* dst[r, c] = scale * sum_{g < n_rep} src[r, g*64 + c]   (gradient of a tied weight; d E_u summed over the graphs) */
int elimrec_fold_blocks(int64_t n_rows, const float* src, int64_t src_ld, int n_rep, float scale, float* dst, int64_t dst_ld,
                        elimrec_stream_t stream);
/* Y[r, :] += row_scale[r] * X[r, :] - the diagonal of adj_type 'norm' / 'mean' (models/EliMRec.py:332-334,349-352),
 * applied after the bipartite SpMM of the layer */
int elimrec_axpy_rows(int64_t n_rows, int width, const float* row_scale, const float* X, int64_t ldx, float* Y, int64_t ldy,
                      elimrec_stream_t stream);
/* out = scale * (((x_0 + x_1) + x_2) + ...) over n_layers slabs (torch.stack + torch.mean, models/EliMRec.py:246-247);
 * used when the mean cannot be fused into the last SpMM.  layers_host / ld_host are HOST arrays of device pointers / strides. */
int elimrec_layer_mean(int64_t n_rows, int width, int n_layers, const float* const* layers_host, const int64_t* ld_host,
                       float scale, float* out, int64_t ld_out, elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * linear schedule (csrc/linsched.cu) - compute_graph x4 + mm_fusion inputs (models/EliMRec.py:238-258) by linearity.
 *
 * light_out_m = mean_k A_hat^k [E_u ; X_m W_m^T + b_m] = mean_k A_hat^k [E_u ; 0] + Zbar_m [W_m | b_m]^T, where
 * Zbar_m = mean_k A_hat^k [0 ; X_m | 1] is a CONSTANT of the dataset (item features are never trained), and the first
 * term is read off the id graph's own layers p_k = A_hat^k [E_u ; E_i]: A_hat is bipartite, so it sits on the user rows
 * of the even layers and on the item rows of the odd ones.  One 64-wide propagation per step instead of one 64-wide plus
 * three 256-wide ones, and no pass over the features.
 *
 * elimrec_lin_assemble: for output row j (node = rows[j], or j when rows == NULL; users first):
 *   out[j, 0:64]                 = scale * ((p_0 + p_1) + ... + p_L)[node]            (light_out of the id graph)
 *   out[j, 64(1+m) : 64(2+m)]  (+)= scale * sum_{k : k even (user) / odd (item)} p_k[node]    for m < n_mod
 * accumulate != 0: the modality blocks already hold Zbar_m[node] W'_m^T (written by the GEMM) and are added to.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t n;                                       /* L + 1 tables p_0 .. p_L */
    const float* user[ELIMREC_MAX_LAYERS + 1];       /* p_k, user rows [U x 64] */
    const float* item[ELIMREC_MAX_LAYERS + 1];       /* p_k, item rows [I x 64] */
    int64_t user_ld[ELIMREC_MAX_LAYERS + 1];
    int64_t item_ld[ELIMREC_MAX_LAYERS + 1];
} elimrec_lin_layers_t;
int elimrec_lin_assemble(int64_t n_rows, const int32_t* rows /* may be NULL */, int32_t num_users,
                         const elimrec_lin_layers_t* layers, float scale, int n_mod, int accumulate, float* out, int64_t ldo,
                         elimrec_stream_t stream);
/* adjoint of elimrec_lin_assemble w.r.t. layer `layer` of the chain, on the instance rows (atomic: nodes repeat):
 *   dst[rows[j], 0:64] += scale * (dO[j, 0:64] + [layer even (user row) / odd (item row)] * sum_m dO[j, 64(1+m):64(2+m)]) */
int elimrec_lin_seed(int n_rows, const int32_t* rows, int32_t num_users, int layer, const float* dO, int64_t ldo, int n_mod,
                     float scale, float* dst /* [N x 64], users first */, int64_t ldd, elimrec_stream_t stream);
/* both seed vectors at once, for the fused form of the chain (the propagation launch adds them in its epilogue):
 *   GA[rows[j], 0:64] += scale * sum_{all blocks} dO[j];   GB[rows[j], 0:64] += scale * dO[j, 0:64]   (atomic)
 * layer k adds GA on user rows for even k / item rows for odd k, and GB on the other side */
int elimrec_lin_seed2(int n_rows, const int32_t* rows, const float* dO, int64_t ldo, int n_mod, float scale, float* GA, float* GB,
                      int64_t ldg, elimrec_stream_t stream);
/* dst_m [64 x Kp] = [ W_m [64 x Dm] | b_m | 0 .. ]: the projection weight with its bias as one more input column (the
 * matching column of Zbar_m is mean_k A_hat^k [0 ; 1]); round_tf32 != 0 rounds to nearest TF32 for the tensor-core GEMM */
typedef struct {
    const float* W;
    const float* b;
    float* dst;
    int64_t Dm, Kp;
} elimrec_pack_proj_t;
int elimrec_pack_proj_weights(int n, const elimrec_pack_proj_t* tensors_host, int round_tf32, elimrec_stream_t stream);
/* dst [3 n_rows x width] = [hi ; hi ; lo] (pattern 0) or [hi ; lo ; hi] (pattern 1) of src [n_rows x width], hi = rna_tf32(x),
 * lo = rna_tf32(x - hi): elimrec_linear_tf32_wgrad over the 3 n_rows stacked rows of a pattern-0 dY and a pattern-1 X computes
 * hi*hi + hi*lo + lo*hi, i.e. the weight gradient in the fp32 accuracy class on the TF32 tensor-core kernel */
int elimrec_split3_rows(int64_t n_rows, int width, const float* src, int64_t lds, float* dst, int64_t ldd, int pattern,
                        elimrec_stream_t stream);
/* Y[r, 0:width] = [Y +] scale * X[r, 0:width]  (accumulating the constant Zbar tables, once per model) */
int elimrec_axpy_2d(int64_t n_rows, int width, float scale, const float* X, int64_t ldx, float* Y, int64_t ldy, int accumulate,
                    elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * gemm - replaces nn.Linear forward/backward for v_dense/a_dense/t_dense (models/EliMRec.py:233-236),
 * embedding_{user,item}_after_GCN (:261-270) and s_dense_* (:146-151).
 *
 * C(m,n) = [C(m,n) +] sum_k A(m,k) * B(k,n) [+ bias[n]]   with fully general element strides, so that
 * transposes and column-block views need no copies.  split_k > 1 writes per-slice partials to
 * `workspace` ([split_k x M x N] floats) and reduces them in slice order (deterministic).
 * precision: 0 = fp32 FFMA.  (tensor-core paths: see elimrec_linear_tf32_* below.)
 * ------------------------------------------------------------------------------------------------ */
int elimrec_gemm(int64_t M, int64_t N, int64_t K, const float* A, int64_t a_sm, int64_t a_sk, const float* B,
                 int64_t b_sk, int64_t b_sn, float* C, int64_t c_sm, int64_t c_sn, const float* bias, int accumulate,
                 int split_k, float* workspace, const float* scale_dev /* may be NULL */, elimrec_stream_t stream);
int64_t elimrec_gemm_workspace_floats(int64_t M, int64_t N, int split_k);
/* out[n] = [out[n] +] sum_m A[m*ld + n] * (*scale_dev)  (bias gradients), deterministic */
int elimrec_colsum(int64_t M, int64_t N, const float* A, int64_t ld, float* out, float* workspace, int accumulate,
                   const float* scale_dev, elimrec_stream_t stream);
int64_t elimrec_colsum_workspace_floats(int64_t M, int64_t N);

/* Tensor-core linear layers (tcgen05, TF32 inputs, fp32 accumulate in TMEM):
 *   fwd : Y[M x 64] (ldy) = X[M x K] * W[64 x K]^T + b          (K % 32 == 0)
 *   wgrad: dW[64 x K] = dY[M x 64]^T (lddy) * X[M x K]            split over M, deterministic reduce
 * return -2 if the shape is unsupported (caller falls back to elimrec_gemm). */
int elimrec_linear_tf32_fwd(int64_t M, int64_t K, const float* X, int64_t ldx, const float* W, const float* b,
                            float* Y, int64_t ldy, elimrec_stream_t stream);
/* Several such layers in ONE persistent launch (grid <= SM count; the three modal projections of a step,
 * models/EliMRec.py:233-236): problem i computes Y_i[M_i, 0:64] = X_i[M_i, K_i] W_i[64, K_i]^T + b_i. */
typedef struct {
    int64_t M, K;
    const float* X;
    int64_t ldx;
    const float* W;
    const float* b; /* may be NULL */
    float* Y;
    int64_t ldy;
} elimrec_linear_desc_t;
int elimrec_linear_tf32_fwd_multi(int n /* <= 4 */, const elimrec_linear_desc_t* problems, elimrec_stream_t stream);
int elimrec_linear_tf32_wgrad(int64_t M, int64_t K, const float* dY, int64_t lddy, const float* X, int64_t ldx,
                              float* dW, float* workspace, elimrec_stream_t stream);
int64_t elimrec_linear_tf32_wgrad_workspace_floats(int64_t M, int64_t K);
/* 3xTF32 forward: same contract as elimrec_linear_tf32_fwd but fp32-class accuracy (error ~1e-6 relative):
 * X is split into TF32 hi/lo parts inside the kernel, W_hi / W_lo come pre-split from elimrec_split_tf32, and
 * X W^T = X_hi W_hi + X_hi W_lo + X_lo W_hi is accumulated in TMEM.  Used for the fusion Linear and the heads. */
int elimrec_linear_x3_fwd(int64_t M, int64_t K, const float* X, int64_t ldx, const float* W_hi, const float* W_lo,
                          const float* b, float* Y, int64_t ldy, elimrec_stream_t stream);
int elimrec_split_tf32(int64_t n, const float* src, float* hi, float* lo, elimrec_stream_t stream);
/* One pass over `rows` rows of the layer-mean slab O [rows x 64(1+n_heads)] (row stride ldo) producing, with 3xTF32:
 *   F_out[r]    = O[r, :] @ Wf^T + bf                          (fusion Linear, models/EliMRec.py:261-270)
 *   S_out[m][r] = O[r, 64(m+1) .. 64(m+2)) @ Ws[m]^T + bs[m]    (single-modal heads, models/EliMRec.py:146-151)
 * Wf_* are [64 x 64(1+n_heads)], Ws_* [64 x 64] hi/lo parts from elimrec_prep_weights_tf32; outputs have row stride 64. */
int elimrec_fuse_heads_x3(int64_t rows, int n_heads, const float* O, int64_t ldo, const float* Wf_hi, const float* Wf_lo,
                          const float* bf, const float* const* Ws_hi_host, const float* const* Ws_lo_host,
                          const float* const* bs_host, float* F_out, float* const* S_out_host, elimrec_stream_t stream);
/* Same, persistent, over the WHOLE slab O [(U+I) x 64(1+n_heads)] (users first): user rows use Wu/bu, item rows Wi/bi;
 * F_out and S_out[m] are [(U+I) x 64].  One CTA per SM, TMEM double-buffered accumulators, epilogue overlapped. */
int elimrec_fuse_heads_x3_all(int64_t num_users, int64_t num_items, int n_heads, const float* O, int64_t ldo,
                              const float* Wu_hi, const float* Wu_lo, const float* bu, const float* Wi_hi,
                              const float* Wi_lo, const float* bi, const float* const* Ws_hi_host,
                              const float* const* Ws_lo_host, const float* const* bs_host, float* F_out,
                              float* const* S_out_host, elimrec_stream_t stream);
/* several small weight tensors prepared for the tensor cores in ONE launch: hi = rna_tf32(src); if lo != NULL,
 * lo = rna_tf32(src - hi) (3xTF32 split), else round only */
#define ELIMREC_PREP_MAX 16
typedef struct {
    const float* src;
    float* hi;
    float* lo;
    int64_t numel;
} elimrec_prep_tensor_t;
int elimrec_prep_weights_tf32(int n, const elimrec_prep_tensor_t* tensors_host, elimrec_stream_t stream);
/* dst[i] = round-to-nearest TF32 of src[i] (tcgen05.mma kind::tf32 truncates its inputs; pre-rounding the constant
 * features once and the weights per step removes the truncation bias).  src == dst allowed. */
int elimrec_round_tf32(int64_t n, const float* src, float* dst, elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * bpr - replaces getEmbedding gathers + original_bpr_loss x (1+M) + their autograd
 * (models/EliMRec.py:115-142, 277-297).
 *
 * tables[t] : [N x 64] fp32, users first then items (t = 0 fused table, t >= 1 single-modal heads)
 * weight[t] : 1 for t = 0, alpha (or 0 if the modality is ablated) for t >= 1
 * users/pos/neg : int64 [B] (what main.py:94-96 passes)
 * loss_out  : 1 float.   inst_rows : int32 [3B] node ids (u | U+pos | U+neg)
 * inst_grad : [3B x 64*n_tables]  d loss / d tables[t][inst_rows[r]]  (t-th 64-column block)
 * workspace : n_tables * B floats
 * ------------------------------------------------------------------------------------------------ */
int elimrec_bpr_forward_backward(int B, int n_tables, const float* const* tables_host, const float* weight_host,
                                 const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t num_users,
                                 float* loss_out, int32_t* inst_rows, float* inst_grad, float* workspace,
                                 elimrec_stream_t stream);

/* The same in parts: 1 = per-triple terms + instance gradients (what the backward waits for), 2 = the fixed-order reduction of
 * the terms to the loss scalar (nothing but the caller's read of the loss waits for it), 3 = both. */
int elimrec_bpr_forward_backward_part(int part, int B, int n_tables, const float* const* tables_host, const float* weight_host,
                                      const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t num_users,
                                      float* loss_out, int32_t* inst_rows, float* inst_grad, float* workspace,
                                      elimrec_stream_t stream);

/* Backward of embedding_{user,item}_after_GCN and s_dense_* (models/EliMRec.py:261-270,146-151) restricted to
 * the 3B instance rows, where alone their gradient is non-zero:
 *   dO_inst [3B x F]   = g * (dF @ W_{u|i} + per-block dS_m @ Ws_m)        (rows < B use Wu, the rest Wi)
 *   dWu/dWi [64 x F], dbu/dbi [64], dWs[m] [64 x 64], dbs[m] [64]             (deterministic chunked reduction)
 * O_inst [3B x F] = rows of the layer-mean slab gathered at inst_rows; F = 64 * n_tables. */
int64_t elimrec_inst_backward_workspace_floats(int B, int n_tables, int F);
int elimrec_inst_backward(int B, int n_tables, int F, const float* inst_grad, const float* O_inst,
                          const float* gscale_dev /* may be NULL */, const float* Wu, const float* Wi,
                          const float* const* Ws_host, float* dO_inst, float* dWu, float* dWi, float* dbu, float* dbi,
                          float* const* dWs_host, float* const* dbs_host, float* workspace, elimrec_stream_t stream);
/* d O[inst] (part 1 of elimrec_inst_backward) with the backward seeds of the linear schedule in its epilogue - what
 * elimrec_lin_seed2 does in a launch of its own: GA[rows[j]] += scale * (sum of the 1 + n_mod 64-column blocks of dO_inst[j]),
 * GB[rows[j]] += scale * dO_inst[j, 0:64]  (float atomics; rows = the 3B instance nodes).  F = 256 only. */
int elimrec_inst_dout_seed(int B, int n_tables, int F, const float* inst_grad, const float* gscale_dev /* may be NULL */,
                         const float* Wu, const float* Wi, const float* const* Ws_host, float* dO_inst, const int32_t* rows,
                         int n_mod, float scale, float* GA, float* GB, int64_t ldg, elimrec_stream_t stream);

/* Forward of the same two layers on the 3B instance rows of a row-sparse step, exact fp32, one launch (csrc/bpr.cu):
 *   F_out[r] = O_inst[r, 0:F] @ W_{u|i}^T + b_{u|i}  (rows 0..B-1 take the user fusion Linear, B..3B-1 the item one;
 *   models/EliMRec.py:261-270),  S_out[m][r] = O_inst[r, 64(m+1) : 64(m+2)] @ Ws[m]^T + bs[m]  (models/EliMRec.py:146-151).
 * F = 64 * n_tables; every pointer 16-byte aligned. */
int elimrec_inst_forward(int B, int n_tables, int F, const float* O_inst, const float* Wu, const float* Wi,
                         const float* const* Ws_host, const float* bu, const float* bi, const float* const* bs_host,
                         float* F_out, float* const* S_out_host, elimrec_stream_t stream);
/* The same in two independent parts, so that the weight gradients (part 2: they only feed Adam) can run on another
 * stream while d O[inst] (part 1) seeds the backward propagation.  part 3 = both, in this order. */
int elimrec_inst_backward_part(int part, int B, int n_tables, int F, const float* inst_grad, const float* O_inst,
                          const float* gscale_dev /* may be NULL */, const float* Wu, const float* Wi,
                          const float* const* Ws_host, float* dO_inst, float* dWu, float* dWi, float* dbu, float* dbi,
                          float* const* dWs_host, float* const* dbs_host, float* workspace, elimrec_stream_t stream);

/* Every weight / bias gradient that contracts over the 3B instance rows in ONE launch (+ one fixed-order reduction),
 * exact fp32 (csrc/wgrad_multi.cu):   out_p [64 x K_p] (row stride ldo) = g * A_p[r0:r1, 0:64]^T B_p[r0:r1, 0:K_p],
 * bias_out_p [64] = g * column sums of A_p[r0:r1] (if not NULL);  g = *gscale_dev for problems with scale_by_g, else 1.
 * Fusion Linears: A = instance gradient block 0, B = O[inst rows]; heads: block m of both; modality projections of the
 * linear schedule: A = block m of dO[inst rows], B = Zbar_m[inst rows] (its ones column yields the bias gradient). */
#define ELIMREC_WGRAD_MAX_PROBLEMS 12
typedef struct {
    const float* A;      /* [rows x >= 64], 16-byte aligned, lda % 4 == 0 */
    int64_t lda;
    const float* B;      /* [rows x >= K] */
    int64_t ldb;
    int64_t K;
    int64_t row_begin, row_end;
    float* out;
    int64_t ldo;
    float* bias_out;     /* may be NULL */
    int32_t scale_by_g;
} elimrec_wgrad_problem_t;
int64_t elimrec_wgrad_multi_workspace_floats(int n, const elimrec_wgrad_problem_t* problems_host, int splits);
int elimrec_wgrad_multi(int n, const elimrec_wgrad_problem_t* problems_host, int splits /* row ranges per tile, <= 64 */,
                        float* workspace, const float* gscale_dev /* may be NULL */, elimrec_stream_t stream);
/* The same contract on the tensor cores (csrc/linear_tc.cu): tcgen05 kind::tf32 on hi / lo split operands (3xTF32: lo*hi +
 * hi*lo + hi*hi in one TMEM accumulator, the bias sums as two more MMAs against an all-ones tile), fp32-class accuracy
 * (relative error ~1e-6 of sum |a||b|) instead of exact FFMA order; same workspace size, same fixed-order reduction of the row-range
 * partials.  B and ldb must additionally be 16-byte aligned / a multiple of 4. */
int elimrec_wgrad_multi_x3(int n, const elimrec_wgrad_problem_t* problems_host, int splits, float* workspace,
                           const float* gscale_dev /* may be NULL */, elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * adam - replaces torch.optim.Adam(lr, weight_decay) .step() (main.py:49,101): coupled L2,
 * betas (0.9, 0.999), eps 1e-8, bias correction from the device-resident step counter.
 * elimrec_adam_tick increments *step_dev and refreshes consts_dev = {lr/bc1, sqrt(bc2)} (2 doubles);
 * elimrec_adam_apply updates one tensor; grad may be a strided view (row_len, grad_ld).
 * ------------------------------------------------------------------------------------------------ */
int elimrec_adam_tick(int64_t* step_dev, double* consts_dev, double lr, double beta1, double beta2,
                      elimrec_stream_t stream);
#define ELIMREC_ADAM_MAX_TENSORS 24
typedef struct {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t numel, row_len, grad_ld;
} elimrec_adam_tensor_t;
/* all parameter tensors in ONE launch (torch's _multi_tensor_adam); descriptors are host memory, copied by value */
int elimrec_adam_apply_multi(int n_tensors, const elimrec_adam_tensor_t* tensors_host, const double* consts_dev,
                             double beta1, double beta2, float eps, float weight_decay, elimrec_stream_t stream);
int elimrec_adam_apply(int64_t n, float* param, const float* grad, int64_t row_len, int64_t grad_ld, float* exp_avg,
                       float* exp_avg_sq, const double* consts_dev, double beta1, double beta2, float eps,
                       float weight_decay, elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * sampler - replaces _pairwise_sampling_v2 + randint_choice (data/sampler.py:93-126,
 * util/cython/random_choice.pyx:12-62).
 *
 * compat (HOST, sequential by construction): bit-exact replay of the reference's libc rand()
 * stream.  `rng_state_host` is 40 uint32 words owned by the caller (seed with *_seed; glibc TYPE_3
 * additive-feedback generator re-implemented so that it is re-entrant and independent of libc).
 * device: counter-based Philox4x32-10, same distribution, one thread per triple.
 * ------------------------------------------------------------------------------------------------ */
void elimrec_compat_rng_seed(uint32_t* rng_state_host, uint32_t seed);
uint32_t elimrec_compat_rng_next(uint32_t* rng_state_host);
int elimrec_sample_epoch_compat(uint32_t* rng_state_host, int32_t n_train_users, const int32_t* user_ids_host,
                                const int64_t* row_ptr_host, const int32_t* items_host, int32_t num_items,
                                int64_t num_samples, int64_t* out_users_host, int64_t* out_pos_host,
                                int64_t* out_neg_host);
int elimrec_sample_triples_device(uint64_t seed, uint64_t epoch, int64_t num_samples, int32_t n_train_users,
                                  const int32_t* user_ids, const int64_t* row_ptr, const int32_t* items,
                                  int32_t num_items, int64_t* out_users, int64_t* out_pos, int64_t* out_neg,
                                  elimrec_stream_t stream);

/* One BATCH of the same stream from inside a CUDA graph: triple t of the launch is sample (*batch_index_dev) * batch_size + t
 * of epoch `epoch` (identical to the corresponding slice of elimrec_sample_triples_device).  batch_index_dev is a device
 * counter the caller advances between replays - the optimizer's step counter (elimrec_adam_tick) does it for free. */
int elimrec_sample_batch_device(uint64_t seed, uint64_t epoch, const int64_t* batch_index_dev, int64_t batch_size,
                                int32_t n_train_users, const int32_t* user_ids, const int64_t* row_ptr, const int32_t* items,
                                int32_t num_items, int64_t* out_users, int64_t* out_pos, int64_t* out_neg,
                                elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * rank - replaces EliMRec.predict + general_cm_fusion (models/EliMRec.py:96-113,155-212), the
 * train-item masking loop (uni_evaluator.py:149-154) and cpp_evaluate_matrix / metric.h
 * (evaluator/backend/cpp/include/evaluate.h:45-64, metric.h:17-114).
 *
 * mode = predict mode + 4 * score-fusion mode.  Predict mode: 0 = 'normal' sigmoid(sigmoid(u.i)), 1 = 'TE', 2 = 'TIE'.
 * Score fusion (s_fusion_mode, EliMRec.py:171-210): 0 = 'rubi' x * prod sigmoid(cos_m); 1 = 'hm'
 * log(z + 1e-12) - log1p(z), z = sigmoid(x) * prod sigmoid(cos_m); 2 = 'sum' log(sigmoid(x + sum cos_m) + 1e-12).
 * s_user/s_item: L2-row-normalised single-modal tables of the ACTIVE modalities (n_mod of them, reference order
 * v,a,t; hm / sum use every modality, as the reference does).  The tensor-core evaluator takes rubi only.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t num_users, num_items, n_mod, mode;
    const float* f_user; /* [U x 64] */
    const float* f_item; /* [I x 64] */
    const float* s_user[ELIMREC_MAX_MODS];
    const float* s_item[ELIMREC_MAX_MODS];
} elimrec_rank_tables_t;

/* out[r, :] = row / max(||row||, 1e-12)   (F.normalize, EliMRec.py:163-170) */
int elimrec_row_normalize(int64_t n_rows, const float* src, float* dst, elimrec_stream_t stream);
/* mean_i sigmoid(u.i) per evaluated user (EliMRec.py:107) */
int elimrec_rank_rowmean(const elimrec_rank_tables_t* t, int n_eval, const int32_t* eval_users, float* ui_mean,
                         elimrec_stream_t stream);
/* full score matrix [n_eval x I] (predict(); parity tests) */
int elimrec_rank_scores(const elimrec_rank_tables_t* t, int n_eval, const int32_t* eval_users, const float* ui_mean,
                        float* scores, elimrec_stream_t stream);
/* fused score + train mask + per-user top-K (ties: lowest index).  train CSR indexed by user id. */
int elimrec_rank_topk(const elimrec_rank_tables_t* t, int n_eval, const int32_t* eval_users, const float* ui_mean,
                      const int64_t* train_ptr, const int32_t* train_items, int K, int32_t* topk_idx, float* topk_val,
                      elimrec_stream_t stream);
/* Tensor-core evaluator (tcgen05 kind::f16, fp32-class accuracy through fp16 hi/lo operand pairs; csrc/rank_tc.cu).
 * Tables are pre-split with elimrec_split_fp16: value * scale = hi + lo (scale a power of two), [rows x 64] fp16.
 * table 0 = fused table, 1..n_mod = L2-normalised single-modal heads; inv_scale[t] = 1 / (user scale * item scale).
 * what = 0: mean_out[r] = mean_i sigmoid(u.i) (row mean for TIE); what = 1: fused score + train mask + top-K. */
typedef struct {
    int32_t num_users, num_items, n_mod, mode;
    const void* user_hi[1 + ELIMREC_MAX_MODS];
    const void* user_lo[1 + ELIMREC_MAX_MODS];
    const void* item_hi[1 + ELIMREC_MAX_MODS];
    const void* item_lo[1 + ELIMREC_MAX_MODS];
    float inv_scale[1 + ELIMREC_MAX_MODS];
} elimrec_rank_tc_tables_t;
int elimrec_split_fp16(int64_t n, const float* src, float scale, void* hi, void* lo, elimrec_stream_t stream);
int elimrec_rank_tc(const elimrec_rank_tc_tables_t* t, int what, int n_eval, const int32_t* eval_users, const float* ui_mean,
                    const int64_t* train_ptr, const int32_t* train_items, int K, int32_t* topk_idx, float* topk_val,
                    float* mean_out, void* workspace, elimrec_stream_t stream);
int64_t elimrec_rank_tc_workspace_bytes(int n_eval);
/* top-K of an explicit score matrix (arg_top_k_2d, util/cython/include/arg_topk.h:15-45), K <= 128; ties: lowest index.
 * With elimrec_rank_scores + elimrec_mask_train it is also the evaluator's path for 32 < K <= 128 (the fused rank kernels
 * keep one list entry per lane; UniEvaluator's default is top_k = 50, uni_evaluator.py:38). */
int elimrec_topk_matrix(int n_rows, int n_cols, const float* scores, int K, int32_t* topk_idx, float* topk_val,
                        elimrec_stream_t stream);
/* scores[r, train items of users[r]] = -inf (uni_evaluator.py:149-154); train CSR indexed by user id */
int elimrec_mask_train(int n_rows, int n_cols, const int32_t* users, const int64_t* train_ptr, const int32_t* train_items,
                       float* scores, elimrec_stream_t stream);
/* metric curves (K <= 128): rows[r, j*K + i] for metric_ids[j] in {1 Precision, 2 Recall, 3 MAP, 4 NDCG, 5 MRR}; truth CSR is
 * indexed by POSITION r (already gathered for the evaluated users), items sorted ascending.
 * inv_log2_host: K doubles 1/log2(i+2).  sums: [n_metrics*K] doubles, accumulated (zero it first). */
int elimrec_metric_rows(int n_eval, int K, const int32_t* topk_idx, const int64_t* truth_ptr,
                        const int32_t* truth_items, int n_metrics, const int32_t* metric_ids_host,
                        const double* inv_log2_host, float* rows, double* sums, elimrec_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * comm - NCCL collectives of the multi-GPU modes (csrc/comm.cu).  The reference is single-device
 * (device hard-wired to cuda:0, main.py:36): these have no reference counterpart, they are what shards its step.
 * One communicator per process / GPU, built from a unique id the caller distributes (rank 0: elimrec_comm_unique_id, then
 * e.g. one torch.distributed broadcast of the ELIMREC_COMM_ID_BYTES).  Every collective is enqueued on the caller's stream,
 * in order with the kernels around it; none synchronises the host; all are capturable in a CUDA graph.  NCCL is resolved
 * at run time from the libnccl.so.2 already loaded in the process.  Buffers are device pointers.
 *   allreduce : recv[i] = sum (or mean) over ranks of send[i]            (send == recv allowed)
 *   allgather : recv[q * bytes : (q+1) * bytes] = rank q's send[0 : bytes]
 *   alltoall  : recv[q * bytes : (q+1) * bytes] = rank q's send[me * bytes : (me+1) * bytes]
 * ------------------------------------------------------------------------------------------------ */
#define ELIMREC_COMM_ID_BYTES 128
/* Column-sharded ("feature-sharded") step, csrc/colshard.cu: rank r of `world` owns columns [r w, (r+1) w), w = 64 / world,
 * of both embedding tables and of every 64-wide slab of the linear schedule; propagation then needs no communication, and
 * only the instance rows of the batches cross NVLink (two all-to-alls per step).  Buffers [world][n][2w] are what
 * elimrec_comm_alltoall sends / receives (n = 3B rows of one rank's batch).
 *   cs_pack        out[j, 0:w] = scale * sum_k p_k[rows[j], 0:w], out[j, w:2w] = scale * parity sum   (layers: w-wide tables)
 *   cs_unpack      O[j, q w + c] = recv[q][j][c];  O[j, 64 (1+m) + q w + c] += recv[q][j][w + c], m < n_mod
 *   cs_seed_pack   send[q][j][c] = scale * sum_b dO[j, 64 b + q w + c];  send[q][j][w + c] = scale * dO[j, q w + c]
 *   cs_seed_scatter GA[rows[j], c] += recv[j][c];  GB[rows[j], c] += recv[j][w + c]      (atomic; rows = all world * n rows) */
/* triples [world][3][B] int64 (users | pos | neg of every rank's batch) -> rows [world][3][B] node ids (items offset by
 * num_users); mask / mask2 (each may be NULL): zeroed over n_nodes, then 1 at every instance row of every batch */
int elimrec_cs_inst_rows(int world, int B, const int64_t* triples, int32_t num_users, int32_t* rows, int64_t n_nodes, uint8_t* mask,
                         uint8_t* mask2, elimrec_stream_t stream);
int elimrec_cs_pack(int64_t n_rows, const int32_t* rows /* may be NULL: row j = node j */, int32_t num_users,
                    const elimrec_lin_layers_t* layers, float scale, int w, float* out, elimrec_stream_t stream);
int elimrec_cs_unpack(int world, int64_t n, int w, const float* recv, int n_mod, float* O, int64_t ldo, elimrec_stream_t stream);
int elimrec_cs_seed_pack(int64_t n, int world, int w, const float* dO, int64_t ldo, int n_mod, float scale, float* send,
                         elimrec_stream_t stream);
int elimrec_cs_seed_scatter(int64_t n_all, const int32_t* rows, int w, const float* recv, float* GA, float* GB, int64_t ldg,
                            elimrec_stream_t stream);
int elimrec_comm_unique_id(void* unique_id_out_host);
int elimrec_comm_init(const void* unique_id_host, int rank, int world, void** comm_out);
int elimrec_comm_destroy(void* comm);
int elimrec_comm_allreduce(void* comm, const float* send, float* recv, int64_t n, int average, elimrec_stream_t stream);
int elimrec_comm_allgather(void* comm, const void* send, void* recv, int64_t bytes_per_rank, elimrec_stream_t stream);
int elimrec_comm_alltoall(void* comm, const void* send, void* recv, int64_t bytes_per_pair, elimrec_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ELIMREC_B200_H */
