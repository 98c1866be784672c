/*
 * vilbert_b200.h — C ABI of libvilbert_b200.so: the sm_100a kernels behind the ViLBERT two-stream
 * co-attentional encoder hot path (reference: vilbert/vilbert.py:396-1107 BertLayer / BertImageLayer /
 * BertConnectionLayer / BertEncoder, :320-367 + :1409-1432 embeddings, :1110-1137 poolers,
 * :1140-1258 + :1638-1722 heads).
 *
 * The reference has no FFI of its own (it is pure PyTorch); every entry point below
 * cites the reference nn.Module / expression whose arithmetic it replaces. The Python host
 * (vilbert-multi-task_b200/) binds these with ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - the caller owns every buffer; the library never allocates or frees device memory;
 *   - every launch is asynchronous on the caller-supplied stream (cudaStream_t passed as void*);
 *   - every function returns a vb_status (0 = ok); vb_last_error() gives a message for the calling
 *     thread; no C++ exception crosses the boundary;
 *   - "bf16" buffers are raw uint16 bfloat16; "f32" are IEEE binary32; matrices are row-major with
 *     an explicit leading dimension in ELEMENTS.
 */
#ifndef VILBERT_B200_H_
#define VILBERT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int vb_status;
enum {
  VB_OK = 0,
  VB_ERR_INVALID = 1,     /* bad shape / alignment / argument                     */
  VB_ERR_UNSUPPORTED = 2, /* device is not sm_100 or an unsupported configuration */
  VB_ERR_CUDA = 3         /* a CUDA runtime / driver call failed                  */
};

enum { VB_ACT_NONE = 0, VB_ACT_GELU = 1, VB_ACT_RELU = 2, VB_ACT_DGELU = 3 };

/* Dropout descriptor (nn.Dropout at vilbert.py:365,443,472,515,604,631,676,778,800,848-851,1233,1430,1678-1695).
 * Masks are not stored: keep(element) = hash32(index ^ hash32(site + step * 0x9E3779B9)) >= p * 2^32 with
 * hash32 = lowbias32 (x^=x>>16; x*=0x7feb352d; x^=x>>15; x*=0x846ca68b; x^=x>>16), kept values scaled by 1/(1-p).
 * `step` is a device uint32 bumped once per training step; `site` names the dropout layer; `index` is the
 * row-major element index of the tensor the reference applies nn.Dropout to (mod 2^32). step == NULL or p == 0
 * disables dropout (the reference's eval mode). */
typedef struct vb_dropout {
  const uint32_t* step;
  uint32_t site;
  float p;
} vb_dropout;

/* ABI version of this header (bumped on incompatible change). */
int vb_version(void);
/* Message for the last non-OK status returned on this thread ("" if none). */
const char* vb_last_error(void);
/* Number of SMs / compute capability (major*10+minor) of the current device. */
vb_status vb_device_info(int* sm_count, int* cc);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores (TMA -> 128B-swizzled smem -> tcgen05.mma -> TMEM).
 *   D[M,N] = alpha * sum_k A(m,k) * B(n,k)   followed by the fused epilogue
 *   v = D (+ bias[n]); act; (+ residual[m,n]); -> out_f32 / out_bf16
 * Replaces every nn.Linear on the path (vilbert.py:410-412,466,492,509 text; :553-555,625,653,670
 * image; :716-725,830,837,865-869 connection; :1116,1131 poolers; heads :1143,1163,1183,1250,1714-1719)
 * and their autograd (dgrad / wgrad):
 *   forward   y = x W^T + b          A = x  [M,K] k-major,  B = W  [N,K] k-major
 *   dgrad     dx = dy W              A = dy [M,N'] k-major, B = W  [N',K'] stored [red, out] -> b_mn_major
 *   wgrad     dW = dy^T x            A = dy [red, out] -> a_mn_major, B = x [red, in] -> b_mn_major
 * Operand storage:
 *   k-major  : element (row r of the M/N extent, reduction index k) at ptr[r*ld + k]
 *   mn-major : element (r, k) at ptr[k*ld + r]
 * Requirements: ld % 8 == 0, base pointers 16-byte aligned. M, N, K arbitrary (TMA zero-fills the
 * edges, stores are predicated).
 * act: VB_ACT_GELU = erf GELU (vilbert.py:111-117; erf by Abramowitz-Stegun 7.1.26, |err| < 5e-7), out_pre receives
 *      gelu'(pre-activation) as bf16 (what the backward needs); VB_ACT_DGELU multiplies by aux[m,n] (that buffer);
 * atomic_out: accumulate into out_f32 with red.global.add (needed when split_k > 1).
 */
typedef struct vb_gemm_args {
  int32_t M, N, K;
  const void* A;      /* bf16 */
  int64_t lda;
  int32_t a_mn_major;
  const void* B;      /* bf16 */
  int64_t ldb;
  int32_t b_mn_major;
  float alpha;
  const float* bias;     /* [N] or NULL */
  const float* residual; /* f32 [M,N] or NULL; may alias out_f32 */
  int64_t ld_res;
  const void* aux;       /* bf16 [M,N] saved gelu'(pre) for VB_ACT_DGELU, else NULL */
  int64_t ld_aux;
  int32_t act;
  float* out_f32;        /* or NULL */
  int64_t ld_out_f32;
  void* out_bf16;        /* or NULL */
  int64_t ld_out_bf16;
  void* out_pre;         /* bf16 gelu'(pre-activation) (GELU) or NULL */
  int64_t ld_out_pre;
  int32_t atomic_out;    /* 0 store, 1 red.add into out_f32 */
  float* out_colsum;     /* [N] or NULL: += column sums of the epilogue value before the residual add (bias gradients) */
  vb_dropout dropout;    /* applied to the epilogue value before the residual add (index m*N + n): LN(dropout(dense(x)) + res) */
  int32_t split_k;       /* >= 1; > 1 requires atomic_out and no act / bf16 outputs */
  int32_t block_n;       /* 0 = auto, else 128 or 256 */
  int32_t max_ctas;      /* 0 = one persistent CTA per SM */
  /* debug/test overrides for the smem matrix descriptors (0 = library default) */
  uint32_t dbg_lbo_a, dbg_sbo_a, dbg_lbo_b, dbg_sbo_b;
  void* dbg_timeline;    /* NULL, or u64 [grid][10]: per-CTA clock64 / globaltimer stamps (development only) */
  int32_t cluster_m;     /* 0 = auto, 1 = no clusters, 2 = CTA pairs (tcgen05 cta_group::2) on adjacent row blocks */
  /* ---- ABI v2: 16-bit operand formats and split precision -------------------------------------------------
   * Forward operands (activations, weights) are IEEE fp16 (11 significant bits; the reference's own reduced
   * precision mode is fp16, train_concap.py:504-505), gradient operands are bf16 (range). a_fp16 / b_fp16 / out_fp16:
   * 0 = bf16, 1 = fp16 (out_fp16 is the format of out_bf16 and out_lo). A and B must have the SAME format: tcgen05
   * kind::f16 encodes them separately but B200 raises an illegal-instruction fault on fp16 x bf16 (measured, round 2),
   * so the backward contractions (dy bf16) read bf16 copies of the forward operands: out_b16 (same ld as out_bf16) is an
   * additional, always-bf16 copy of the 16-bit output, written by the forward GEMM for the weight-gradient GEMM.
   * Split precision ("fp32 parity mode", 1e-3): an operand x is stored as hi = fp16(x), lo = fp16(x - hi); with
   * A_lo and/or B_lo given the contraction is A.B + A_lo.B + A.B_lo (three passes over K into the same TMEM
   * accumulator; the lo.lo term, 2^-22 relative, is dropped). A_lo / B_lo use lda / ldb and the major of A / B.
   * out_lo (same ld as out_bf16) receives the low part of the value written to out_bf16. */
  int32_t a_fp16, b_fp16, out_fp16;
  const void* A_lo;
  const void* B_lo;
  void* out_lo;
  void* out_b16;
} vb_gemm_args;

vb_status vb_gemm_bf16(const vb_gemm_args* args, void* stream);

/* Host-only query: the tile configuration vb_gemm_bf16 would use for `args` (fields block_n / cluster_m / split_k that are
   non-zero in `args` are honoured) on a device with `sm_count` SMs (0 = the current CUDA device). No GPU work, no device
   needed when sm_count > 0; only the shape, layout, epilogue and output fields of `args` are read. */
vb_status vb_gemm_plan(const vb_gemm_args* args, int32_t sm_count, int32_t* block_n, int32_t* cluster_m, int32_t* split_k);

/* ------------------------------------------------------------------------------------------------
 * Fused attention:  P = softmax(Q K^T * scale + mask[b, key]),  O = P V, heads merged in the output.
 * Replaces BertSelfAttention.forward (vilbert.py:424-460), BertImageSelfAttention.forward (:571-619,
 * dynamic_attention off) and both directions of BertBiAttention.forward (:771-809), including the
 * transpose_for_scores / permute().contiguous() layout ops (:416-422, :447-449).
 *   Q: bf16, element (b, i, h, d) at Q[(b*Nq + i)*ldq + h*D + d]   (read in place from a packed QKV buffer)
 *   K, V: same with Nk / ldk / ldv;  O: bf16 [B*Nq, H*D] with ldo
 *   mask: f32 [B, Nk] additive (0 / -10000, vilbert.py:1350-1362) or NULL
 *   lse:  f32 [B, H, Nq] row log-sum-exp in the log2 domain (saved for backward; may be NULL in fwd)
 * Backward recomputes P from lse: needs dO (bf16), writes dQ/dK/dV (bf16, same indexing as Q/K/V with
 * their own ld) and uses delta [B, H, Nq] f32 as scratch. D in {16, 32, 64, 128}; Nq, Nk <= ~320.
 */
typedef struct vb_attn_args {
  int32_t B, H, Nq, Nk, D;
  const void* Q; int64_t ldq;
  const void* K; int64_t ldk;
  const void* V; int64_t ldv;
  const float* mask;
  float scale;
  void* O; int64_t ldo;
  float* lse;
  const void* dO; int64_t lddo;
  void* dQ; int64_t lddq;
  void* dK; int64_t lddk;
  void* dV; int64_t lddv;
  float* delta;
  /* optional (backward): += column sums of dQ / dK / dV, f32 [H*D] each — the bias gradients of the projections */
  float* dbias_q; float* dbias_k; float* dbias_v;
  vb_dropout dropout;   /* on the probabilities; element index ((b*H + h)*Nq + q)*Nk + k */
  /* ---- ABI v2. qkv_fp16: Q, K, V and O are fp16 (forward operands) instead of bf16; dO / dQ / dK / dV are always bf16
   * (the backward converts its Q / K / V panels to bf16 in shared memory). Split precision (forward only): with Q_lo,
   * K_lo, V_lo given (same ld and indexing as Q / K / V) S = Q K^T + Q_lo K^T + Q K_lo^T and O = P V + P_lo V + P V_lo
   * with P split in registers; O_lo (same ld as O) receives the low part of O. */
  int32_t qkv_fp16;
  const void* Q_lo; const void* K_lo; const void* V_lo;
  void* O_lo;
  void* O_b16;   /* forward: optional always-bf16 copy of O (same ldo): the operand of the out-projection's weight gradient.
                    backward: if given, delta = rowsum(dO o O) reads this copy (consistent with the bf16 products of the backward) */
} vb_attn_args;

vb_status vb_attention_fwd(const vb_attn_args* args, void* stream);
vb_status vb_attention_bwd(const vb_attn_args* args, void* stream);
/* Attention-probability export of config.visualization (attn_data["attn"], vilbert.py:451-458, 610-617, 813-821):
 * probs f32 [B, H, Nq, Nk] = softmax(Q K^T * scale + mask) from the Q / K / mask / scale fields of args (eval mode: no dropout). */
vb_status vb_attention_probs(const vb_attn_args* args, float* probs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise (HBM-bound) kernels. One warp per row, 128-bit accesses.
 */

/* BertLayerNorm.forward (vilbert.py:304-317): biased variance, eps inside the sqrt, affine after.
 * x f32 [M,H] (ldx); writes y as f32 and/or bf16 (either may be NULL; both use ldy) and the row
 * statistics mean/rstd [M] (may be NULL). H % 4 == 0, H <= 2048. */
vb_status vb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                           float* y_f32, void* y_bf16, int64_t ldy, float* mean, float* rstd,
                           int32_t M, int32_t H, const vb_dropout* out_dropout /* may be NULL: dropout(LN(x)), embeddings */,
                           int32_t y_fp16 /* format of y_bf16 / y_lo: 0 = bf16, 1 = fp16 */,
                           void* y_lo /* NULL, or the split-precision low part of y_bf16 (same ldy) */,
                           void* y_b16 /* NULL, or an always-bf16 copy of y (same ldy): weight-gradient operand */, void* stream);
/* Autograd of the above. dx as f32 and/or bf16; dgamma/dbeta are ACCUMULATED (atomics) and may be NULL.
 * If gelu_pre (bf16 [M,H], the GELU derivative saved by the forward GEMM) is given, dx_bf16 is multiplied by it — the
 * Linear -> GELU -> LayerNorm head transforms (vilbert.py:1152-1156, 1172-1176, 1714-1718).
 * dbias (may be NULL) += column sums of the dx value written to dx_bf16 (or of dx when dx_bf16 is NULL): the
 * bias gradient of the Linear that produced the LayerNorm input. */
vb_status vb_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                           const float* mean, const float* rstd, float* dx_f32, void* dx_bf16, int64_t lddx,
                           const void* gelu_pre, int64_t ld_pre, float* dgamma, float* dbeta, float* dbias,
                           int32_t M, int32_t H,
                           const vb_dropout* out_dropout /* NULL or the mask applied to this LN's OUTPUT in forward: dy is masked first */,
                           const vb_dropout* in_dropout  /* NULL or the mask applied to the dense output feeding this LN: dx_bf16 / dbias are masked */,
                           void* stream);

/* fp32 -> 16-bit casts: flat (weights shadow, region-feature ingest; bf16 or fp16, optionally hi + lo) and 2-D with independent leading
 * dimensions and a scale (pads operands whose row length is not a multiple of 8). */
vb_status vb_cast_f32_to_bf16(const float* src, void* dst, int64_t n, int32_t fp16 /* 0 = bf16, 1 = fp16 */,
                              void* dst_lo /* NULL or split-precision low part */, void* dst_b16 /* NULL or always-bf16 copy */,
                              void* stream);
vb_status vb_cast2d_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int32_t rows, int32_t cols,
                                float scale, void* stream);

/* BertEmbeddings.forward before its LayerNorm (vilbert.py:346-362): out[b,p,:] = word[ids] + pos[arange] +
 * type[token_type_ids]; if task_ids != NULL the task embedding row is inserted at position 1 (no pos/type
 * term) and the output has Nt+1 rows per sample. ids / token_type_ids [B,Nt] int64, task_ids [B] int64.
 * Backward scatter-adds into the tables (word row 0 = padding_idx gets no gradient, :328-330). */
vb_status vb_embed_text_fwd(const int64_t* ids, const int64_t* token_type_ids, const int64_t* task_ids,
                            const float* word, const float* pos, const float* type, const float* task,
                            float* out, int32_t B, int32_t Nt, int32_t H, void* stream);
vb_status vb_embed_text_bwd(const float* dout, const int64_t* ids, const int64_t* token_type_ids,
                            const int64_t* task_ids, float* dword, float* dpos, float* dtype, float* dtask,
                            int32_t B, int32_t Nt, int32_t H, void* stream);

/* BertImageEmbeddings.image_location_embeddings (vilbert.py:1416,1424): out[m,:] = loc[m,:5] W^T + b,
 * W [H,5]; consumed as the residual of the 2048 -> Hv region-feature GEMM. Backward accumulates dW, db. */
vb_status vb_loc_proj_fwd(const float* loc, const float* W, const float* b, float* out, int32_t M, int32_t H, void* stream);
vb_status vb_loc_proj_bwd(const float* dy, const float* loc, float* dW, float* db, int32_t M, int32_t H, void* stream);

/* Bias gradients: out[n] += sum_m X[m,n]; X is bf16 (is_bf16 != 0) or f32, [M,N] with ld. */
vb_status vb_colsum(const void* X, int32_t is_bf16, int64_t ld, float* out, int32_t M, int32_t N, void* stream);

/* Linears with 1..8 outputs (vil_logit, vil_tri_prediction, vision_logit, linguisic_logit,
 * bi_seq_relationship, the 2-way output of vil_binary_prediction; vilbert.py:1231,1620-1628,1684-1695):
 * y[m,j] = x[m,:] . W[j,:] + b[j] (+ row_addend[m]). Backward: dx (=, or += when accumulate_dx),
 * dW and db ACCUMULATED. */
vb_status vb_small_linear_fwd(const float* x, int64_t ldx, const float* W, const float* b, const float* row_addend,
                              float* y, int32_t M, int32_t K, int32_t N,
                              const vb_dropout* in_dropout /* NULL or dropout applied to x first (index m*K + k; needs ldx == K) */, void* stream);
vb_status vb_small_linear_bwd(const float* dy, const float* x, int64_t ldx, const float* W, float* dx, int64_t lddx,
                              int32_t accumulate_dx, float* dW, float* db, int32_t M, int32_t K, int32_t N,
                              const vb_dropout* in_dropout, void* stream);

/* pooled_output = pooled_t (*|+) pooled_v (fusion_method, vilbert.py:1677-1682, 1236-1241); backward
 * ACCUMULATES into da / db. */
vb_status vb_fuse_pooled_fwd(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t n, int32_t mul,
                             const vb_dropout* dropout /* NULL or dropout on the fused vector (index i) */,
                             int32_t out_fp16, void* out_lo /* format / split-precision low part of out_bf16 */,
                             void* out_b16 /* NULL or always-bf16 copy */, void* stream);
vb_status vb_fuse_pooled_bwd(const float* d, const float* a, const float* b, float* da, float* db, int64_t n, int32_t mul,
                             const vb_dropout* dropout, void* stream);
/* ReLU backward of the poolers (vilbert.py:1121,1136): dx = dy * (y > 0). */
vb_status vb_relu_bwd(const float* dy, const float* y, void* dx_bf16, float* dx_f32, int64_t n, void* stream);
/* y += alpha * x (f32): merges gradient contributions. */
vb_status vb_axpy_f32(const float* x, float* y, int64_t n, float alpha, void* stream);

/* VQA objective (task_utils.py:325-327): loss = mean(BCEWithLogits(logits, target)) * cols, written to
 * *loss (device scalar); dlogits = grad_scale * d loss / d logits as f32 and/or bf16 (ld). */
vb_status vb_bce_logits_loss(const float* logits, const float* target, float* loss, float* dlogits_f32, void* dlogits_bf16,
                             int64_t ld_dlogits_bf16, int32_t rows, int32_t cols, float grad_scale, void* stream);

/* Softmax cross-entropy, reduction = mean over the rows whose label != ignore_index (F.cross_entropy / nn.CrossEntropyLoss as
 * used at vilbert.py:1578-1590 for the masked-LM (30522-way, ignore_index -1) and alignment objectives and at
 * task_utils.py:339-343, 366-374 for the VL-logit / binary / tri heads). *loss (device scalar) = the mean (+= when
 * accumulate_loss); dlogits = grad_scale * d loss / d logits as f32 and/or bf16 (the operand of the head's backward GEMMs),
 * zero on ignored rows. No rows to average -> loss = NaN like torch, gradients 0. */
vb_status vb_ce_loss(const float* logits, int64_t ld_logits, const int64_t* labels, int64_t ignore_index, float* loss,
                     float* dlogits_f32, int64_t ld_d32, void* dlogits_bf16, int64_t ld_d16, int32_t rows, int32_t cols,
                     float grad_scale, int32_t accumulate_loss, void* stream);

/* Masked-region KL objective of BertForMultiModalPreTraining (visual_target == 0, vilbert.py:1506-1525):
 *   loss = sum_{b,r: label[b,r]==1} sum_c t_c (log t_c - log_softmax(scores[b, r+1, :])_c) / max(#(label == 1), 0)
 * scores f32 [B, Nv, C] (region 0 = the global feature is skipped, :1506), target f32 [B, Nv-1, C], label int64 [B, Nv-1].
 * dscores (f32 [B,Nv,C] and/or bf16 with row pitch ld_d16) = grad_scale * d loss / d scores, zero on unmasked rows. */
vb_status vb_kl_masked_loss(const float* scores, const float* target, const int64_t* label, float* loss, float* dscores_f32,
                            void* dscores_bf16, int64_t ld_d16, int32_t B, int32_t Nv, int32_t C, float grad_scale,
                            int32_t accumulate_loss, void* stream);

/* config.dynamic_attention (BertImageSelfAttention, vilbert.py:557-586): the image self-attention's queries and keys are scaled per
 * (sample, channel) by gate = 1 + sigmoid(dyLinear(pool)), pool = mean of the current text states over the unmasked tokens.
 *   vb_masked_mean_fwd  pool[b,:] = sum_n m[b,n] x[b,n,:] / sum_n m[b,n]; x f32 [B,N,H]; add_mask f32 [B,N] is the additive text mask
 *                       ((1-m) * -10000, what vb_mask_to_additive writes); outputs pool f32 [B,H] + its GEMM operand copies
 *                       (pool16 in the forward format, optional low part and bf16 copy)
 *   vb_masked_mean_bwd  dx[b,n,:] (+)= m[b,n] / sum_n m[b,n] * dpool[b,:]
 *   vb_gate_scale_fwd   qk[b*N+n, c] *= 1 + sigmoid(z[b,c]) for c < cols, in place on the Q|K sections of the 16-bit projection
 *                       buffer (row pitch ld elements, hi (+ lo) parts, fp16 or bf16); z f32 [B, cols] = dyLinear_q | dyLinear_k outputs
 *   vb_gate_scale_bwd   dqk (bf16, in place) <- gate * dqk;  dz[b,c] = s(1-s) * sum_n dqk[b,n,c] qk[b,n,c] / gate, s = sigmoid(z);
 *                       qk is the GATED forward buffer; dz f32 [B, cols] and its bf16 operand copy dz16 */
vb_status vb_masked_mean_fwd(const float* x, const float* add_mask, float* pool, void* pool16, void* pool16_lo, void* pool16_b, int32_t out_fp16,
                             int32_t B, int32_t N, int32_t H, void* stream);
vb_status vb_masked_mean_bwd(const float* dpool, const float* add_mask, float* dx, int32_t accumulate, int32_t B, int32_t N, int32_t H, void* stream);
vb_status vb_gate_scale_fwd(void* qk, void* qk_lo, int64_t ld, const float* z, int32_t B, int32_t N, int32_t cols, int32_t fp16, void* stream);
vb_status vb_gate_scale_bwd(void* dqk, int64_t ldd, const void* qk, const void* qk_lo, int64_t ld, const float* z, float* dz, void* dz16, int32_t B,
                            int32_t N, int32_t cols, int32_t fp16, void* stream);

/* Masked-LM head without materialising the [tokens, 30522] logits when only the loss is wanted: only rows with label != ignore_index
 * enter the cross-entropy (vilbert.py:1578-1583), so the tied decoder GEMM, its CE and its backward run on those rows alone.
 *   vb_compact_rows      idx[r] = r-th row with labels[row] != ignore_index (-1 beyond the count), labels_compact[r] its label
 *                        (ignore_index beyond), *count = number of such rows (compare with cap on the host: rows past cap are dropped)
 *   vb_gather_rows16     dst[r,:] = src[idx[r],:] (zeros where idx[r] < 0), 16-bit rows, optionally a second (src2, dst2) pair
 *   vb_scatter_rows_f32  dst[idx[r],:] = src[r,:] for idx[r] >= 0 (the caller zeroes dst): gradient of the gather. With count and
 *                        poison given, *poison = NaN when *count > cap (rows were dropped: the loss must not look valid) */
vb_status vb_compact_rows(const int64_t* labels, int64_t ignore_index, int32_t rows, int32_t cap, int32_t* idx, int32_t* count,
                          int64_t* labels_compact, void* stream);
vb_status vb_gather_rows16(const void* src, void* dst, const void* src2, void* dst2, const int32_t* idx, int32_t cap, int32_t cols, void* stream);
vb_status vb_scatter_rows_f32(const float* src, float* dst, const int32_t* idx, int32_t cap, int32_t cols, const int32_t* count, float* poison,
                              void* stream);

/* Additive attention masks of BertModel.forward (vilbert.py:1341-1362): out[b,j] = (1 - mask[b,j]) * -10000,
 * mask int64 0/1 [B,N]; prepend_one != 0 emits N+1 entries per row with a leading 0 (task-token mask
 * extension, :1331-1334). */
vb_status vb_mask_to_additive(const int64_t* mask, float* out, int32_t B, int32_t N, int32_t prepend_one, void* stream);

/* dst[r] = src for r < repeats (`bytes` each, multiple of 16): BertEncoder's FAST_MODE (vilbert.py:1042-1053) — the text stream
 * of ONE caption, computed at batch 1 up to the first connection layer, broadcast to the image batch (txt_embedding.expand). */
vb_status vb_broadcast_rows(const void* src, void* dst, int64_t bytes, int32_t repeats, void* stream);

/* dst[i * repeats + r] = src[i] for `items` items of `bytes` each (multiple of 16): the batch expansion of the reference's
 * `process: expand / dialog` tasks (task_utils.py:248-274, 198-246): region features / boxes / masks of one image replicated once
 * per answer option, features.unsqueeze(1).expand(B, options, ...).contiguous().view(-1, ...), done in one pass on the device. */
vb_status vb_repeat_rows(const void* src, void* dst, int64_t bytes, int64_t items, int32_t repeats, void* stream);

/* dst[k*n + e] (+)= sum_{r < count_r} src[k*stride_k + r*stride_r + e] for k < count_k, e < n (f32; n and the strides multiples of 4):
 * autograd of BertEncoder's in_batch_pairs expansion (vilbert.py:1008-1040), where every text / image item is expanded to B pairs. */
vb_status vb_sum_strided(const float* src, float* dst, int64_t n, int32_t count_k, int64_t stride_k, int32_t count_r, int64_t stride_r,
                         int32_t accumulate, void* stream);

/* step += 1 on the device (the dropout step counter; one launch per training step, capturable in a CUDA graph). */
vb_status vb_step_counter_bump(uint32_t* step, void* stream);

vb_status vb_memset_zero(void* ptr, int64_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-tensor AdamW on flat buffers. Replaces pytorch_transformers==1.0.0 AdamW as the reference
 * builds it (train_tasks.py:401-426: one param group per tensor with its own lr / weight_decay, correct_bias=False),
 * optimizer.step() + model.zero_grad() (train_tasks.py:550-551), and the engine's own weight-shadow cast:
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= step_size m / (sqrt(v) + eps);  p -= lr wd p   (decay after, on the new p)
 *   step_size = lr, or lr sqrt(1-b2^t)/(1-b1^t) when correct_bias (t = *step, a device int32 the caller advances)
 * p / g / m / v: flat f32 buffers with one layout. Work list: chunk c covers elements [chunk_start[c], +chunk_count[c]) of one
 * tensor (starts multiples of 4) and uses groups[chunk_group[c]] (a DEVICE array, rewritten by the host when a scheduler
 * changes an lr). g is multiplied by grad_scale on read (gradient accumulation / loss scaling) and zeroed when zero_grad.
 * p16 (may be NULL): the 16-bit operand copy of the updated weights (bf16, or fp16 when p16_fp16), p16_lo its
 * split-precision low part (may be NULL), p16_b an always-bf16 copy for the backward (may be NULL). */
typedef struct vb_adamw_group {
  float lr, beta1, beta2, eps, weight_decay;
  int32_t correct_bias;
} vb_adamw_group;

vb_status vb_adamw_step(float* p, float* g, float* m, float* v, void* p16, void* p16_lo, void* p16_b, int32_t p16_fp16,
                        const int64_t* chunk_start, const int32_t* chunk_count, const int32_t* chunk_group, int32_t n_chunks,
                        const vb_adamw_group* groups, const int32_t* step, float grad_scale, int32_t zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VILBERT_B200_H_ */
