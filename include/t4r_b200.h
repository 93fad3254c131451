/* t4r_b200.h -- C ABI of libt4r_b200.so: the sm_100a hot path behind the
 * TabularSequenceFeatures / TransformerBlock / NextItemPredictionTask surface
 * of NVIDIA-Merlin/Transformers4Rec.
 *
 * The reference has no FFI of its own (it is pure Python dispatching to ATen);
 * each entry point below replaces the torch call sequence cited next to it
 * (paths relative to the upstream repository; "HF:" = Hugging Face transformers,
 * where the encoder arithmetic lives).  INTEGRATION.md shows the ctypes binding
 * a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *    parameter comment says "host"; `stream` is a cudaStream_t passed as void*.
 *  - every call is asynchronous w.r.t. the host (no implicit synchronisation),
 *    never allocates or frees caller memory, and takes scratch from the caller
 *    (`*_workspace_bytes` queries).
 *  - return value 0 = OK; non-zero = error, message via t4r_last_error()
 *    (thread-local).  There is NO CPU fallback: a missing GPU or an unsupported
 *    shape is an error, never a silent slow path.
 *  - "planes" = the split-bf16 operand format consumed by the tcgen05 GEMMs:
 *    for a logical fp32 matrix X[rows, K] two bf16 matrices hi = bf16(X),
 *    lo = bf16(X - hi), each [rows, Kp] row-major with Kp = round_up(K, 64) and
 *    zero padding, stored back to back ([2, rows, Kp]).  Three bf16 tensor-core
 *    products (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) reproduce the
 *    fp32 product to ~2^-16 relative error.
 */
#ifndef T4R_B200_H_
#define T4R_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T4R_MAX_FEATURES 32
#define T4R_OK 0
#define T4R_ERR_INVALID 1
#define T4R_ERR_CUDA 2
#define T4R_ERR_UNSUPPORTED 3

const char* t4r_last_error(void);
int t4r_version(void);
/* number of kernels this library has launched since load (all threads). */
long long t4r_launch_count(void);
/* sizeof() of the argument structs below as this library was compiled (0 t4r_head_args, 1 t4r_linear_args,
 * 2 t4r_feature_list, 3 t4r_feature, 4 t4r_xlnet_layer, 5 t4r_gpt2_layer; 0 for anything else) and the byte offset
 * of the LAST field of t4r_head_args -- lets a binding check its mirror of the layouts before the first call. */
size_t t4r_sizeof_struct(int which);
size_t t4r_head_args_last_offset(void);

static inline int t4r_round_up64(int k) { return (k + 63) / 64 * 64; }

/* ------------------------------------------------------------------------- *
 * K1  fused multi-table embedding gather + continuous + concat
 *     replaces: EmbeddingFeatures.forward  transformers4rec/torch/features/embedding.py:226-249
 *               (nn.Embedding per feature, features/sequence.py:75-81),
 *               ContinuousFeatures.forward features/continuous.py:60-63,
 *               ConcatFeatures.forward     tabular/aggregation.py:35-47
 *     The caller lists features already in sorted-name order with their column
 *     offsets in the concatenated row.
 * ------------------------------------------------------------------------- */
typedef struct {
  int n_cat;
  int n_cont;
  const float* table[T4R_MAX_FEATURES];     /* [rows, dim] fp32 row-major            */
  const int64_t* ids[T4R_MAX_FEATURES];     /* [M] int64                              */
  int64_t table_rows[T4R_MAX_FEATURES];
  int dim[T4R_MAX_FEATURES];
  int cat_col[T4R_MAX_FEATURES];            /* first output column of the feature     */
  const float* cont[T4R_MAX_FEATURES];      /* [M] fp32                               */
  int cont_col[T4R_MAX_FEATURES];
} t4r_feature_list;

/* out_f32: [M, C] or NULL; out_planes: bf16 [2, M, Cp] or NULL (Cp = round_up64(C)).
 * err_flag: optional int32 device word set to 1 if any id is outside [0, rows). */
int t4r_embed_concat_fwd(const t4r_feature_list* feats /*host*/, int64_t M, int C, float* out_f32,
                         void* out_planes, int32_t* err_flag, void* stream);

/* ------------------------------------------------------------------------- *
 * K1b general input block: every feature kind, per-feature LayerNorm and aggregation that
 *     TabularSequenceFeatures can be configured with, one kernel.
 *     replaces: SoftEmbedding.forward            features/embedding.py:517-556
 *               TabularLayerNorm.forward         tabular/transformations.py:95-141
 *               ConcatFeatures / ElementwiseSum / ElementwiseSumItemMulti
 *                                                tabular/aggregation.py:35-47,139-193
 *               _expand_non_sequential_features  tabular/base.py:53-63 (per_session)
 *               the "continuous_projection" branch features/tabular.py:88-118 enters as a DENSE
 *               feature (its MLP runs through t4r_linear_fwd first)
 *     Features are listed in sorted-name order (aggregation.py:42-47).  For the element-wise
 *     aggregations every feature must have width C (the reference raises otherwise).
 * ------------------------------------------------------------------------- */
#define T4R_FEAT_CAT 0   /* input: int64 ids; table [card, dim]                                  */
#define T4R_FEAT_CONT 1  /* input: fp32 scalars; dim = 1                                         */
#define T4R_FEAT_SOFT 2  /* input: fp32 scalars; softmax(x*soft_w + soft_b) [card] @ table [card, dim] */
#define T4R_FEAT_DENSE 3 /* input: fp32 [rows, dim] already computed                             */
#define T4R_AGG_CONCAT 0
#define T4R_AGG_SUM 1            /* "element-wise-sum"            */
#define T4R_AGG_SUM_ITEM_MULTI 2 /* "element-wise-sum-item-multi": item * sum(others) */
typedef struct {
  int kind;
  int dim;               /* output width of the feature                                  */
  int col;               /* first output column (concat only)                            */
  int card;              /* CAT: table rows (ids outside [0, card) raise err_flag); SOFT: bins */
  int per_session;       /* 1: one input entry per session, repeated over its L positions */
  int reserved;
  const void* input;
  const float* table;
  const float* soft_w;   /* [card]  Linear(1, card).weight                                */
  const float* soft_b;   /* [card]                                                        */
  const float* ln_gamma; /* [dim] or NULL: LayerNorm over this feature before aggregation */
  const float* ln_beta;
} t4r_feature;
/* M = B*L output rows; C = output width (sum of widths for concat, the common width otherwise). */
int t4r_input_block_fwd(const t4r_feature* feats /*host, n_feats*/, int n_feats, int64_t M, int L, int agg,
                        int item_feature /*index into feats, item-multi only*/, float ln_eps, int C, float* out_f32,
                        void* out_planes, int32_t* err_flag, void* stream);

/* StochasticSwapNoise.augment (tabular/transformations.py:54-92) on one tensor of n 4- or 8-byte
 * elements: positions with keep_mask and u < replacement_prob take values drawn (through `perm`,
 * a permutation of the kept positions' count) from the kept values of the same tensor.
 * keep_mask (uint8, NULL = all) is read at [(i / inner) * keep_stride]; scratch: int32 [n]. */
int t4r_swap_noise(const void* values, int elem_bytes, int64_t n, const uint8_t* keep_mask, int64_t keep_stride,
                   int inner, const float* u, float replacement_prob, const int64_t* perm, int32_t* scratch_pool_pos,
                   void* out, void* stream);

/* ------------------------------------------------------------------------- *
 * K3  mask / label generation (integer, bit-exact)
 *     replaces: MaskedLanguageModeling._compute_masked_targets masking.py:376-470
 *               CausalLanguageModeling._compute_masked_targets masking.py:274-300
 *               MaskSequence.predict_all masking.py:182-213
 *     Randomness enters only through `u` (uniforms in [0,1), [B, L+2]:
 *     u[b, 0..L) bernoulli draws, u[b, L] "force one label", u[b, L+1] "un-label
 *     one"); see DESIGN.md "Random draws".
 *     row_code (uint8 [B, Lout]) tells the projection epilogue what
 *     apply_mask_to_inputs (masking.py:473-498 / 302-337) does to that position:
 *     0 keep, 1 replace by masked_item_embedding, 2 zero.
 * ------------------------------------------------------------------------- */
#define T4R_MLM_TRAIN 0
#define T4R_MLM_EVAL_LAST 1
#define T4R_MLM_EVAL_ALL 2
#define T4R_MLM_INFERENCE 3 /* outputs have L+1 columns */
int t4r_mask_mlm(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode, float mlm_probability,
                 const float* u, uint8_t* mask_schema, int64_t* masked_targets, uint8_t* row_code, void* stream);

#define T4R_CLM_ALL 0       /* training (or eval with eval_on_last_item_seq_only=False) */
#define T4R_CLM_LAST 1      /* last item only (eval default / train_on_last_item_seq_only) */
#define T4R_CLM_INFERENCE 2
int t4r_mask_clm(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode, uint8_t* mask_schema,
                 int64_t* masked_targets, uint8_t* row_code, void* stream);

/* Permutation Language Modeling (XLNet): PermutationLanguageModeling._compute_masked_targets_extended
 * masking.py:548-727.  One thread per session (the reference loops over the sessions in python).  Outputs:
 * mask_schema / masked_targets [B, L] and perm_mask [B, L, L] bytes (1 = query i may NOT attend key j);
 * target_mapping is the identity in every mode the reference builds it in and is not materialised.
 * Training consumes explicit draws, per session: u_span / u_start [B, L] (iteration n of the span loop:
 * span = 1 + floor(u_span * max_span), start = cur + floor(u_start * (ctx - span + 1))), u_force / u_unmask [B],
 * perm [B, L] int32 = the factorisation order (torch.randperm(L) per session).  ctx_len[span] (HOST array,
 * max_span + 1 entries) = int(span / plm_probability) as the reference computes it.  L <= 64, max_span <= 15.
 * t4r_debug_mask_plm_host is the same code compiled for the host (HOST pointers) -- test infrastructure. */
#define T4R_PLM_TRAIN 0
#define T4R_PLM_EVAL_LAST 1
#define T4R_PLM_EVAL_ALL 2
int t4r_mask_plm(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode, int max_span,
                 const int32_t* ctx_len /*host*/, const float* u_span, const float* u_start, const float* u_force,
                 const float* u_unmask, const int32_t* perm, uint8_t* mask_schema, int64_t* masked_targets,
                 uint8_t* perm_mask, void* stream);
int t4r_debug_mask_plm_host(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode, int max_span,
                            const int32_t* ctx_len, const float* u_span, const float* u_start, const float* u_force,
                            const float* u_unmask, const int32_t* perm, uint8_t* mask_schema, int64_t* masked_targets,
                            uint8_t* perm_mask);

/* Row compaction of label positions (row-major order), replaces the
 * masked_select pair at model/prediction_task.py:436-443,472-479.
 * tgt_rows[i] = flat index (b*L+l) of the i-th non-pad label, tgt_labels[i] its id,
 * *count_dev = T.  Entries i >= T are set to row 0 / label 0. */
int t4r_compact_targets(const int64_t* masked_targets, int64_t n, int64_t padding_idx, int32_t* tgt_rows,
                        int64_t* tgt_labels, int32_t* count_dev, void* stream);

/* ------------------------------------------------------------------------- *
 * N1  ragged (values, offsets) -> dense right-padded / truncated [rows, pad_len]
 *     replaces: _pad_ragged_tensor / _pad_dense_tensor  utils/padding.py:20-68 (sparse_coo -> to_dense
 *               + F.pad), called by pad_batch :71-122 and pad_inputs :125-164 (model/base.py:551).
 *     offsets == NULL means a dense input [rows, in_len] that is padded / truncated to pad_len.
 *     elem_bytes is 8 (int64 ids) or 4 (fp32 continuous features).
 * ------------------------------------------------------------------------- */
int t4r_pad_ragged(const void* values, const int64_t* offsets, int64_t rows, int in_len, int pad_len, int elem_bytes,
                   void* out, void* stream);

/* ------------------------------------------------------------------------- *
 * operand packing / elementwise helpers
 * ------------------------------------------------------------------------- */
/* fp32 [rows, K] (row stride ld) -> planes bf16 [2, rows, Kp].  Optional row_code/
 * mask_vec apply apply_mask_to_inputs (used when there is no projection GEMM to
 * fuse it into); out_f32 (optional, [rows, K]) receives the masked fp32 rows. */
int t4r_split_planes(const float* x, int64_t rows, int K, int ld, const uint8_t* row_code, const float* mask_vec,
                     float* out_f32, void* out_planes, void* stream);
/* Operands of the 2-unit product (nprod = 2 of t4r_head_softmax_ce_fwd): fp32 [rows, K] (row stride ld) ->
 * 16-bit words [2, rows, Kp]: plane 0 = fp16(x * s_row); plane 1 = per 64-wide K block and row 64 e4m3 bytes
 * of the fp16 value (x 2^-6) followed by 64 e4m3 bytes of the fp16 rounding residual (x 2^6);
 * s_row = the power of two that brings max|x_row| into [2^13, 2^14); out_inv_scale[row] = 1 / s_row.
 * Same footprint as the split-bf16 planes.  Layout and rounding are defined in csrc/t4r_mixed_pack.cuh;
 * tools/precision_study.py holds the error analysis (about 2.3x the error of the 3-product bf16 split at
 * 2/3 of its tensor time).  t4r_debug_split_planes_mixed_host is the same code compiled for the host
 * (HOST pointers, no CUDA call) -- test infrastructure. */
int t4r_split_planes_mixed(const float* x, int64_t rows, int K, int ld, void* out_planes, float* out_inv_scale,
                           void* stream);
/* the same with a device-side row count (the head's label rows: capacity B*L, ~13 % of it valid): rows from
 * round_up(*count_dev, 256) on are not touched */
int t4r_split_planes_mixed_n(const float* x, int64_t rows, int K, int ld, const int32_t* count_dev, void* out_planes,
                             float* out_inv_scale, void* stream);
int t4r_debug_split_planes_mixed_host(const float* x, int64_t rows, int K, int ld, void* out_planes,
                                      float* out_inv_scale);
/* gather rows then split: out[i] = x[idx[i]] for i < *count_dev (all `cap` rows when
 * count_dev is NULL); with a count, rows from it up to the next multiple of 256 are zero and the
 * rest of the outputs is left untouched.  out_f32 optional. */
int t4r_gather_rows_split(const float* x, int K, int ld, const int32_t* idx, const int32_t* count_dev, int cap,
                          float* out_f32, void* out_planes, void* stream);
/* idx variant for int64 indices (embedding rows for sampled softmax) */
int t4r_gather_rows_split_i64(const float* x, int K, int ld, const int64_t* idx, int cap, float* out_f32,
                              void* out_planes, void* stream);

/* ------------------------------------------------------------------------- *
 * K2  dense layer on tcgen05:  Y = epilogue(X * W^T)
 *     replaces: DenseBlock (Linear + ReLU) block/mlp.py:123-144 built at
 *               features/sequence.py:213-219, fused with
 *               apply_mask_to_inputs masking.py:473-498 / 302-337; also the
 *               head's task_block Linear model/prediction_task.py:390-397.
 * ------------------------------------------------------------------------- */
#define T4R_ACT_NONE 0
#define T4R_ACT_RELU 1
#define T4R_ACT_GELU 2
typedef struct {
  int64_t M;                 /* rows of X (capacity if m_dev != NULL)                    */
  int N;                     /* output features (any; LayerNorm needs 64, 128 or 256)    */
  int K;                     /* input features (planes are padded to round_up64(K))      */
  const void* x_planes;      /* bf16 [2, M, Kp]                                          */
  const void* w_planes;      /* bf16 [2, N, Kp]  (W is [N, K] like nn.Linear.weight)     */
  const int32_t* m_dev;      /* optional device row count; tiles beyond it are skipped   */
  const float* bias;         /* [N] or NULL                                              */
  int act;                   /* T4R_ACT_*                                                */
  const uint8_t* row_code;   /* [M] or NULL                                              */
  const float* mask_vec;     /* [N] masked_item_embedding (needed when row_code)         */
  const float* residual;     /* [M, N] fp32 or NULL (added after act / mask)             */
  const float* ln_gamma;     /* [N] or NULL: LayerNorm over the N outputs (N <= 256)     */
  const float* ln_beta;
  float ln_eps;
  float* out_pre_ln;         /* optional fp32 [M, N]: value before LayerNorm             */
  float* out_f32;            /* optional fp32 [M, N]: final value                        */
  void* out_planes;          /* optional bf16 [2, M, round_up64(N)]: final value, split  */
  int nprod;                 /* 3 = fp32-grade split product (default), 1 = plain bf16   */
} t4r_linear_args;
int t4r_linear_fwd(const t4r_linear_args* a /*host*/, void* stream);

/* Fused feed-forward block (K7): out = LayerNorm(residual + gelu_erf(X W1^T + b1) W2^T + b2).
 * Replaces XLNetFeedForward (HF modeling_xlnet.py:297-305) and GPT2MLP + residual + the following
 * LayerNorm (HF modeling_gpt2.py:229-243, :305-309).  The [M, hidden] intermediate never reaches HBM:
 * it is produced in TMEM, GELU'd in registers and fed back to the tensor core from TMEM.
 *   x_planes  bf16 [2, M, d]       w1_planes bf16 [2, hidden, d]     w2_planes bf16 [2, d, hidden]
 *   b1 [hidden], b2 [d] (b2 may be NULL), ln_gamma/ln_beta [d]
 *   residual: fp32 [M, d], or NULL to use x itself (hi + lo of x_planes; the XLNet post-LN form)
 *   out_f32 [M, d] / out_pre_ln [M, d] / out_planes [2, M, d]: each optional, at least one.
 * d in {64, 128, 256}; hidden a multiple of 128. */
int t4r_ffn_fwd(const void* x_planes, int64_t M, int d, int hidden, const void* w1_planes, const float* b1,
                const void* w2_planes, const float* b2, const float* residual, const float* ln_gamma,
                const float* ln_beta, float ln_eps, float* out_pre_ln, float* out_f32, void* out_planes, void* stream);

/* debug: probe of the A-from-TMEM form of tcgen05.mma: D[128, N] = bf16(A[128, 64]) * B_hi[N, 64]^T,
 * B given as planes [2, N, 64]; N in {64, 128, 256}. */
int t4r_debug_ts_mma(const float* A, const void* b_planes, int N, float* D, void* stream);

/* debug/reference: plain fp32 SIMT GEMM  C[M,N] = A[M,K] * B[N,K]^T (+bias). Used by
 * the GPU tests to check the tensor-core path at sizes the CPU oracle cannot reach. */
int t4r_debug_sgemm_nt(const float* A, const float* B, const float* bias, float* C, int64_t M, int N, int K,
                       void* stream);

/* ------------------------------------------------------------------------- *
 * K4-K7  XLNet encoder (relative attention, post-LN), all layers
 *     replaces: TransformerBlock.forward block/transformer.py:179-199 ->
 *               HF XLNetModel.forward HF:models/xlnet/modeling_xlnet.py:979-1205
 *               (XLNetRelativeAttention :245-277, rel_attn_core :95-140,
 *                rel_shift_bnij :81-93, post_attention :142-152,
 *                XLNetFeedForward :285-305, relative_positional_encoding :940-976)
 *               built by XLNetConfig.build config/transformer.py:432-482.
 * ------------------------------------------------------------------------- */
typedef struct {
  const void* wqkv_planes;   /* bf16 [2, 3d, d]: rows = [q | k | v] output features       */
  const float* wr;           /* fp32 [d, d] = rel_attn.r reshaped [d_in, H*dh]            */
  const float* r_w_bias;     /* [H*dh]                                                    */
  const float* r_r_bias;     /* [H*dh]                                                    */
  const void* wo_planes;     /* bf16 [2, d, d]   = rel_attn.o reshaped [d_out, H*dh]      */
  const float* ln1_gamma;    /* rel_attn.layer_norm                                       */
  const float* ln1_beta;
  const void* w1_planes;     /* bf16 [2, 4d, d]  = ff.layer_1.weight                      */
  const float* b1;
  const void* w2_planes;     /* bf16 [2, d, 4d]  = ff.layer_2.weight                      */
  const float* b2;
  const float* ln2_gamma;    /* ff.layer_norm                                             */
  const float* ln2_beta;
} t4r_xlnet_layer;
size_t t4r_xlnet_encoder_workspace_bytes(int B, int L, int d, int n_head);
/* x_f32 [B*L, d]; x_planes [2, B*L, d] or NULL (then split internally);
 * out_f32 [B*L, d]; out_planes optional. d multiple of 64, d <= 256, L <= 64. */
int t4r_xlnet_encoder_fwd(const t4r_xlnet_layer* layers /*host*/, int n_layer, int B, int L, int d, int n_head,
                          float ln_eps, const float* x_f32, const void* x_planes, float* out_f32, void* out_planes,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * GPT-2 encoder (pre-LN, learned positions, causal attention)
 *     replaces: TransformerBlock.forward + GPT2Prepare block/transformer.py:55-73,179-199 ->
 *               HF GPT2Model.forward HF:models/gpt2/modeling_gpt2.py:522-636
 *               (GPT2Block :246-309, GPT2Attention :144-226, GPT2MLP :229-243)
 *               built by GPT2Config.build config/transformer.py:217-260.
 * ------------------------------------------------------------------------- */
typedef struct {
  const float* ln1_gamma;
  const float* ln1_beta;
  const void* wqkv_planes;   /* bf16 [2, 3d, d] = c_attn.weight^T                         */
  const float* bqkv;         /* [3d]                                                      */
  const void* wo_planes;     /* bf16 [2, d, d]  = attn.c_proj.weight^T                    */
  const float* bo;
  const float* ln2_gamma;
  const float* ln2_beta;
  const void* w1_planes;     /* bf16 [2, 4d, d] = mlp.c_fc.weight^T                       */
  const float* b1;
  const void* w2_planes;     /* bf16 [2, d, 4d] = mlp.c_proj.weight^T                     */
  const float* b2;
} t4r_gpt2_layer;
/* XLNet two-stream forward for Permutation Language Modeling (HF XLNetModel.forward with perm_mask and
 * target_mapping = identity, as block/transformer.py:179-199 passes masking.transformer_arguments): the caller
 * stacks the content stream h (the masked input embeddings) and the query stream g (mask_emb in every row) into
 * x_f32 [2 B L, d]; both streams go through every layer's GEMMs together, attention reads keys / values from the
 * h rows for both and applies perm_mask [B, L, L] bytes (h: except on the diagonal).  out_f32 [2 B L, d]: the g
 * rows (second half) are what HF returns as output[0].  Workspace: t4r_xlnet_encoder_workspace_bytes(2 B, ...).
 * Needs the tensor-path attention (L + 2 <= 32, or <= 64 with T4R_ATTN_MMA64=1). */
int t4r_xlnet_encoder_plm_fwd(const t4r_xlnet_layer* layers /*host*/, int n_layer, int B, int L, int d, int n_head,
                              float ln_eps, const float* x_f32, const uint8_t* perm_mask, float* out_f32,
                              void* workspace, size_t workspace_bytes, void* stream);

size_t t4r_gpt2_encoder_workspace_bytes(int B, int L, int d, int n_head);
int t4r_gpt2_encoder_fwd(const t4r_gpt2_layer* layers /*host*/, int n_layer, int B, int L, int d, int n_head,
                         float ln_eps, const float* wpe /*[n_positions, d]*/, const float* lnf_gamma,
                         const float* lnf_beta, const float* x_f32, float* out_f32, void* out_planes,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * K8  tied-weight item logits + softmax cross-entropy, never materialising [T, V]
 *     replaces: _NextItemPredictionTask.forward model/prediction_task.py:648-671
 *               + nn.CrossEntropyLoss (mean) :446/:347, and for evaluation
 *               RecallAt ranking_metric.py:111-147 (rank of the label instead of
 *               one-hot + topk, utils/torch_utils.py:226-238).
 * K9  sampled softmax: _NextItemPredictionTask.sampled :673-696 with the
 *               LogUniformSampler probabilities :766-796.
 * ------------------------------------------------------------------------- */
typedef struct {
  int T_cap;                 /* capacity of the target-row buffers                       */
  const int32_t* t_dev;      /* device count of valid rows (NULL -> T_cap)                */
  int De;                    /* embedding dim of the output table                         */
  int64_t V;                 /* number of classes (rows of the table / of the shard)      */
  const void* xt_planes;     /* bf16 [2, T_cap, Dep] hidden rows at label positions       */
  const float* xt_f32;       /* fp32 [T_cap, De] same rows (exact target logit)           */
  const int64_t* labels;     /* [T_cap] class ids                                         */
  const void* w_planes;      /* bf16 [2, V, Dep] split table                              */
  const float* w_f32;        /* fp32 [V, De] table (exact target logit); may be NULL when
                                tgt_logit_in is given                                     */
  float inv_temperature;     /* 1 / softmax_temperature                                   */
  /* sampled softmax (all NULL/0 for the full softmax): the V "classes" are then the
   * S sampled negatives (w_planes = gathered rows) and class 0 is the positive. */
  const float* col_bias;     /* [V] added to each logit (= -log(q+1e-16))                 */
  const int64_t* col_ids;    /* [V] item id of each column, for accidental-hit removal    */
  float hit_value;           /* value accidental hits are set to (-655.04)                */
  const float* pos_logit;    /* [T_cap] positive logit incl. its logQ term, or NULL       */
  /* shard support: this call covers table rows [v_offset, v_offset + V) */
  int64_t v_offset;
  /* outputs */
  float* row_lse;            /* [T_cap] log-sum-exp over this call's classes (+pos)       */
  float* row_tgt;            /* [T_cap] logit of the label (0 if not in this shard)       */
  float* row_loss;           /* [T_cap] lse - target logit (single-shard use)             */
  float* loss;               /* [1] mean over valid rows, or NULL                         */
  int32_t* row_rank;         /* [T_cap] #classes scoring above the label (ties: lower id
                                first), or NULL                                           */
  void* workspace;
  size_t workspace_bytes;
  int nprod;
  /* optional cudaEvent_t pair recorded on `stream` immediately around the logits/LSE
   * GEMM launch (bench.py's per-kernel roofline timing); NULL = off */
  void* ev_gemm_start;
  void* ev_gemm_stop;
  /* nn.CrossEntropyLoss(label_smoothing=e) (transformers4rec/torch/losses.py:4-20): row_loss =
   * lse - (1-e) z_label - (e/V) sum_j z_j.  Full, unsharded softmax only; 0 = off. */
  float label_smoothing;
  /* sharded evaluation: the label's logit over the WHOLE table (summed over shards by the caller)
   * that row_rank counts against; NULL = this call's own row_tgt (single shard) */
  const float* rank_tgt;
  /* nprod = 2 (the 2-unit product: fp16 x fp16 plus two e4m3 cross terms, t4r_split_planes_mixed below):
   * xt_planes / w_planes are then in the mixed format and these are the per-row inverse scales it returns
   * ([T_cap] and [V]); both NULL otherwise.  Full and sampled softmax alike; not for t4r_head_logits. */
  const float* xt_inv_scale;
  const float* w_inv_scale;
  /* sampled softmax: non-zero promises that col_ids is strictly ascending (the reference's negatives are
   * `multinomial(...).unique()[:S]`, model/prediction_task.py:843-845: sorted and unique).  The accidental hit of a row
   * is then found once per row by binary search instead of one id comparison per logit (measured: 45 % of the sampled
   * head's time).  0 = no assumption (per-logit comparison). */
  int col_ids_sorted_unique;
} t4r_head_args;
size_t t4r_head_workspace_bytes(int T_cap, int64_t V, int De);
int t4r_head_softmax_ce_fwd(const t4r_head_args* a /*host*/, void* stream);

/* exact fp32 logit of each row's label plus an optional per-class bias:
 * out[t] = (xt[t] . W[label] + class_bias[label]) * inv_temperature -- the positive
 * score of the sampled softmax incl. its logQ term (model/prediction_task.py:681-687).
 * w_f32 holds table rows [v_offset, v_offset + V); labels outside that range give 0 (so that
 * the per-shard results of a row-sharded table can be summed). */
int t4r_label_logit(const float* xt_f32, const float* w_f32, const int64_t* labels, int T_cap, const int32_t* t_dev,
                    int De, int64_t V, const float* class_bias, float inv_temperature, int64_t v_offset, float* out,
                    void* stream);

/* materialise logits [T_cap, V] = xt * W^T * inv_temperature (the reference's
 * "predictions" output, model/prediction_task.py:447-451) -- optional, on demand. */
int t4r_head_logits(const void* xt_planes, const void* w_planes, int T_cap, const int32_t* t_dev, int64_t V, int De,
                    float inv_temperature, float* out /*[T_cap, ldo]*/, int64_t ldo, int nprod, void* stream);

/* the same from the operands of the 2-unit product (t4r_split_planes_mixed) */
int t4r_head_logits_mixed(const void* xt_planes, const void* w_planes, int T_cap, const int32_t* t_dev, int64_t V, int De,
                          float inv_temperature, float* out /*[T_cap, ldo]*/, int64_t ldo, const float* xt_inv_scale,
                          const float* w_inv_scale, void* stream);

/* ------------------------------------------------------------------------- *
 * N3  element / row kernels of the training step (csrc/t4r_train.cu; composed by
 *     transformers4rec_b200/training.py with the GEMMs above on transposed operands).  All fp32, row-major.
 *     `on_host` != 0 runs the same per-item code in a loop on HOST pointers (no CUDA call): test infrastructure.
 *     Reference counterparts: torch autograd of the modules at block/mlp.py:123-144, HF XLNet / GPT-2 layers,
 *     masking.py:473-498 / :302-337, model/prediction_task.py:648-671 + :446, features/embedding.py:226-249.
 * ------------------------------------------------------------------------- */
int t4r_train_transpose(const float* x, int64_t R, int64_t C, float* out /*[C, R]*/, void* stream, int on_host);
int t4r_train_act_fwd(int kind, const float* x, float* y, int64_t n, void* stream, int on_host);  /* exact erf GELU / ReLU */
int t4r_train_act_bwd(int kind, const float* pre, const float* dy, float* dx, int64_t n, void* stream, int on_host);
int t4r_train_add_positions(const float* x, const float* wpe, int B, int L, int d, float* y, void* stream, int on_host);
int t4r_train_sum_sessions(const float* x, int B, int L, int d, float* out /*[L, d]*/, void* stream, int on_host);
int t4r_train_row_codes_fwd(const float* y, const uint8_t* code, const float* mask_vec, int64_t M, int d, float* out,
                            void* stream, int on_host);
int t4r_train_row_codes_bwd(const float* dx, const uint8_t* code, int64_t M, int d, float* dy, float* dmask /*[d]*/,
                            float* tmp /*[M, d] scratch*/, void* stream, int on_host);
int t4r_train_gather_rows(const float* x, const int32_t* idx, int64_t n, int d, float* out, void* stream, int on_host);
int t4r_train_scatter_rows(const float* src, const int32_t* idx, int64_t n, int d, int64_t out_rows, float* out,
                           void* stream, int on_host);
/* in place: z[t, j] <- (exp(z[t, j] - lse[t]) - (1 - e) [labels[t] == v0 + j] - e / V_total) * scale,
 * e = label_smoothing (0 = plain cross-entropy) */
int t4r_train_softmax_ce_bwd(float* z, const float* lse, const int64_t* labels, int64_t T, int64_t Vc, int64_t v0,
                             float scale, float label_smoothing, int64_t V_total, void* stream, int on_host);
/* sampled softmax, in place: z[t, s] (= x_t . w_s / tau) <- exp(z + col_bias[s] / tau - lse[t]) * scale, 0 where
 * col_ids[s] == labels[t] (accidental hits were constants in the forward) */
int t4r_train_sampled_ce_bwd(float* z, const float* lse, const int64_t* labels, const float* col_bias,
                             const int64_t* col_ids, int64_t T, int64_t S, float inv_tau, float scale, void* stream,
                             int on_host);
/* dst[idx[r], :] += src[r, col : col + width] (dst rows are `width` wide); rows with idx == skip_index are skipped */
int t4r_train_index_add_rows(float* dst, const int64_t* idx, const float* src, int64_t n, int64_t ld_src, int col,
                             int width, int64_t skip_index, void* stream, int on_host);
int t4r_train_col_sum(const float* x, int64_t M, int64_t N, float* out /*[N]*/, void* stream, int on_host);
/* soft embedding of one scalar per row (features/embedding.py:517-556): p[M, n] = softmax(x w + b), out[M, dim] = p table;
 * backward per row: dlogit[M, n] and dlogit * x (d table = p^T dout, d w / d b = column sums: composed by the caller) */
int t4r_train_soft_emb_fwd(const float* x, const float* w, const float* b, const float* table, int64_t M, int n, int dim,
                           float* p, float* out, void* stream, int on_host);
int t4r_train_soft_emb_bwd(const float* x, const float* table, const float* p, const float* dout, int64_t M, int n, int dim,
                           float* dlogit, float* dlogit_x, void* stream, int on_host);
/* out = a + b (op 0) or a * b (op 1), element-wise (aggregation.py:139-193 and its product rule) */
int t4r_train_binary(int op, const float* a, const float* b, float* out, int64_t n, void* stream, int on_host);
/* one AdamW step (torch.optim.AdamW's rule, decoupled weight decay) on a flat fp32 tensor, in place on p / m / v */
int t4r_train_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, void* stream, int on_host);
int t4r_train_layer_norm_fwd(const float* x, const float* gamma, const float* beta, int64_t M, int d, float eps, float* y,
                             void* stream, int on_host);
/* dx = LayerNorm backward of dy at x (+ add, optional); dgamma / dbeta [d] are overwritten (column sums of dy * xhat,
 * kept in tmp, and of dy: no per-element atomics) */
int t4r_train_layer_norm_bwd(const float* x, const float* gamma, int64_t M, int d, float eps, const float* dy,
                             const float* add, float* dx, float* dgamma, float* dbeta, float* tmp /*[M, d] scratch*/,
                             void* stream, int on_host);
/* attention backward, scores recomputed (L <= 64), one work item per (session, head) owning its slices of dq / dk / dv and
 * of the per-session partials of dR / drw / drr (no atomics; the partials are then reduced over the sessions).
 * qkv / dqkv [M, 3d] (q | k | v), dout [M, d].  XLNet relative form: R [2L, d], rw / rr [d], their gradients and
 * part = scratch of B (2L + 2) d floats; all seven NULL selects GPT-2's causal form. */
int t4r_train_attn_bwd(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout, int B, int L,
                       int d, int H, float* dqkv, float* dR, float* drw, float* drr, float* part,
                       const uint8_t* plm_mask /* NULL, or [B, L, L]: two-stream form over 2 B L rows */, void* stream,
                       int on_host);
/* Dropout of the training step (HF:xlnet dropout sites :129, :147, :300-303, :1085-1180; HF:gpt2 :66, :225, :241, :584):
 * y[i] = x[i] * keep(i) / (1 - p) with keep(i) a pure function of (seed, site, i) -- Philox4x32-10, so the backward
 * regenerates the forward's mask by calling the same entry on the gradient.  Works in place (y == x). */
int t4r_train_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t site, void* stream, int on_host);
/* Attention forward of the training graph with dropout of the probabilities (mask index = HF's [stream, B, H, L, L]
 * "bnij" order), one work item per (session, head) like the backward: out [M, d] fp32 (2 M rows in the two-stream
 * form).  R / rw / rr NULL selects GPT-2's causal form.  t4r_train_attn_drop_bwd = t4r_train_attn_bwd under the
 * same (p_drop, seed, site). */
int t4r_train_attn_drop_fwd(const float* qkv, const float* R, const float* rw, const float* rr, int B, int L, int d, int H,
                            const uint8_t* plm_mask, float p_drop, uint64_t seed, uint32_t site, float* out, void* stream,
                            int on_host);
int t4r_train_attn_drop_bwd(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout, int B,
                            int L, int d, int H, float* dqkv, float* dR, float* drw, float* drr, float* part,
                            const uint8_t* plm_mask, float p_drop, uint64_t seed, uint32_t site, void* stream, int on_host);
/* two-stream (PLM) attention forward on split planes [2, 2 B L, 3d] (h rows then g rows), R planes [2, 2L, d] */
int t4r_train_xlnet_attn_plm_fwd(const void* qkv_planes, const void* r_planes, const float* rw, const float* rr, int B,
                                 int L, int d, int H, const uint8_t* plm_mask, void* out_planes, void* stream);
/* forward pieces of the training graph that reuse inference kernels on fp32 q|k|v (device only) */
int t4r_train_xlnet_attn_fwd(const float* qkv, const float* R, const float* rw, const float* rr, int B, int L, int d, int H,
                             void* out_planes /*[2, M, d]*/, void* stream);
int t4r_train_causal_attn_fwd(const float* qkv, int B, int L, int d, int H, void* out_planes, void* stream);
int t4r_train_rel_pos_proj(const float* const* wr /*host array of device ptrs*/, int n_layer, int L, int d,
                           float* r_out /*[n_layer, 2L, d]*/, void* stream);

/* K10 Recall@k from label ranks: out[j] = mean_t(rank[t] < ks[j]).
 * replaces RecallAt._metric + RankingMetric.update ranking_metric.py:52-63,111-147 */
int t4r_recall_from_ranks(const int32_t* row_rank, const int32_t* t_dev, int T_cap, const int32_t* ks /*host*/,
                          int n_ks, float* out, void* stream);

/* The other ranking metrics of ranking_metric.py:73-319 from the same ranks (one relevant item per
 * row): precision = [r<k]/k, reciprocal rank (avg_precision == mrr) = [r<k]/(r+1),
 * dcg (== ndcg) = [r<k]/log2(r+2); out[j] = mean over rows. */
#define T4R_METRIC_RECALL 0
#define T4R_METRIC_PRECISION 1
#define T4R_METRIC_RR 2
#define T4R_METRIC_DCG 3
int t4r_metrics_from_ranks(const int32_t* row_rank, const int32_t* t_dev, int T_cap, int kind,
                           const int32_t* ks /*host*/, int n_ks, float* out, void* stream);

/* top-k (k <= 64) scores and ids of materialised logits, lower id first on ties
 * (model/prediction_task.py:467-470 torch.topk on the inference path). */
int t4r_topk(const float* logits, int64_t rows, int64_t V, int64_t ld, int k, float* out_scores, int64_t* out_ids,
             void* stream);

/* combine per-shard (lse, target-logit) rows gathered from `world` ranks:
 * parts [world, T_cap, 2] -> loss (mean of lse_total - tgt_total). */
int t4r_combine_shard_lse(const float* parts, int world, int T_cap, const int32_t* t_dev, float* row_loss,
                          float* loss, void* stream);

/* ------------------------------------------------------------------------- *
 * K11 / K12 over NVLink PEER MEMORY (csrc/t4r_peer.cu): the row-sharded item table and tied head of
 * BASELINE configs 4-5 with one process per GPU.  The reference replicates the table on every
 * rank (DDP only, docs/source/multi_gpu_train.md:5-50; nn.Embedding features/embedding.py:226-249 and the
 * tied logits model/prediction_task.py:648-671); here rank r holds rows [r * rows_per_shard, ...) and every
 * rank maps its peers' shards and two small per-step windows through CUDA IPC, so that the kernels below read
 * remote rows with plain loads over NVLink -- no bulk collective, no routing plan, exactly the needed bytes.
 * ------------------------------------------------------------------------- */
#define T4R_MAX_PEERS 16
#define T4R_PEER_HANDLE_BYTES 64
/* the same buffer on every rank of a group, as seen from THIS process: base[rank] is the local allocation,
 * the others are mappings obtained with t4r_peer_open (in a single-process test they may all be local). */
typedef struct {
  const void* base[T4R_MAX_PEERS];
  int world;
  int rank;
} t4r_peer_ptrs;
/* Export the cudaMalloc'ed allocation that contains dev_ptr: handle_out (host, 64 bytes) is sent to the peers
 * (any byte transport), offset_out (host) is dev_ptr's offset inside that allocation. */
int t4r_peer_export(const void* dev_ptr, void* handle_out /*host*/, int64_t* offset_out /*host*/);
/* Map a peer's exported allocation into this process (cudaIpcOpenMemHandle, lazy peer access);
 * *mapped_out = the peer's dev_ptr as addressable here.  Not for handles of this same process. */
int t4r_peer_open(const void* handle /*host*/, int64_t offset, void** mapped_out /*host*/);
int t4r_peer_close(void* mapped, int64_t offset);
/* out[i] = row ids[i] of the sharded table (owner = min(id / rows_per_shard, world - 1)) for i < *count_dev
 * (all `cap` rows when NULL; with a count, rows from it up to the next multiple of 256 are zero and the rest of
 * the outputs is left untouched).  K % 4 == 0, K <= 1024.  pad_id (-1: none): that row is read
 * once per CTA and served from shared memory.  out_f32 [cap, K] and / or out_planes bf16 [2, cap, round_up64(K)].
 * err_flag (optional) is set to 1 on ids outside [0, V) (their rows are zero). */
int t4r_peer_gather_rows(const t4r_peer_ptrs* shards /*host*/, int64_t V, int64_t rows_per_shard, int K,
                         const int64_t* ids, const int32_t* count_dev, int cap, int64_t pad_id, float* out_f32,
                         void* out_planes, int32_t* err_flag, void* stream);
/* Pull the label rows of every rank into one compact operand: mail_x->base[r] = fp32 [cap, K] rows of rank r,
 * mail_y->base[r] = int64 [cap] labels, counts [world] (DEVICE, e.g. the output of a 4-byte all-gather) the
 * number of valid rows per rank.  Rank-major concatenation -> out_f32 [world*cap, K] (optional), out_planes
 * bf16 [2, world*cap, round_up64(K)] (optional), out_labels [world*cap]; *t_total = sum(counts),
 * *my_start = first row of this rank (optional); rows up to round_up(t_total, 256) are zeroed. */
int t4r_peer_pull_rows(const t4r_peer_ptrs* mail_x /*host*/, const t4r_peer_ptrs* mail_y /*host*/,
                       const int32_t* counts, int cap, int K, float* out_f32, void* out_planes, int64_t* out_labels,
                       int32_t* t_total, int32_t* my_start, void* stream);
/* stats->base[s] = fp32 [3, cap_g] of shard s: row log-sum-exp over the shard's classes | label logit (0 when
 * the label lives elsewhere) | int32 count of the shard's classes scoring above the label.  row_loss [cap_g] =
 * logsumexp_s(lse_s) - sum_s(tgt_s) for rows < *t_total (0 beyond), row_rank [cap_g] = sum_s(count_s) when
 * with_rank, loss (optional) = mean of row_loss over the *t_total rows. */
int t4r_peer_combine_lse(const t4r_peer_ptrs* stats /*host*/, int64_t cap_g, const int32_t* t_total, int with_rank,
                         float* row_loss, int32_t* row_rank, float* loss, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T4R_B200_H_ */
