/* panst3r_b200 — C ABI of the Blackwell (sm_100a) PanSt3R inference hot path.
 *
 * Drop-in boundary for naver/panst3r's multi-view forward path.  Every entry point is plain C:
 * raw device pointers + sizes, a CUDA stream handle, int return code (0 = ok, <0 = error, message via
 * pst3r_last_error()).  The caller (PyTorch host code, see panst3r_b200/ops.py) owns all memory; the
 * library never allocates device memory and never synchronises the stream.
 *
 * Reference interfaces replaced (paths relative to the naver/panst3r checkout):
 *   - curope `rope_2d(tokens, positions, base, fwd)` (croco/curope, installed per README.md:67-71;
 *     selected by 'RoPE100' at src/panst3r/model/input_mixer.py:16)            -> pst3r_rope2d
 *   - `nn.Linear` / croco `Mlp` / `nn.MultiheadAttention` in-proj GEMMs
 *     (model/upscalers/pixel_shuffle.py:17-27,40-54; model/mask_transformer.py:314,372,435-437;
 *     model/input_mixer.py:14,19; upstream encoder/decoder blocks driven from engine/must3r.py:17-24,45,93)
 *                                                                               -> pst3r_gemm_bf16
 *   - xformers `memory_efficient_attention` / croco `Attention`,`CrossAttention` / torch MHA
 *     (gradio_panst3r.py:25; model/blocks.py:18-19,32; model/mask_transformer.py:395-398)
 *                                                                               -> pst3r_attention
 *   - `torch.einsum("bqc,bnchw->bnqhw")` (model/mask_transformer.py:279-288)   -> pst3r_gemm_bf16 with
 *     PST3R_STORE_TRANSPOSED (mask-logit planes) + pst3r_attn_mask_bits (:264-272, :172)
 *   - `F.pixel_shuffle` (model/upscalers/pixel_shuffle.py:42,46,50)             -> PST3R_STORE_PIXSHUF2
 *   - upstream LinearHead depth-to-space (pointmaps at engine/must3r.py:93)     -> PST3R_STORE_D2S
 *   - `nn.LayerNorm`, `nn.GroupNorm`, sine PE, DINO preprocessing (model/dino.py:61-66), LoftUp featuriser
 *     (model/upscalers/loftup.py:9-79)                                          -> pst3r_layernorm etc.
 */
#ifndef PANST3R_B200_H_
#define PANST3R_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pst3r_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
const char* pst3r_last_error(void);
/* ABI version of this header: bumped whenever a struct layout or an entry-point signature changes
 * (2: pst3r_gemm_epilogue grew the folded-LayerNorm fields; batched GEMM / LayerNorm, SM budget, post-processing;
 *  3: reference-precision "split bf16" operands (PST3R_KIND_SPLIT) through the GEMM and the row kernels, masked row
 *     softmax, TMA-store mask-logit planes).  Entry points ADDED without touching existing ones keep the version
 *     (pst3r_set_split_k); a library that lacks one fails at bind time, symbol by symbol (panst3r_b200/lib.py). */
#define PST3R_ABI_VERSION 3

/* Element kinds of a matrix argument.  PST3R_KIND_SPLIT is the reference-precision representation used for the
 * panoptic head, which the reference runs in fp32 (src/panst3r/panst3r.py:236-245): a value x is held as TWO bf16
 * numbers hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits, relative error 2^-17).  A split row stores its hi parts
 * followed, `lo_off` elements later, by its lo parts; unless an entry point says otherwise lo_off == the number of
 * logical columns, i.e. a row is [hi(0..C) | lo(0..C)] and its stride is >= 2*C bf16 elements. */
enum {
  PST3R_KIND_BF16 = 0,
  PST3R_KIND_F32 = 1,
  PST3R_KIND_SPLIT = 2
};
int pst3r_version(void);
/* Returns 0 if the current CUDA device is sm_100 (B200); <0 otherwise. */
int pst3r_check_device(void);
int pst3r_num_sms(void);
/* Limits the SMs the following launches may fill (persistent grids, tile / split heuristics); 0 = all.  Returns the
 * previous budget.  Host-side setting read at launch time: lets the latency-bound sequential memory build
 * (engine/must3r.py:40-54) and the throughput-bound DINOv2 encoder (model/dino.py:59-71) run side by side on
 * two streams, each on its own share of the 148 SMs. */
int pst3r_set_sm_budget(int32_t n_sms);
/* Programmatic dependent launch (every kernel is launched with the programmatic-stream-serialization attribute and
 * executes griddepcontrol.wait before its first global access) on / off; returns the previous setting.  Off for
 * per-kernel profiling: a dependent kernel that starts early spends the wait inside ITS measured duration. */
int pst3r_set_pdl(int32_t on);
/* Split-K for small GEMMs on / off (default off); returns the previous setting.  While on, an all-bf16 pst3r_gemm_bf16 whose
 * 128 x 64 output tiles fill at most half of the SMs and whose reduction has >= 8 k-blocks runs on two-CTA clusters, each CTA
 * over half of K, the partial tile reduced through distributed shared memory (csrc/gemm_splitk.cuh).  The fp32 summation order
 * differs from the unsplit kernel's, so results are equal to rounding, not bitwise; the decision depends on M.  The host side
 * switches it on around the sequential memory build (engine/must3r.py:40-54: a chain of M = 768 nn.Linear calls), whose
 * shapes are the same on every rank, and leaves the per-view stages on the unsplit kernels (bit-identical sharded runs). */
int pst3r_set_split_k(int32_t on);

/* ---- GEMM: C[M,N] = epilogue(A[M,K] * B[N,K]^T) ---------------------------------------------
 * A, B are bf16, K contiguous (row strides lda/ldb in elements, multiples of 8).  tcgen05 tensor cores,
 * fp32 accumulation in TMEM, TMA-fed.  The epilogue is described by pst3r_gemm_epilogue. */
enum {
  PST3R_ACT_NONE = 0,
  PST3R_ACT_GELU = 1, /* exact erf GELU (nn.GELU default) */
  PST3R_ACT_RELU = 2
};
enum {
  PST3R_STORE_PLAIN = 0,      /* out[row*ldo + col]; if rows_per_batch > 0:
                                 out[(row / rows_per_batch)*batch_stride + (row % rows_per_batch)*ldo + col] */
  PST3R_STORE_TRANSPOSED = 1, /* out[(row / rows_per_batch)*batch_stride + col*ldt + row % rows_per_batch] */
  PST3R_STORE_PIXSHUF2 = 2,   /* row=(b,y,x) on grid_h x grid_w, col=4c+2i+j -> out[((b*2gh+2y+i)*2gw+2x+j)*ldo + c] */
  PST3R_STORE_D2S = 3         /* row=(b,y,x), col=(i*P+j)*C+c -> fp32 out[((b*gh*P+y*P+i)*gw*P + x*P+j)*C + c] */
};

typedef struct pst3r_gemm_epilogue {
  void* out;              /* bf16 or fp32 device pointer */
  int64_t ldo;            /* row stride of out in elements (PLAIN / PIXSHUF2) */
  int32_t out_kind;       /* PST3R_KIND_*: bf16, fp32 or split bf16 (hi at column c, lo at column c + out_lo_off) */
  int32_t act;            /* PST3R_ACT_* (applied after bias) */
  const float* bias;      /* [N] fp32 or NULL */
  const float* col_scale; /* [N] fp32 or NULL (LayerScale; applied after act) */
  const void* residual;   /* [M, ldr] of res_kind (bf16 unless set) or NULL (added last) */
  int64_t ldr;
  int32_t res_mod_rows;   /* > 0: residual row = row % res_mod_rows (broadcast, e.g. position embeddings) */
  float alpha;            /* accumulator scale (applied first) */
  int32_t store_mode;     /* PST3R_STORE_* */
  int64_t rows_per_batch; /* TRANSPOSED */
  int64_t batch_stride;   /* TRANSPOSED */
  int64_t ldt;            /* TRANSPOSED */
  int32_t grid_h, grid_w; /* PIXSHUF2 / D2S token grid */
  int32_t d2s_patch;      /* D2S patch size P */
  int32_t d2s_ch;         /* D2S channels C */
  /* fused 2-D RoPE on the leading rope_cols columns (head_dim 64, pairs (j, j+16) in each 32-wide half;
   * first half rotates by y, second by x) — replaces cuRoPE2D after the QKV projection */
  const float* rope_cs;   /* [rope_maxpos][16][2] fp32 (cos, sin) or NULL */
  const int32_t* rope_pos; /* [M][2] (y, x) */
  int32_t rope_cols;
  int32_t rope_maxpos;
  /* LayerNorm folded into the GEMM (nn.LayerNorm directly ahead of an nn.Linear: norm1/norm2/norm3 of every croco /
   * MUSt3R / DINOv2 block).  A holds the RAW rows x [M, K]; B holds gamma-scaled weights bf16(gamma[k] * W[n][k]);
   * bias holds bias + W beta.  The epilogue computes rstd[m] * (acc - mean[m] * ln_colsum[n]) + bias[n], with
   * mean / rstd of row m from ln_stats[m][0 .. K/32) = per-32-column (sum x, sum x^2) partial sums (float2) written
   * by the GEMM that produced x (its stats_out).  ln_colsum[n] = sum_k B[n][k] (fp32). */
  const void* ln_stats;    /* float2 [M][K/32] or NULL */
  int32_t ln_slots;        /* K / 32 */
  const float* ln_colsum;  /* [N] */
  float ln_eps;
  /* Producer side: float2 [M][N/32] partial sums of the bf16 values this GEMM stores (PLAIN bf16 store, N % 32 == 0). */
  void* stats_out;
  /* Reference-precision mode (fp32 policy of the panoptic head, panst3r.py:236-245, on bf16 tensor cores):
   *   split_terms = 0: A, B plain bf16.
   *   split_terms = 3: A and B are PST3R_KIND_SPLIT; acc = A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T (fp32 in TMEM).
   *   split_terms = 2: A plain bf16 (exactly representable inputs, e.g. the bf16 trunk features), B split.
   * a_lo_off / b_lo_off: element distance between the hi and the lo part inside a row of A / B (multiples of 8);
   * the part is a dimension of the TMA tensor maps.  K stays the logical reduction length. */
  int32_t split_terms;
  int64_t a_lo_off, b_lo_off;
  int64_t out_lo_off;     /* out_kind == PST3R_KIND_SPLIT */
  int32_t res_kind;       /* PST3R_KIND_* of residual */
  int64_t res_lo_off;
} pst3r_gemm_epilogue;

int pst3r_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N, int32_t K,
                    const pst3r_gemm_epilogue* epi, pst3r_stream_t stream);

/* Strided-batched variant: problem b (0 <= b < batches) computes epi(A_b * B_b^T) with A_b = A + b*a_batch_stride,
 * B_b = B + b*b_batch_stride (elements), out_b = epi->out + b*out_batch_stride, bias_b = epi->bias + b*bias_batch_stride.
 * One launch covers all problems (3-D TMA maps; the tile scheduler walks (batch, m, n)).  PLAIN store with bias /
 * activation only.  Used for the per-layer K|V projections of freshly appended memory tokens: 12 decoder layers,
 * one launch (upstream MUSt3R memory update driven from engine/must3r.py:45). */
int pst3r_gemm_bf16_batched(const void* A, int64_t lda, int64_t a_batch_stride, const void* B, int64_t ldb,
                            int64_t b_batch_stride, int32_t M, int32_t N, int32_t K, int32_t batches,
                            const pst3r_gemm_epilogue* epi, int64_t out_batch_stride, int64_t bias_batch_stride,
                            pst3r_stream_t stream);

/* 3x3 convolution, stride 1, zero padding 1, over a pixel-major bf16 map x [V, H, W, ldx] (C <= ldx valid channels)
 * as an implicit GEMM: out[(v, y, x), o] = epi(sum_{ky,kx,c} x[v, y+ky-1, x+kx-1, c] * w[o, (ky*3+kx)*cpad + c]).
 * w is bf16 [O, 9*cpad] (Conv2d weight [O, C, 3, 3] permuted to tap-major and zero-padded to cpad, cpad % 64 == 0).
 * The A operand is fetched by a 4-D TMA map whose out-of-bounds zero fill IS the padding.  PLAIN store only
 * (out rows = V*H*W pixels).  Replaces nn.Conv2d(k=3, padding=1) at model/upscalers/loftup.py:124,127. */
int pst3r_conv3x3_nhwc(const void* x, int64_t ldx, int32_t V, int32_t H, int32_t W, int32_t C, const void* w,
                       int32_t cpad, int32_t O, const pst3r_gemm_epilogue* epi, pst3r_stream_t stream);

/* ---- Attention: O = softmax(scale * Q K^T [+ mask]) V ----------------------------------------
 * bf16 Q/K/V with head_dim 64 or 96, element strides given per (batch, token, head); head_dim contiguous.
 * kv_batch_stride may be 0 (all batch items attend the same memory tokens: MUSt3R render pass).
 * mask_bits (optional): uint32 words, bit (k & 31) of word [b][q][k >> 5] set => key k is BLOCKED for
 * query q (shared by all heads, mask_transformer.py:272).
 * O is written bf16 as [B, Nq, H*hd] with row stride ldo.
 * workspace: fp32 scratch of pst3r_attention_workspace_bytes() bytes (split-KV partials), may be NULL when 0. */
typedef struct pst3r_attn_args {
  const void* q; int64_t q_sb, q_sn, q_sh;
  const void* k; int64_t k_sb, k_sn, k_sh;
  const void* v; int64_t v_sb, v_sn, v_sh;
  void* o; int64_t o_sb, o_sn; /* o[b*o_sb + n*o_sn + h*hd + d] */
  int32_t B, H, Nq, Nk, head_dim;
  float scale;
  const uint32_t* mask_bits; int64_t mask_sb, mask_sq; /* strides in words */
  int32_t kv_splits;     /* 0 = auto */
  void* workspace; int64_t workspace_bytes;
} pst3r_attn_args;

int64_t pst3r_attention_workspace_bytes(int32_t B, int32_t H, int32_t Nq, int32_t head_dim, int32_t kv_splits);
int32_t pst3r_attention_auto_splits(int32_t B, int32_t H, int32_t Nq, int32_t Nk);
int pst3r_attention(const pst3r_attn_args* args, pst3r_stream_t stream);

/* ---- Normalisation / elementwise ------------------------------------------------------------- */
/* y = LN(x [+ add]) * gamma + beta ; x, add, y, sum_out of any PST3R_KIND_* (row strides in elements of the stored
 * type; split rows are [hi(dim) | lo(dim)]).
 * If sum_out != NULL the pre-norm sum (x + add) is also written (row stride ld_sum): fused
 * "residual add + post-norm" of the Mask2Former-style query decoder (mask_transformer.py:339-340).
 * If x_rows_per_batch > 0, input row r lives at x + (r / rpb)*x_batch_stride + (r % rpb)*ldx (used to drop the
 * DINOv2 CLS token while normalising, model/dino.py:69). */
int pst3r_layernorm(const void* x, int32_t x_kind, int64_t ldx, const void* add, int32_t add_kind, int64_t ld_add,
                    const float* gamma, const float* beta, float eps, void* y, int32_t y_kind, int64_t ldy, void* sum_out,
                    int32_t sum_kind, int64_t ld_sum, int32_t rows, int32_t dim, int32_t x_rows_per_batch,
                    int64_t x_batch_stride, pst3r_stream_t stream);

/* Batched LayerNorm over `batches` equally shaped bf16 matrices with per-batch parameters:
 *   y[b][r] = LN(x[b][r] + add[r]) * gamma[b] + beta[b],   x[b] = x + b*x_batch_stride, y[b] = y + b*y_batch_stride,
 *   gamma[b] = gamma + b*param_batch_stride (same for beta); `add` (may be NULL) is shared by all batches.
 * One launch normalises the 12 per-layer memory-token sets of a MUSt3R memory update (norm_y of
 * layer input + feedback, memory_mode='norm_y', configs/base.yaml:12-15). */
int pst3r_layernorm_batched(const void* x, int64_t ldx, int64_t x_batch_stride, const void* add, int64_t ld_add,
                            const float* gamma, const float* beta, int64_t param_batch_stride, float eps, void* y,
                            int64_t ldy, int64_t y_batch_stride, int32_t rows_per_batch, int32_t batches, int32_t dim,
                            pst3r_stream_t stream);

/* In-place 2-D RoPE, curope semantics: tokens bf16 [B, N, H, D] (D contiguous, D % 4 == 0),
 * positions int32 [B, N, 2] = (y, x); angle = p * base^(-j/(D/4)); fwd = +1 / -1. */
int pst3r_rope2d(void* tokens, int64_t s_b, int64_t s_n, int64_t s_h, const int32_t* pos, int32_t B, int32_t N,
                 int32_t H, int32_t D, float base, float fwd, pst3r_stream_t stream);

/* out[r, c] = a[r, c] + b[r % b_rows, c]  (any PST3R_KIND_*; pos-embedding / level-embedding / query-embedding adds) */
int pst3r_add_bcast(const void* a, int32_t a_kind, int64_t lda, const void* b, int32_t b_kind, int64_t ldb, int32_t b_rows,
                    void* out, int32_t out_kind, int64_t ldo, int32_t rows, int32_t cols, pst3r_stream_t stream);

/* fp32 [rows, cols] (ld) -> bf16, and back */
int pst3r_cast_f32_to_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t cols,
                           pst3r_stream_t stream);
int pst3r_cast_bf16_to_f32(const void* x, int64_t ldx, float* y, int64_t ldy, int32_t rows, int32_t cols,
                           pst3r_stream_t stream);
/* y = x between any two PST3R_KIND_* (e.g. fp32 -> split bf16 when fp32 features enter the reference-precision head) */
int pst3r_convert(const void* x, int32_t x_kind, int64_t ldx, void* y, int32_t y_kind, int64_t ldy, int32_t rows,
                  int32_t cols, pst3r_stream_t stream);

/* Masked row softmax of the reference-precision attention (the query decoder's nn.MultiheadAttention,
 * mask_transformer.py:314,372,395-398, evaluated as S = QK^T GEMM -> this -> PV GEMM on split operands):
 * out[r][k] = softmax over the unblocked keys of S[r][0..Nk) (fp32, scale already applied); mask_bits as in
 * pst3r_attention, row q = r % Q (shared by all heads), NULL = no mask.  out of kind out_kind, row stride ldo; a split
 * output keeps its lo parts out_lo_off (>= Nk, a multiple of 8 when it feeds the P V GEMM) elements after the hi parts. */
int pst3r_softmax_rows(const float* S, int64_t lds, int32_t rows, int32_t Nk, const uint32_t* mask_bits, int64_t mask_sq,
                       int32_t Q, void* out, int32_t out_kind, int64_t ldo, int64_t out_lo_off, pst3r_stream_t stream);

/* Patchify (im2col) for the ViT patch embeddings: img fp32 [B,3,H,W] -> bf16 [B*(H/P)*(W/P), ldo] with
 * column = c*P*P + i*P + j (Conv2d weight flattening).  Columns [3*P*P, ldo) are zero-filled. */
int pst3r_patchify(const float* img, int32_t B, int32_t H, int32_t W, int32_t P, void* out, int64_t ldo,
                   pst3r_stream_t stream);

/* DINOv2 preprocessing fused with patchify (model/dino.py:61-66): x*0.5+0.5, ImageNet normalise, bilinear
 * resize (align_corners=False) to (Ho, Wo), then im2col with patch P (14). */
int pst3r_dino_preprocess_patchify(const float* img, int32_t B, int32_t H, int32_t W, int32_t Ho, int32_t Wo,
                                   int32_t P, void* out, int64_t ldo, pst3r_stream_t stream);

/* Mean of the centre 2x2 of every 8x8 cell of a pixel-major feature map: feats bf16 [B, Hm, Wm, C] ->
 * bf16 [B, Hm/8, Wm/8, C] (kind: PST3R_KIND_BF16, or PST3R_KIND_SPLIT with [hi(C) | lo(C)] pixel rows in and out).
 * The 8x bilinear downsample (align_corners=False) of the mask logits
 * (mask_transformer.py:286) is linear in the features, so the attention mask only needs these. */
int pst3r_center_pool8(const void* feats, int32_t kind, int32_t B, int32_t Hm, int32_t Wm, int32_t C, void* out,
                       pst3r_stream_t stream);

/* logits_t fp32 [Q, ld] (TRANSPOSED store: row q, column = flattened key token) -> mask bits
 * [Q, ceil(Nk/32)] with bit set iff logit < 0 (sigmoid < 0.5 => blocked); rows that would be fully blocked
 * are cleared (mask_transformer.py:172). */
int pst3r_attn_mask_bits(const float* logits_t, int64_t ld, int32_t Q, int32_t Nk, uint32_t* bits,
                         pst3r_stream_t stream);

/* L2-normalise rows: y = x / (||x|| + eps)  (fp32 in, y of any PST3R_KIND_*; mask_transformer.py:227) */
int pst3r_l2norm_rows(const float* x, int64_t ldx, void* y, int32_t y_kind, int64_t ldy, int32_t rows, int32_t cols,
                      float eps, pst3r_stream_t stream);

/* bf16 / split-bf16 pixel-major [B, HW, C] -> fp32 channel-major [B, C, HW] (reference NCHW layout of mask_feats / fpn) */
int pst3r_nhwc_to_nchw_f32(const void* x, int32_t x_kind, int32_t B, int32_t HW, int32_t C, float* y, pst3r_stream_t stream);

/* ---- Panoptic post-processing front half (engine/postprocess.py:18-27, 38-45, 63-120) ------------------
 * scores[q] = max_k sigmoid(logits[q][k]), labels[q] = first arg max (postprocess.py:39). */
int pst3r_class_scores(const float* logits, int64_t ldl, int32_t Q, int32_t K, float* scores, int32_t* labels,
                       pst3r_stream_t stream);
/* Fused sigmoid -> bilinear resize (align_corners=False, postprocess.py:24-25) -> score-weighted argmax over the
 * kept queries (:63, :77) for V equally shaped views.  masks fp32 [V][Q][hm][wm] (strides in elements); keep_idx /
 * keep_scores [nkeep]: surviving query indices (ascending) and their class scores.  Per output pixel:
 * ids = position in keep_idx of the winning query (first on ties), win = that query's mask probability.
 * area_half[k] += #pixels with probability >= 0.5 (:84), area_won[k] += #pixels query k wins with probability >=
 * mask_threshold (:85-86); the caller zeroes both before the first view group of a round. */
int pst3r_panoptic_argmax(const float* masks, int64_t view_stride, int64_t query_stride, int32_t V, int32_t hm, int32_t wm,
                          const int32_t* keep_idx, const float* keep_scores, int32_t nkeep, int32_t H, int32_t W,
                          float mask_threshold, int32_t* ids, float* win, int64_t out_view_stride, int32_t out_row_stride,
                          int32_t* area_half, int32_t* area_won, pst3r_stream_t stream);
/* The same kernel on a BAND of the map: output rows [y0, y0 + rows) of every view, with `masks` holding only the source
 * rows [src_row0, src_row0 + src_rows) of every plane (plane pitch query_stride >= src_rows * wm; hm stays the height of
 * the whole mask grid, ids / win still point at row 0 of the whole output map).  The band must contain every source row
 * the bilinear resize reads for its output rows (checked).  This is what lets a caller produce the mask logits chunk by
 * chunk into a scratch buffer that stays in the 126 MB L2 and never materialise the (V, Q, h, w) tensor the reference
 * builds at postprocess.py:18-27 (panst3r_b200/postprocess.py, LazyMasks); counters accumulate across bands. */
int pst3r_panoptic_argmax_band(const float* masks, int64_t view_stride, int64_t query_stride, int32_t V, int32_t hm, int32_t wm,
                               int32_t src_row0, int32_t src_rows, const int32_t* keep_idx, const float* keep_scores,
                               int32_t nkeep, int32_t H, int32_t W, int32_t y0, int32_t rows, float mask_threshold,
                               int32_t* ids, float* win, int64_t out_view_stride, int32_t out_row_stride,
                               int32_t* area_half, int32_t* area_won, pst3r_stream_t stream);
/* pan[i] = lut[ids[i]] if win[i] >= mask_threshold else 0; conf[i] = win[i] where pan[i] != 0 else void_confidence
 * (:103-105); lut[k] = segment id of kept query k or 0 if it was filtered out. */
int pst3r_panoptic_finalize(const int32_t* ids, const float* win, const int32_t* lut, int32_t nkeep, float mask_threshold,
                            float void_confidence, int32_t* pan, float* conf, int64_t n, pst3r_stream_t stream);

/* ---- LoftUp guidance path (model/upscalers/loftup.py:9-79,122-130,152-157) ------------------------ */
int64_t pst3r_loftup_workspace_bytes(int32_t V, int32_t C, int32_t groups);
/* img fp32 [V,3,H,W] -> half fp32 [V,3,H/2,W/2] (bilinear x0.5) and minmax fp32 [3][2] = per-channel (min, max)
 * over the WHOLE batch (MinMaxScaler, loftup.py:14-19). */
int pst3r_loftup_guidance(const float* img, int32_t V, int32_t H, int32_t W, float* half, float* minmax,
                          void* workspace, pst3r_stream_t stream);
/* MinMaxScaler -> ImplicitFeaturizer (n_freqs sin/cos features of [gy, gx, r, g, b] + scaled rgb; channel order
 * [sin(f*5+m) | cos(f*5+m) | rgb]) -> GroupNorm(1, C) with (gamma, beta) -> bf16 pixel-major out [V, Hh, Wh, ldo]
 * (channels [C, ldo) zeroed; out_kind PST3R_KIND_SPLIT: pixel rows [hi(ldo) | lo(ldo)]).  gy/gx = torch.linspace(-1,1,.) tables, freqs = exp(linspace(-2,10,n_freqs)),
 * biases = the flat (2, 5, n_freqs) parameter. */
int pst3r_loftup_fourier_gn(const float* half, const float* minmax, const float* gy, const float* gx,
                            const float* freqs, const float* biases, int32_t V, int32_t Hh, int32_t Wh,
                            int32_t n_freqs, const float* gamma, const float* beta, float eps, void* out, int32_t out_kind,
                            int64_t ldo, void* workspace, pst3r_stream_t stream);
/* In-place GroupNorm (+ optional ReLU) on a pixel-major map x [V, npix, C] (nn.GroupNorm(groups, C)); kind:
 * PST3R_KIND_BF16, or PST3R_KIND_SPLIT with [hi(C) | lo(C)] pixel rows. */
int pst3r_groupnorm_nhwc(void* x, int32_t kind, int32_t V, int32_t npix, int32_t C, int32_t groups, const float* gamma,
                         const float* beta, float eps, int32_t relu, void* workspace, pst3r_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PANST3R_B200_H_ */
