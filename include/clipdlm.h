/* clipdlm.h — C-ABI of libclipdlm.so: the B200 (sm_100a) implementation of the CLIP-Diffusion-LM hot path.
 *
 * The reference (xu-shitong/diffusion-image-captioning, CLIP-DDPM.py) has no FFI: its boundary is a set of
 * Python functions over torch tensors.  Every entry point below therefore cites the reference function (or the
 * third-party op that function dispatches to) that it replaces; the Python host in
 * diffusion-image-captioning_b200/ keeps the reference's names/signatures and binds these symbols with ctypes
 * (see INTEGRATION.md).
 *
 * Conventions: plain C types only; every pointer is a DEVICE pointer owned by the caller unless stated; the
 * library allocates nothing on the hot path (workspace is caller-provided); all work is enqueued on the given
 * cudaStream_t (passed as void*); return 0 = ok, <0 = error with text in clipdlm_last_error() (thread-local).
 * "bf16 pair" (hi, lo): lo == NULL means plain bf16 storage; lo != NULL means split storage value = hi + lo
 * ("bf16x3" precision mode: GEMMs run three tensor-core passes hi*hi + lo*hi + hi*lo, fp32-class accuracy).
 */
#ifndef CLIPDLM_H
#define CLIPDLM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* clipdlm_stream; /* cudaStream_t */

const char* clipdlm_last_error(void);
int clipdlm_version(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM present), else 0; <0 on CUDA error. */
int clipdlm_device_ok(void);

/* ------------------------------------------------------------------------------------------------------------
 * tcgen05 GEMM   D[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate in TMEM), fused epilogues.
 * Replaces: nn.Linear forward (HF modeling_distilbert.py:187-189,206,224-226,514; CLIP-DDPM.py:299,323) and the
 * autograd dgrad / wgrad GEMMs of l.backward() (CLIP-DDPM.py:483).
 * a_major/b_major: 0 = operand stored row-major [M|N][K] (K contiguous); 1 = stored [K][M|N] (MN contiguous).
 * ---------------------------------------------------------------------------------------------------------- */
enum {
  CLIPDLM_EPI_STORE = 0,  /* out = f(acc): +bias, dropout, +residual, gelu dual-store, *gelu'(u); bf16 pair / f32 */
  CLIPDLM_EPI_WGRAD = 1,  /* acc_f32[M,N] += acc   (split-K, fp32 red.add) */
  CLIPDLM_EPI_LSE = 2,    /* per (row, 128-col half tile) max / sum-exp / argmax partials + target logit; no logits in HBM */
  CLIPDLM_EPI_SMGRAD = 3, /* out = (exp(acc - lse[m]) - [n == tgt[m]]) * scale   (softmax-CE gradient, bf16 pair) */
  /* Factored softmax gradient (plain bf16; see clipdlm_ce_row_terms): d x = sum_v softmax_v W_v - W_tgt, and
   * softmax_v = exp(s_v - c) / sum_u exp(s_u - c) for ANY per-launch constant c, so the 1 / sum factor leaves the reduction over v:
   * the lm_head pass stores exp(s - c) (LSE_EXP), the gradient GEMM reads that array as it is and scales its accumulator ROWS
   * (STORE_ROWSCALE) - the pass over the [M, V] array that forms (softmax - onehot) * scale disappears. */
  CLIPDLM_EPI_LSE_EXP = 4,        /* as LSE without arg-max tracking, partials relative to the constant *exp_shift (part_max = shift,
                                   * part_sum = sum exp(acc - shift)); out_hi (required) receives bf16(exp(acc - shift)), exponent clamped
                                   * to 2^100 */
  CLIPDLM_EPI_STORE_ROWSCALE = 5, /* out = acc * row_scale[m] + residual  (plain bf16, 32-byte aligned rows, N % 256 == 0, K-major A) */
  /* GELU layer whose backward is a plain multiply (plain bf16, same alignment rules): the forward stores gelu'(u) where STORE's dual
   * mode stores u, the backward's epilogue multiplies by it instead of evaluating gelu' (two exp2 and a degree-6 polynomial per
   * element in the epilogue of a K = 3072 -> N = 3072 gradient GEMM that is epilogue-issue bound). */
  CLIPDLM_EPI_STORE_GELU_DERIV = 6, /* out = gelu'(acc + bias), out2 = gelu(acc + bias)   (K-major operands, bias required) */
  CLIPDLM_EPI_STORE_MULAUX = 7      /* out = acc * u   (u_hi = an array a STORE_GELU_DERIV forward wrote; MN-major B). If acc_f32 is set
                                     * (fp32 [N], 8-byte aligned, unscattered output): acc_f32[n] += sum_m out[m, n] of the bf16-rounded
                                     * outputs - the bias gradient of the layer in front, without a pass over out */
};

typedef struct clipdlm_gemm {
  const void* a_hi; const void* a_lo;
  const void* b_hi; const void* b_lo;
  int64_t lda, ldb;               /* row pitch in elements of the stored 2-D arrays */
  int32_t M, N, K;
  int32_t a_major, b_major;
  int32_t gather_len, gather_stride; /* >0: logical A row m = stored row (m / len) * stride + m % len  (K-major A only) */
  int32_t epilogue;
  int32_t k_splits;               /* WGRAD only; 0 = auto */
  /* STORE */
  void* out_hi; void* out_lo; float* out_f32; int64_t ldo;
  void* out2_hi; void* out2_lo;   /* if set: out = acc+bias (pre-activation), out2 = gelu(out) */
  const float* bias;              /* [N] or NULL */
  const void* res_hi; const void* res_lo; int64_t ldr; /* residual (same row mapping as out) or NULL */
  const void* u_hi; const void* u_lo; int64_t ldu;     /* if set: out = acc * gelu'(u) */
  int32_t scatter_len, scatter_stride; /* >0: output/residual row = (m / len) * stride + m % len */
  uint64_t drop_seed; uint32_t drop_site; float drop_p; /* dropout on (acc + bias) before the residual; p = 0 off */
  /* WGRAD */
  float* acc_f32;                 /* [M, ldo] fp32, accumulated in place */
  /* LSE / SMGRAD */
  float* part_max; float* part_sum; int32_t* part_arg; /* [2 * ceil(N/256)][M]: slot 2 * tile + half; part_arg NULL = no arg-max tracking */
  float* tgt_logit;               /* [M] */
  const int32_t* targets; int32_t tgt_period; /* target of row m = targets[m % tgt_period] */
  const float* lse;               /* [M] (SMGRAD) */
  float grad_scale;               /* SMGRAD (part_max, unused by this epilogue otherwise, may point to ONE device float multiplying it: NULL = 1) */
  const float* exp_shift;         /* LSE_EXP: device scalar c (natural-log units), NULL = 0 */
  const float* row_scale;         /* STORE_ROWSCALE: [M] fp32, indexed by the GEMM row m (before scatter) */
} clipdlm_gemm_t;

int clipdlm_gemm(const clipdlm_gemm_t* g, clipdlm_stream stream);
/* Debug hook for bring-up: override the MN-major smem descriptor strides (bytes); 0,0 restores defaults. */
void clipdlm_gemm_debug_mn_desc(uint32_t lbo_bytes, uint32_t sbo_bytes);
/* Debug hook for performance triage (tools/gemm_perf.py): bit 0 = drop the epilogue's global stores, bit 1 = drop its auxiliary
 * operand loads (residual / gelu' input), bit 2 = drain TMEM only (no epilogue math), bit 3 = force single-CTA tiles (cta_group::1), bit 4 = issue no MMAs, bit 5 = issue no TMA loads, bit 6 = use the generic STORE epilogue where a specialised one exists, bit 9 = launch GEMMs without programmatic dependent launch, bit 10 = row-wise global stores instead of TMA stores in the specialised epilogue, bit 12 = no band tile order in the lm_head passes, bit 13 = experiment: the specialised epilogues read the bias straight from global memory instead of the per-tile shared-memory array (two epilogue-wide barriers per tile). 0 restores normal operation. */
void clipdlm_gemm_debug_flags(uint32_t flags);

/* Reduce LSE partials: lse[m], argmax[m] (may be NULL), and adds sum_m(lse[m] - tgt_logit[m]) * scale to *loss_acc (double).
 * Replaces softmax -> gather -> log -> sum -> mean (CLIP-DDPM.py:436-437) and softmax/argmax (CLIP-DDPM.py:620). */
int clipdlm_lse_combine(const float* part_max, const float* part_sum, const int32_t* part_arg, int32_t n_tiles, int32_t M,
                        const float* tgt_logit, float* lse, int32_t* argmax, double* loss_acc, double scale,
                        clipdlm_stream stream);

/* Row terms of the factored softmax-CE gradient, after clipdlm_lse_combine produced lse[m] from LSE_EXP partials:
 *   row_scale[m] = scale * exp(*exp_shift - lse[m])            (= scale / sum_v exp(s_v - shift); 0 if that is not finite)
 *   dx[row(m), 0..D) -= scale * W[targets[m % tgt_period], 0..D)   (the one-hot term; dx bf16, read-modify-write; W = the bf16
 *                                                                    lm_head operand of the gradient GEMM, pitch ldw)
 * row(m) = (m / scatter_len) * scatter_stride + m % scatter_len when scatter_len > 0, else m. D % 8 == 0, 16-byte aligned rows.
 * Replaces, together with LSE_EXP / STORE_ROWSCALE, softmax -> gather -> log -> backward of CLIP-DDPM.py:436-437 for the lm_head rows. */
int clipdlm_ce_row_terms(const float* lse, const float* exp_shift, const int32_t* targets, int32_t tgt_period, float scale, int32_t M,
                         const void* w_bf16, int64_t ldw, void* dx_bf16, int64_t ldx, int32_t scatter_len, int32_t scatter_stride, int32_t D,
                         float* row_scale, clipdlm_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * HBM-bound kernels
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct clipdlm_bf { void* hi; void* lo; } clipdlm_bf_t;

/* K1-K4: embedding gather + q_sample + CLIP concat/add fusion + segment/position add + embedding LayerNorm (+dropout).
 * Replaces model.embedding (CLIP-DDPM.py:459), diffuse_t (:347-362), forward's hstack/segment add (:296-307) and HF
 * Embeddings.forward (modeling_distilbert.py:96-122).
 * mode 0: x = x_in[r, p, :] (fp32 [R, Ltxt, D]);  mode 1: x = ca[r / B] * E[ids[b, p]] + cb[r / B] * noise[b, p] with b = r % B.
 * Row r (0..R), position p (0..L): concat: p < Ltxt text, p == Ltxt image proj, p == Ltxt+1 text proj, + seg[p >= Ltxt] + pos[p];
 * add-fusion (fusion == 1, L == Ltxt): x + img_proj[b] (+ txt_proj[b] if guided) + pos[p].
 * Writes z (pre-LN) and h (post-LN, post-dropout). */
typedef struct clipdlm_embed {
  int32_t R, B, Ltxt, L, D, fusion, mode, guided;
  const float* x_in;            /* mode 0: fp32, row r at x_in + r * x_in_stride, position p at + p * D */
  int64_t x_in_stride;          /* elements between consecutive rows; 0 = Ltxt * D (dense) */
  const float* emb_table; const int32_t* ids; const float* noise; const float* coef_a; const float* coef_b; /* mode 1 */
  const float* img_proj; const float* txt_proj; /* [B, D] fp32 (image_linear / text_linear outputs) */
  const float* seg; const float* pos;           /* [2, D], [max_pos, D] */
  const float* ln_w; const float* ln_b; float ln_eps;
  clipdlm_bf_t z; clipdlm_bf_t h;               /* [R*L, D] */
  uint64_t drop_seed; uint32_t drop_site; float drop_p;
} clipdlm_embed_t;
int clipdlm_embed_fwd(const clipdlm_embed_t* e, clipdlm_stream stream);
/* Backward of the fusion (after LN bwd produced dz): accumulates d pos [L rows], d seg [2], d img_proj / d txt_proj [B, D] (fp32, +=). */
int clipdlm_embed_bwd(const clipdlm_bf_t* dz, int32_t R, int32_t B, int32_t Ltxt, int32_t L, int32_t D, int32_t fusion, int32_t guided,
                      float* d_pos, float* d_seg, float* d_img_proj, float* d_txt_proj, clipdlm_stream stream);

/* LayerNorm over the last dim (HF nn.LayerNorm eps=1e-12; modeling_distilbert.py:120,257,261,516). y = LN(z)*w + b (+dropout). */
int clipdlm_layernorm_fwd(const clipdlm_bf_t* z, const float* w, const float* b, float eps, int64_t rows, int32_t D,
                          const clipdlm_bf_t* y, float* y_f32 /* optional fp32 copy of y, may be NULL */,
                          uint64_t drop_seed, uint32_t drop_site, float drop_p, clipdlm_stream stream);
/* dz = LN'(z) applied to dy (dy masked by the output dropout if drop_p_out > 0); dw, db += (fp32 atomics);
 * optional dz_drop = dz * mask_in / (1 - p_in) (gradient entering a dropout that preceded the residual add);
 * optional gelu_u: dz *= gelu'(u) (MLM transform head, z = gelu(u)); optional dbias[D] += column sums of
 * (dz_drop if given else dz) = the bias gradient of the Linear that produced z's non-residual branch. */
int clipdlm_layernorm_bwd(const clipdlm_bf_t* z, const clipdlm_bf_t* dy, const float* w, float eps, int64_t rows, int32_t D,
                          const clipdlm_bf_t* dz, float* dw, float* db,
                          uint64_t drop_seed, uint32_t drop_site_out, float drop_p_out,
                          const clipdlm_bf_t* dz_drop, uint32_t drop_site_in, float drop_p_in,
                          const clipdlm_bf_t* gelu_u, float* dbias, clipdlm_stream stream);

/* Multi-head self-attention over L <= 128 positions, one warp per (row, head); qkv [R*L, 3*D] (q | k | v).
 * Replaces DistilBertSelfAttention / sdpa (modeling_distilbert.py:126-151,177-207). keymask[r] bit j = key j visible
 * (CLIP-DDPM.py:296-297 hstack([mask, 1, 0])); for L > 32 keymask is [R][ceil(L/32)] words. */
int clipdlm_attn_fwd(const clipdlm_bf_t* qkv, const uint32_t* keymask, int32_t R, int32_t L, int32_t D, int32_t H,
                     const clipdlm_bf_t* ctx, uint64_t drop_seed, uint32_t drop_site, float drop_p, clipdlm_stream stream);
int clipdlm_attn_bwd(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int32_t R, int32_t L, int32_t D,
                     int32_t H, const clipdlm_bf_t* dqkv, uint64_t drop_seed, uint32_t drop_site, float drop_p,
                     clipdlm_stream stream);

/* clipdlm_attn_bwd that may also form the bias gradients of the q|k|v projection (the column sums of dqkv the caller would otherwise compute
 * with clipdlm_colsum; autograd of the Linear at modeling_distilbert.py:187-189): dbias_qkv [3*D] fp32 (+=). *bias_folded = 1 when the kernel
 * that ran accumulated d(q bias) and d(v bias) there (plain bf16, L = 16 / 18: the packed tcgen05 kernel; d(k bias) is analytically zero -
 * every row of dS sums to zero - and is left untouched), 0 when the caller still has to run clipdlm_colsum over dqkv. */
int clipdlm_attn_bwd_bias(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int32_t R, int32_t L, int32_t D,
                          int32_t H, const clipdlm_bf_t* dqkv, uint64_t drop_seed, uint32_t drop_site, float drop_p, float* dbias_qkv,
                          int32_t* bias_folded, clipdlm_stream stream);

/* Test hook: attention path for plain bf16, L <= 32: 0 = default (L = 16 / 18: tcgen05 tiles of back-to-back packed sequences in both
 * directions; other L: forward tcgen05 with 32-row slots, backward mma.sync TMA ring), 1 = fp32 SIMT kernels, 2 = mma.sync TMA-ring
 * kernels, 3 = tcgen05 with 32-row slots for both directions, 4 = as 0 with the non-pipelined packed backward (A/B of the software pipeline). */
void clipdlm_attn_force_simt(int32_t on);

/* Column sums (bias gradients): out[n] += sum_m x[m, n]. */
int clipdlm_colsum(const clipdlm_bf_t* x, int64_t rows, int32_t N, float* out, clipdlm_stream stream);

/* Embedding-space loss between x_out[:, :Ltxt] and x_0[b] = E[ids[b]] (LOSS_FUNC, CLIP-DDPM.py:77-89,418,428), forward
 * value added to *loss_acc (double) and gradient written to dx [R*L, D] (rows p >= Ltxt zeroed).
 * kind: 0 series_sum_sample_mean, 1 series_sum, 2 mse_series_mean, 3 mse_series_sum. batch_size = reference BATCH_SIZE
 * (divisor of kinds 1 and 3); R_total = rows of the whole pass (divisor of the means) when called per chunk. */
int clipdlm_embed_loss(const clipdlm_bf_t* x_out, const float* emb_table, const int32_t* ids,
                       const float* target /* optional explicit target [target_rows, Ltxt, D] fp32 (x_tgt of the x_{t-1} objective,
                                              CLIP-DDPM.py:421): row r uses target[r % target_rows]; NULL = E[ids[r % B]] */,
                       int32_t target_rows, int32_t R, int32_t B, int32_t Ltxt,
                       int32_t L, int32_t D, int32_t kind, int64_t R_total, int32_t batch_size, float weight,
                       double* loss_acc, const clipdlm_bf_t* dx, clipdlm_stream stream);

/* Small fp32 linear for the CLIP projections (image_linear / text_linear, CLIP-DDPM.py:252-253,299): y[B,N] = x[B,K] W[N,K]^T + b. */
int clipdlm_small_linear_fwd(const float* x, const float* w, const float* b, int32_t B, int32_t K, int32_t N, float* y,
                             clipdlm_stream stream);
/* dW[N,K] += dy^T x ; db[N] += colsum(dy). */
int clipdlm_small_linear_bwd(const float* x, const float* dy, int32_t B, int32_t K, int32_t N, float* dw, float* db,
                             clipdlm_stream stream);

/* Flat multi-tensor AdamW (torch.optim.AdamW defaults, decoupled weight decay on every element; CLIP-DDPM.py:335,484),
 * refreshing the bf16 (pair) shadow copy the GEMMs read. grad_scale multiplies g (1/world_size after a sum all-reduce). */
int clipdlm_adamw(float* p, float* g, float* m, float* v, void* shadow_hi, void* shadow_lo, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                  int32_t zero_grad /* 1: g is cleared in the same pass (trainer.zero_grad(), CLIP-DDPM.py:471) */,
                  clipdlm_stream stream);
/* Data-parallel step fused with the optimizer (csrc/dp_fused.cu): reduce-scatter of the gradients, AdamW on this rank's slice,
 * all-gather of the updated fp32 weights + bf16 shadow - one kernel over NVLink / NVSwitch peer memory, replacing
 * ncclAllReduce(flat gradients) + clipdlm_adamw. p / g / shadow_* [r] are rank r's flat buffers as mapped into THIS process (symmetric
 * memory: same size and layout on every GPU); *_mc are the multicast (NVSwitch multimem) addresses of the same buffers, or NULL: then the
 * slice is summed with loads from every peer and written with stores to every peer. m / v are indexed like the flat buffer but only
 * the elements of this rank's slice (clipdlm_dp_slice) are touched, so a caller may pass (slice storage - slice begin). The caller
 * orders the kernel against the other ranks (a barrier over the group before and after). */
#define CLIPDLM_MAX_PEERS 16
typedef struct clipdlm_dp_buffers {
  int32_t rank, world;
  float* p[CLIPDLM_MAX_PEERS]; float* g[CLIPDLM_MAX_PEERS];
  void* shadow_hi[CLIPDLM_MAX_PEERS]; void* shadow_lo[CLIPDLM_MAX_PEERS]; /* shadow_lo[*] NULL unless split precision */
  float* p_mc; float* g_mc; void* shadow_hi_mc; void* shadow_lo_mc;
} clipdlm_dp_buffers_t;
/* element range [*begin, *end) of the flat buffer of n elements owned by `rank` (multiples of 8 elements) */
int clipdlm_dp_slice(int64_t n, int32_t rank, int32_t world, int64_t* begin, int64_t* end);
int clipdlm_adamw_dp(const clipdlm_dp_buffers_t* d, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     float weight_decay, int32_t step, float grad_scale /* 1 / world for a mean */, clipdlm_stream stream);
/* fp32 -> bf16 (pair) conversion (weights shadow refresh, frozen embedding table). */
int clipdlm_to_bf16(const float* x, void* hi, void* lo, int64_t n, clipdlm_stream stream);
/* bf16 (pair) -> fp32. */
int clipdlm_to_f32(const void* hi, const void* lo, float* y, int64_t n, clipdlm_stream stream);
/* y[r, :] = x[(r / len) * stride + r % len, :] gather of the first `len` of every `stride` rows, to fp32. */
int clipdlm_gather_rows_f32(const clipdlm_bf_t* x, int64_t rows_out, int32_t len, int32_t stride, int32_t D, float* y,
                            clipdlm_stream stream);
/* q_sample / diffuse_t (CLIP-DDPM.py:347-362) as a standalone op: out[s, :] = coef_a[s] * x0 + coef_b[s] * noise over n elements
 * (n = B * Ltxt * D, one noise draw shared by the S samples; coef_a = sqrt(alpha_bar[t]), coef_b = sqrt(1 - alpha_bar[t])).
 * The training path never materialises this tensor (clipdlm_embed_fwd mode 1 fuses it); this entry point serves diffuse_t(). */
int clipdlm_q_sample(const float* x0, const float* noise, const float* coef_a, const float* coef_b, int64_t n, int32_t S, float* out,
                     clipdlm_stream stream);
/* Key-visibility words for the attention kernels, keymask[r][ceil(L/32)]: text keys from attn_mask[r % B] (NULL = all
 * visible), image-CLIP key always visible, text-CLIP key visible iff guided (CLIP-DDPM.py:296-297). */
int clipdlm_keymask(const int32_t* attn_mask, int32_t R, int32_t B, int32_t Ltxt, int32_t L, int32_t fusion, int32_t guided,
                    uint32_t* keymask, clipdlm_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * TRAIN_EMBEDDING=True glue (CLIP-DDPM.py:238-243,292-293,319-320): the learned embedding is IN_CHANNEL (16) wide, features are fp32.
 * ---------------------------------------------------------------------------------------------------------- */
/* LOSS_FUNC (CLIP-DDPM.py:77-89) between y[:, :Ltxt] (y fp32 [R, L, ch], model feature_out) and target[r % target_rows] (fp32
 * [target_rows, Ltxt, ch]; x_0.repeat(...) :418,428 or x_tgt :421). kind / R_total / batch_size / weight / loss_acc as clipdlm_embed_loss.
 * dy (optional, fp32 [R, L, ch]) is WRITTEN: text positions = loss gradient + dce[(r * Ltxt + p) * ld_dce + c] (optional gradient coming
 * from the rounding cross-entropy), other positions = 0. d_target (optional, fp32 like target) -= loss gradient (atomic): with a learned
 * embedding the target x_0 / x_tgt carries gradient too. */
int clipdlm_feature_loss_f32(const float* y, const float* target, int32_t target_rows, int32_t R, int32_t Ltxt, int32_t L, int32_t ch,
                             int32_t kind, int64_t R_total, int32_t batch_size, float weight, double* loss_acc, const float* dce,
                             int32_t ld_dce, float* dy, float* d_target, clipdlm_stream stream);
/* Compact zero-padded GEMM operand: out[m, c] = c < ch ? y[(m / Ltxt) * L + m % Ltxt, c] : 0 for m < rows_out, c < ld; bf16 (pair). */
int clipdlm_pack_rows_bf16(const float* y, int64_t rows_out, int32_t Ltxt, int32_t L, int32_t ch, int32_t ld, void* hi, void* lo,
                           clipdlm_stream stream);
/* Classifier-free guidance on fp32 encoder outputs [R, row_len] (TRAIN_EMBEDDING path; CLIP-DDPM.py:313-317):
 * x_unguided[r] <- guided[r] ? (1 + w) x_guided[r] - w x_unguided[r] : unchanged. */
int clipdlm_row_mix_f32(float* x_unguided, const float* x_guided, const int32_t* guided, float w, int32_t R, int64_t row_len,
                        clipdlm_stream stream);
/* Backward of the mix: d_other[r] = guided[r] ? (1 + w) d_self[r] : 0 (gradient of the guided pass), then d_self[r] *= guided[r] ? -w : 1. */
int clipdlm_row_split_f32(float* d_self, float* d_other, const int32_t* guided, float w, int32_t R, int64_t row_len, clipdlm_stream stream);
/* a += b over n fp32 elements (sum of the input gradients of two encoder passes over the same input). */
int clipdlm_add_f32(float* a, const float* b, int64_t n, clipdlm_stream stream);
/* Gradient of nn.Embedding through q_sample: d_table[ids[j], c] += sum_s scale[s] * dx[s, j, c]; dx fp32 [S, tokens, ch], ids int32
 * [tokens], scale [S] (sqrt(alpha_bar[t_s]), CLIP-DDPM.py:360) or NULL (= 1). Replaces autograd through :459 and :347-362. */
int clipdlm_embedding_bwd(const float* dx, const float* scale, const int32_t* ids, int32_t S, int64_t tokens, int32_t ch, float* d_table,
                          clipdlm_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Engine: the composite hot path (model forward / loss+backward / denoise step) orchestrated natively.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct clipdlm_config {
  int32_t n_layers, dim, n_heads, hidden_dim, vocab, max_len, clip_dim, max_pos;
  int32_t fusion;     /* 0 = "concat", 1 = "add"  (CLIP_ADDING_METHOD, CLIP-DDPM.py:93-94) */
  int32_t precision;  /* 0 = bf16 tensor-core passes, 1 = split bf16x3 (parity mode) */
  float ln_eps;       /* 1e-12 */
  float dropout, attn_dropout; /* DistilBertConfig defaults 0.1 / 0.1 */
} clipdlm_config_t;

typedef struct clipdlm_engine clipdlm_engine_t;

/* Flat parameter layout (fp32 master / grads / Adam moments / bf16 shadows share it). Slot ids below;
 * per-layer slots are CLIPDLM_P_LAYER0 + layer * CLIPDLM_P_PER_LAYER + k. */
enum {
  CLIPDLM_P_POS = 0, CLIPDLM_P_EMB_LN_W, CLIPDLM_P_EMB_LN_B, CLIPDLM_P_VT_W, CLIPDLM_P_VT_B, CLIPDLM_P_VLN_W, CLIPDLM_P_VLN_B,
  CLIPDLM_P_IMG_W, CLIPDLM_P_IMG_B, CLIPDLM_P_TXT_W, CLIPDLM_P_TXT_B, CLIPDLM_P_SEG, CLIPDLM_P_LAYER0,
  /* per layer: */
  CLIPDLM_PL_QKV_W = 0, CLIPDLM_PL_QKV_B, CLIPDLM_PL_O_W, CLIPDLM_PL_O_B, CLIPDLM_PL_LN1_W, CLIPDLM_PL_LN1_B,
  CLIPDLM_PL_FF1_W, CLIPDLM_PL_FF1_B, CLIPDLM_PL_FF2_W, CLIPDLM_PL_FF2_B, CLIPDLM_PL_LN2_W, CLIPDLM_PL_LN2_B, CLIPDLM_P_PER_LAYER
};
int64_t clipdlm_param_count(const clipdlm_config_t* cfg);                 /* total fp32 elements of the flat buffer */
int64_t clipdlm_param_offset(const clipdlm_config_t* cfg, int32_t slot);  /* element offset of a slot, <0 if invalid */
int64_t clipdlm_param_size(const clipdlm_config_t* cfg, int32_t slot);

size_t clipdlm_workspace_bytes(const clipdlm_config_t* cfg, int32_t max_rows, int32_t batch, int32_t training);

typedef struct clipdlm_buffers {
  float* params;            /* flat fp32 master weights (trainable; reference model.parameters(), CLIP-DDPM.py:258-269) */
  float* grads;             /* flat fp32 gradients (accumulated) */
  void* shadow_hi; void* shadow_lo;   /* flat bf16 (pair) copy of params */
  const float* emb_table;   /* frozen E [vocab, dim] fp32 (model.embedding, CLIP-DDPM.py:245) */
  void* emb_hi; void* emb_lo;         /* bf16 (pair) copy of the frozen lm_head weight [vocab, dim] (CLIP-DDPM.py:246) */
  void* workspace; size_t workspace_bytes;
} clipdlm_buffers_t;

clipdlm_engine_t* clipdlm_engine_create(const clipdlm_config_t* cfg, const clipdlm_buffers_t* bufs, int32_t max_rows, int32_t batch,
                                        int32_t training);
void clipdlm_engine_destroy(clipdlm_engine_t* e);

/* One encoder pass over R rows (R <= max_rows), = DistilBertModel.forward without the lm_head (CLIP-DDPM.py:271-322).
 * Inputs as clipdlm_embed_t mode 0/1 (x_in XOR ids+noise+coef); image_clip/text_clip are [B, clip_dim] fp32 with row r using
 * caption r % B; attn_mask [B, max_len] int32 (0/1); guided = concat_mask[:,1] (uniform over rows).
 * train = 1 enables dropout and keeps activations for clipdlm_engine_backward. x_out (fp32 [R, L, D]) may be NULL. */
typedef struct clipdlm_pass {
  int32_t R, B, mode, guided, train;
  int32_t reuse_proj; /* 1: image_clip / text_clip / attn_mask / guided are those of the previous pass (denoise loop): skip the CLIP projections + key mask */
  const float* x_in; int64_t x_in_stride; /* mode 0 input and its row pitch in elements (0 = dense max_len * dim) */
  const int32_t* ids; const float* noise; const float* coef_a; const float* coef_b;
  const float* image_clip; const float* text_clip; const int32_t* attn_mask;
  uint64_t drop_seed;
  float* x_out;
} clipdlm_pass_t;
int clipdlm_engine_forward(clipdlm_engine_t* e, const clipdlm_pass_t* p, clipdlm_stream stream);

/* lm_head over x_out[:, :max_len] of the last forward: logits fp32 [R*max_len, ld_logits] with ld_logits >= vocab rounded
 * up to 32 (may be NULL; the padding columns receive zeros) and/or argmax ids int32 [R*max_len] (may be NULL; computed by
 * the fused GEMM + running-argmax epilogue, logits never reach HBM). The bf16 copy of the lm_head weight (emb_hi/lo) must be
 * zero-padded to a multiple of 256 rows. Replaces self.lm_head(x_out[:, :MAX_LENGTH]) and argmax (CLIP-DDPM.py:323,620). */
int clipdlm_engine_lm_head(clipdlm_engine_t* e, float* logits, int64_t ld_logits, int32_t* argmax, clipdlm_stream stream);

/* Loss of the last forward + full backward into bufs.grads (+=). Adds to losses[0] (embedding loss) and losses[1]
 * (cross-entropy, unweighted) as doubles. Row means use R_total (rows of the whole pass when chunked).
 * Replaces loss() terms (CLIP-DDPM.py:415-437) and l.backward() (:483) for the rows of this pass. */
typedef struct clipdlm_loss_cfg {
  int32_t loss_kind;      /* LOSS_FUNC id, see clipdlm_embed_loss */
  int32_t use_embed_loss; /* USE_X_T_LOSS / USE_X_1_LOSS */
  int32_t use_prob_loss;  /* USE_PROB_LOSS */
  int32_t batch_size;     /* BATCH_SIZE */
  int64_t R_total;
  float rounding_weight;  /* ROUNDING_WEIGHT */
  int32_t backward;       /* 0 = losses only (validate, CLIP-DDPM.py:488-501) */
  const float* target;    /* optional explicit embedding-loss target, see clipdlm_embed_loss */
  int32_t target_rows;
  /* Classifier-free-guidance training (CLIP-DDPM.py:313-317,406-410): x_out of this engine is the MIX of an unguided pass (this
   * engine) and a guided pass (another engine, see clipdlm_engine_cfg_mix). After the loss gradient d(x_out) is formed, row r of
   * it is multiplied by row_scale_self[r] before this engine's backward continues, and - if export_engine is set - d(x_out) *
   * row_scale_export[r] is written into that engine's upstream-gradient buffer for clipdlm_engine_backward. NULL = plain path. */
  const float* row_scale_self;    /* [R] device, or NULL */
  const float* row_scale_export;  /* [R] device, or NULL */
  clipdlm_engine_t* export_engine;
  /* Optional DEVICE scalar multiplying rounding_weight in the cross-entropy gradient (NULL = 1): the reference's dynamic rounding weight
   * (CLIP-DDPM.py:535-536) is a tensor recomputed every step from the running loss sums; with it on the device no step ever reads a
   * scalar back to the host. (The loss VALUES are returned unweighted either way.) */
  const float* rounding_weight_dev;
} clipdlm_loss_cfg_t;
int clipdlm_engine_loss_backward(clipdlm_engine_t* e, const clipdlm_loss_cfg_t* lc, double* losses, clipdlm_stream stream);
/* x_out(e_unguided)[r] <- guided[r] ? (1 + w) * x_out(e_guided)[r] - w * x_out(e_unguided)[r] : unchanged, over the R rows of the two
 * engines' last forward passes (same R, B). Replaces the guidance mix of DistilBertModel.forward (CLIP-DDPM.py:313-317); the
 * mixed x_out is what clipdlm_engine_loss_backward / clipdlm_engine_lm_head of e_unguided then see. */
int clipdlm_engine_cfg_mix(clipdlm_engine_t* e_unguided, clipdlm_engine_t* e_guided, const int32_t* guided /* [R] device */, float w,
                           clipdlm_stream stream);
/* Backward of the last forward of `e` from the upstream gradient d(x_out) another engine exported into it (row_scale_export
 * above): transform head, blocks, embeddings; gradients accumulate into bufs.grads. */
int clipdlm_engine_backward(clipdlm_engine_t* e, clipdlm_stream stream);

/* Backward of the last forward of `e` from an explicit upstream gradient dx_out (fp32 [R, L, dim], the gradient of the fp32 x_out the
 * forward returned) and, optionally, the gradient of the mode-0 input x_in (fp32 [R, max_len, dim]) written to dx_in. Used when the
 * encoder sits between caller-side layers (TRAIN_EMBEDDING=True: input_projection / output_projection, CLIP-DDPM.py:292-293,319-320). */
int clipdlm_engine_backward_from(clipdlm_engine_t* e, const float* dx_out, float* dx_in /* may be NULL */, clipdlm_stream stream);

/* Engine options (default 0). FUSED_SOFTMAX_GRAD = 1: plain-bf16 training passes take the factored softmax gradient
 * (LSE_EXP lm_head pass -> clipdlm_ce_row_terms -> STORE_ROWSCALE gradient GEMM) instead of the in-place softmax-gradient pass over the
 * stored logits; EXP_SHIFT_PTR: device pointer (as int64) to the fp32 scalar c of that path, 0 = use c = 0. The path needs
 * max_v s_v - 69 <= c <= max_v s_v + 87 for every row (exp(s - c) must neither saturate the 2^100 clamp nor flush to zero);
 * c = 0 holds whenever the largest logit of every row lies in [-87, 69]. Experimental in round 1: validated on the GPU in round 2. */
/* GELU_DERIV_STORE = 1 (plain bf16): lin1 of every block stores gelu'(u) instead of u in passes of an engine with training buffers
 * (STORE_GELU_DERIV) and the lin2 gradient GEMM multiplies by it (STORE_MULAUX). Set it before the forward whose backward should use it.
 * GELU_DERIV_STORE = 2: additionally that GEMM's epilogue accumulates lin1's bias gradient (column sums of its output), which
 * replaces one clipdlm_colsum pass over the [tokens, hidden_dim] gradient per block.
 * Experimental in round 1 as well. */
enum { CLIPDLM_OPT_FUSED_SOFTMAX_GRAD = 1, CLIPDLM_OPT_EXP_SHIFT_PTR = 2, CLIPDLM_OPT_GELU_DERIV_STORE = 3 };
int clipdlm_engine_set_option(clipdlm_engine_t* e, int32_t option, int64_t value);

/* Number of kernel launches issued by this engine since creation (bench "gpu_launches"). */
int64_t clipdlm_engine_launch_count(const clipdlm_engine_t* e);

/* Per-launch device timing for the roofline report: when enabled every engine launch is bracketed by two CUDA events on
 * the caller's stream; profile_read() waits for them and returns per-category totals (ms on the device, ALGORITHMIC flops
 * = 2MNK for GEMMs, algorithmic bytes for the HBM-bound kernels, launch count); reset = 1 recycles the events. */
enum {
  CLIPDLM_PROF_GEMM_FWD = 0, CLIPDLM_PROF_GEMM_DGRAD, CLIPDLM_PROF_GEMM_WGRAD, CLIPDLM_PROF_GEMM_LSE, CLIPDLM_PROF_GEMM_SMGRAD,
  CLIPDLM_PROF_ATTN_FWD, CLIPDLM_PROF_ATTN_BWD, CLIPDLM_PROF_LN_FWD, CLIPDLM_PROF_LN_BWD, CLIPDLM_PROF_EMBED, CLIPDLM_PROF_LOSS,
  CLIPDLM_PROF_COLSUM, CLIPDLM_PROF_OTHER, CLIPDLM_PROF_NCAT
};
typedef struct clipdlm_prof { double ms; double flops; double bytes; int64_t launches; } clipdlm_prof_t;
int clipdlm_engine_profile(clipdlm_engine_t* e, int32_t enable);
int clipdlm_engine_profile_read(clipdlm_engine_t* e, clipdlm_prof_t* out /* [CLIPDLM_PROF_NCAT] */, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* CLIPDLM_H */
