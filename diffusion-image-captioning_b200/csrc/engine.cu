// Native orchestration of the composite hot path: DistilBertModel.forward (CLIP-DDPM.py:271-323 + HF DistilBertForMaskedLM),
// the loss terms of loss() (:415-437) with the full hand-written backward (replaces autograd's l.backward(), :483), and the
// lm_head rounding / argmax used by the denoise loop (:616-621).  All kernels are launched on the caller's stream; the engine
// owns no memory (it carves the caller's workspace) and keeps, per pass, only the activations the backward needs.
#include "common.cuh"
#include "../../include/clipdlm.h"
#include <string.h>
#include <new>
#include <vector>

namespace clipdlm {

int num_sms();
int gemm_dispatch(const clipdlm_gemm_t* g, cudaStream_t st);
int lse_combine_dispatch(const float* pmax, const float* psum, const int* parg, int n_tiles, int M, const float* tgt_logit, float* lse,
                         int* argmax, double* loss_acc, double scale, cudaStream_t st);
int cfg_mix_dispatch(const clipdlm_bf_t* xu, const clipdlm_bf_t* xg, const int* guided, float w, int R, int L, int D, cudaStream_t st);
int row_scale_dispatch(const clipdlm_bf_t* g, const float* s_self, const clipdlm_bf_t* ex, const float* s_ex, int R, int L, int D,
                       cudaStream_t st);
int softmax_grad_inplace_dispatch(void* logits, long long ld, int M, int N, const float* lse, const int* targets, int tgt_period, float scale,
                                  cudaStream_t st, const float* scale_mul);
int embed_fwd_dispatch(const clipdlm_embed_t* e, cudaStream_t st);
int embed_bwd_dispatch(const clipdlm_bf_t* dz, int R, int B, int Ltxt, int L, int D, int fusion, int guided, float* d_pos, float* d_seg,
                       float* d_img, float* d_txt, cudaStream_t st);
int layernorm_fwd_dispatch(const clipdlm_bf_t* z, const float* w, const float* b, float eps, long long rows, int D, const clipdlm_bf_t* y,
                           float* y_f32, unsigned long long seed, uint32_t site, float p, cudaStream_t st);
int layernorm_bwd_dispatch(const clipdlm_bf_t* z, const clipdlm_bf_t* dy, const float* w, float eps, long long rows, int D,
                           const clipdlm_bf_t* dz, float* dw, float* db, unsigned long long seed, uint32_t site_out, float p_out,
                           const clipdlm_bf_t* dz_drop, uint32_t site_in, float p_in, const clipdlm_bf_t* gelu_u, float* dbias,
                           cudaStream_t st);
int colsum_dispatch(const clipdlm_bf_t* x, long long rows, int N, float* out, cudaStream_t st);
int ce_row_terms_dispatch(const float* lse, const float* exp_shift, const int* targets, int tgt_period, float scale, int M, const void* w,
                          long long ldw, void* dx, long long ldx, int scatter_len, int scatter_stride, int D, float* row_scale, cudaStream_t st,
                          const float* scale_mul);
int embed_loss_dispatch(const clipdlm_bf_t* x_out, const float* emb, const int* ids, const float* tgt, int tgt_rows, int R, int B, int Ltxt,
                        int L, int D, int kind, long long R_total, int batch_size, float weight, double* loss_acc, const clipdlm_bf_t* dx,
                        cudaStream_t st);
int small_linear_fwd_dispatch(const float* x, const float* w, const float* b, int B, int K, int N, float* y, cudaStream_t st);
int small_linear_bwd_dispatch(const float* x, const float* dy, int B, int K, int N, float* dw, float* db, cudaStream_t st);
int keymask_dispatch(const int* attn_mask, int R, int B, int Ltxt, int L, int fusion, int guided, uint32_t* km, cudaStream_t st);
int to_bf16_dispatch(const float* x, void* hi, void* lo, long long n, cudaStream_t st);
int gather_rows_f32_dispatch(const clipdlm_bf_t* x, long long rows_out, int len, int stride, int D, float* y, cudaStream_t st);
int attn_fwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, int R, int L, int D, int H, const clipdlm_bf_t* ctx,
                      unsigned long long seed, uint32_t site, float p, cudaStream_t st);
int attn_bwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int R, int L, int D, int H,
                      const clipdlm_bf_t* dqkv, unsigned long long seed, uint32_t site, float p, cudaStream_t st, float* dbias, int* folded);

// ------------------------------------------------------------------------------------------------------------------
// flat parameter layout
// ------------------------------------------------------------------------------------------------------------------
static long long slot_size(const clipdlm_config_t* c, int slot) {
  const long long D = c->dim, F = c->hidden_dim, C = c->clip_dim;
  if (slot < 0) return -1;
  if (slot < CLIPDLM_P_LAYER0) {
    switch (slot) {
      case CLIPDLM_P_POS: return (long long)c->max_pos * D;
      case CLIPDLM_P_EMB_LN_W: case CLIPDLM_P_EMB_LN_B: case CLIPDLM_P_VT_B: case CLIPDLM_P_VLN_W: case CLIPDLM_P_VLN_B:
      case CLIPDLM_P_IMG_B: case CLIPDLM_P_TXT_B: return D;
      case CLIPDLM_P_VT_W: return D * D;
      case CLIPDLM_P_IMG_W: case CLIPDLM_P_TXT_W: return D * C;
      case CLIPDLM_P_SEG: return c->fusion == 0 ? 2 * D : 0;
    }
    return -1;
  }
  const int rel = slot - CLIPDLM_P_LAYER0;
  if (rel / CLIPDLM_P_PER_LAYER >= c->n_layers) return -1;
  switch (rel % CLIPDLM_P_PER_LAYER) {
    case CLIPDLM_PL_QKV_W: return 3 * D * D;
    case CLIPDLM_PL_QKV_B: return 3 * D;
    case CLIPDLM_PL_O_W: return D * D;
    case CLIPDLM_PL_FF1_W: case CLIPDLM_PL_FF2_W: return D * F;
    case CLIPDLM_PL_FF1_B: return F;
    default: return D;  // O_B, LN1_W/B, FF2_B, LN2_W/B
  }
}
static int num_slots(const clipdlm_config_t* c) { return CLIPDLM_P_LAYER0 + c->n_layers * CLIPDLM_P_PER_LAYER; }
static long long slot_offset(const clipdlm_config_t* c, int slot) {
  if (slot < 0 || slot > num_slots(c)) return -1;
  long long off = 0;
  for (int s = 0; s < slot; ++s) off += slot_size(c, s);
  return off;
}

// ------------------------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------------------------
struct Carver {
  uint8_t* base; size_t off;
  void* take(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};
typedef clipdlm_bf_t Act;

struct LayerBufs { Act qkv, ctx, z1, h1, u, g, z2; };

// Optional per-launch timing: every launch is bracketed by two CUDA events on the caller's stream (bench.py's roofline leg).
struct ProfRec { cudaEvent_t a, b; int cat; double flops, bytes; };
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};

}  // namespace clipdlm

using namespace clipdlm;

struct clipdlm_engine {
  clipdlm_config_t cfg;
  clipdlm_buffers_t bufs;
  int max_rows, batch, training, pair;
  int L, Ltxt, n_vtiles;
  long long ldl;  // pitch of the d(logits) buffer
  // activations
  Act z0;
  Act* h;           // [n_layers + 1]
  LayerBufs* lay;   // [n_layers]
  Act uv, gv, xo;
  // backward temporaries
  Act g0, g1, g2, gf, gq, dlog;
  float *img_proj, *txt_proj, *d_img_proj, *d_txt_proj;
  uint32_t* keymask;
  float *part_max, *part_sum, *tgt_logit, *lse;
  int32_t* part_arg;
  long long launches;
  Profiler* prof;
  // last forward
  clipdlm_pass_t last;
  int have_fwd;
  // options (clipdlm_engine_set_option)
  int fused_smgrad;          // CLIPDLM_OPT_FUSED_SOFTMAX_GRAD
  const float* exp_shift;    // CLIPDLM_OPT_EXP_SHIFT_PTR (device scalar) or NULL
  int gelu_deriv;            // CLIPDLM_OPT_GELU_DERIV_STORE
  int last_gelu_deriv;       // the last forward stored gelu'(u) in the blocks' u buffers
};

namespace clipdlm {

static Act take_act(Carver& c, size_t elems, int pair) {
  Act a;
  a.hi = c.take(elems * 2);
  a.lo = pair ? c.take(elems * 2) : nullptr;
  return a;
}

static size_t carve(clipdlm_engine* e, uint8_t* base) {
  const clipdlm_config_t& c = e->cfg;
  Carver cv{base, 0};
  const size_t T = (size_t)e->max_rows * e->L, T16 = (size_t)e->max_rows * e->Ltxt;
  const size_t D = c.dim, F = c.hidden_dim;
  const int pair = e->pair, NL = c.n_layers;
  const bool tr = e->training != 0;
  e->z0 = tr ? take_act(cv, T * D, pair) : Act{nullptr, nullptr};
  for (int i = 0; i <= NL; ++i) e->h[i] = (tr || i < 2) ? take_act(cv, T * D, pair) : e->h[i & 1];
  for (int l = 0; l < NL; ++l) {
    if (tr || l == 0) {
      LayerBufs& b = e->lay[l];
      b.qkv = take_act(cv, T * 3 * D, pair);
      b.ctx = take_act(cv, T * D, pair);
      b.z1 = take_act(cv, T * D, pair);
      b.h1 = take_act(cv, T * D, pair);
      b.u = tr ? take_act(cv, T * F, pair) : Act{nullptr, nullptr};
      b.g = take_act(cv, T * F, pair);
      b.z2 = take_act(cv, T * D, pair);
    } else {
      e->lay[l] = e->lay[0];
    }
  }
  e->uv = tr ? take_act(cv, T * D, pair) : Act{nullptr, nullptr};
  e->gv = take_act(cv, T * D, pair);
  e->xo = take_act(cv, T * D, pair);
  if (tr) {
    e->g0 = take_act(cv, T * D, pair);
    e->g1 = take_act(cv, T * D, pair);
    e->g2 = take_act(cv, T * D, pair);
    e->gf = take_act(cv, T * F, pair);
    e->gq = take_act(cv, T * 3 * D, pair);
    e->dlog = take_act(cv, T16 * (size_t)e->ldl, pair);
  } else {
    e->g0 = e->g1 = e->g2 = e->gf = e->gq = e->dlog = Act{nullptr, nullptr};
  }
  e->img_proj = (float*)cv.take((size_t)e->batch * D * 4);
  e->txt_proj = (float*)cv.take((size_t)e->batch * D * 4);
  e->d_img_proj = (float*)cv.take((size_t)e->batch * D * 4);
  e->d_txt_proj = (float*)cv.take((size_t)e->batch * D * 4);
  e->keymask = (uint32_t*)cv.take((size_t)e->max_rows * ((e->L + 31) / 32) * 4);
  const size_t lse_slots = 2 * (size_t)e->n_vtiles;  // the LSE epilogue emits one partial per 128-column half tile
  e->part_max = (float*)cv.take(lse_slots * T16 * 4);
  e->part_sum = (float*)cv.take(lse_slots * T16 * 4);
  e->part_arg = (int32_t*)cv.take(lse_slots * T16 * 4);
  e->tgt_logit = (float*)cv.take(T16 * 4);
  e->lse = (float*)cv.take(T16 * 4);
  return (cv.off + 255) & ~size_t(255);
}

static int init_shape(clipdlm_engine* e, const clipdlm_config_t* cfg, int max_rows, int batch, int training) {
  CLIPDLM_CHECK(cfg != nullptr, "null config");
  CLIPDLM_CHECK(cfg->n_layers >= 1 && cfg->n_layers <= 64, "n_layers %d out of range", cfg->n_layers);
  CLIPDLM_CHECK(cfg->dim % 256 == 0 && cfg->dim <= 1024, "dim %d must be a multiple of 256, <= 1024", cfg->dim);
  CLIPDLM_CHECK(cfg->n_heads * 64 == cfg->dim, "head dim must be 64 (dim %d, heads %d)", cfg->dim, cfg->n_heads);
  CLIPDLM_CHECK(cfg->hidden_dim % 256 == 0, "hidden_dim %d must be a multiple of 256", cfg->hidden_dim);
  CLIPDLM_CHECK(cfg->fusion == 0 || cfg->fusion == 1, "fusion must be 0 (concat) or 1 (add)");
  CLIPDLM_CHECK(cfg->max_len >= 1 && 128 % cfg->max_len == 0, "max_len %d must divide 128 (TMA row gather of x_out[:, :max_len])", cfg->max_len);
  CLIPDLM_CHECK(cfg->clip_dim % 8 == 0 && cfg->vocab > 0, "bad clip_dim / vocab");
  CLIPDLM_CHECK(max_rows > 0 && batch > 0, "max_rows / batch must be positive");
  e->cfg = *cfg;
  e->max_rows = max_rows; e->batch = batch; e->training = training;
  e->pair = cfg->precision == 1;
  e->Ltxt = cfg->max_len;
  e->L = cfg->fusion == 0 ? cfg->max_len + 2 : cfg->max_len;
  CLIPDLM_CHECK(e->L <= cfg->max_pos && e->L <= 128, "sequence length %d exceeds max_pos %d / 128", e->L, cfg->max_pos);
  e->n_vtiles = (cfg->vocab + 255) / 256;
  e->ldl = (long long)e->n_vtiles * 256;
  return 0;
}

// ---- small helpers -------------------------------------------------------------------------------------------------
static inline Act shadow(const clipdlm_engine* e, int slot) {
  const long long off = slot_offset(&e->cfg, slot);
  Act a;
  a.hi = (uint8_t*)e->bufs.shadow_hi + off * 2;
  a.lo = (e->pair && e->bufs.shadow_lo) ? (uint8_t*)e->bufs.shadow_lo + off * 2 : nullptr;
  return a;
}
static inline float* param(const clipdlm_engine* e, int slot) { return e->bufs.params + slot_offset(&e->cfg, slot); }
static inline float* grad(const clipdlm_engine* e, int slot) { return e->bufs.grads + slot_offset(&e->cfg, slot); }
static inline int lslot(int layer, int k) { return CLIPDLM_P_LAYER0 + layer * CLIPDLM_P_PER_LAYER + k; }

static clipdlm_gemm_t gemm_desc(const Act& a, long long lda, int a_major, const Act& b, long long ldb, int b_major, int M, int N, int K) {
  clipdlm_gemm_t g;
  memset(&g, 0, sizeof(g));
  g.a_hi = a.hi; g.a_lo = a.lo; g.b_hi = b.hi; g.b_lo = b.lo;
  g.lda = lda; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.a_major = a_major; g.b_major = b_major;
  g.epilogue = CLIPDLM_EPI_STORE;
  return g;
}
struct ProfScope {
  clipdlm_engine* e; cudaStream_t st; int idx;
  ProfScope(clipdlm_engine* e_, cudaStream_t st_, int cat, double flops, double bytes) : e(e_), st(st_), idx(-1) {
    if (e->prof && e->prof->on) {
      ProfRec r; r.a = e->prof->get(); r.b = e->prof->get(); r.cat = cat; r.flops = flops; r.bytes = bytes;
      cudaEventRecord(r.a, st);
      e->prof->recs.push_back(r);
      idx = (int)e->prof->recs.size() - 1;
    }
  }
  void end() { if (idx >= 0) cudaEventRecord(e->prof->recs[idx].b, st); }
};
#define RUNP(CAT, FLOPS, BYTES, expr)                                     \
  do {                                                                    \
    ProfScope _ps(e, st, (CAT), (double)(FLOPS), (double)(BYTES));        \
    int _rc = (expr);                                                     \
    _ps.end();                                                            \
    if (_rc) return _rc;                                                  \
    e->launches++;                                                        \
  } while (0)
#define RUN(expr) RUNP(CLIPDLM_PROF_OTHER, 0, 0, expr)
// a GEMM launch: category from the operand majors / epilogue, algorithmic flops 2MNK, bytes = operands + outputs once
static int gemm_cat(const clipdlm_gemm_t& g) {
  switch (g.epilogue) {
    case CLIPDLM_EPI_WGRAD: return CLIPDLM_PROF_GEMM_WGRAD;
    case CLIPDLM_EPI_LSE: case CLIPDLM_EPI_LSE_EXP: return CLIPDLM_PROF_GEMM_LSE;
    case CLIPDLM_EPI_SMGRAD: return CLIPDLM_PROF_GEMM_SMGRAD;
    default: return g.b_major ? CLIPDLM_PROF_GEMM_DGRAD : CLIPDLM_PROF_GEMM_FWD;
  }
}
static double gemm_bytes(const clipdlm_gemm_t& g, int es) {
  double b = ((double)g.M * g.K + (double)g.N * g.K) * es;
  if (g.epilogue == CLIPDLM_EPI_WGRAD) b += (double)g.M * g.N * 8;  // fp32 accumulate: read + write
  else if (g.epilogue == CLIPDLM_EPI_SMGRAD) b += (double)g.M * g.N * es;
  else if (g.epilogue == CLIPDLM_EPI_STORE || g.epilogue >= CLIPDLM_EPI_STORE_ROWSCALE) {
    if (g.out_hi) b += (double)g.M * g.N * es;
    if (g.out2_hi) b += (double)g.M * g.N * es;
    if (g.out_f32) b += (double)g.M * g.N * 4;
    if (g.res_hi) b += (double)g.M * g.N * es;
    if (g.u_hi) b += (double)g.M * g.N * es;
  }
  return b;
}
#define RUNG(gd) RUNP(gemm_cat(gd), 2.0 * (gd).M * (gd).N * (gd).K, gemm_bytes((gd), e->pair ? 4 : 2), gemm_dispatch(&(gd), st))

// y[T, N] = x[T, K] W[N, K]^T + bias  (+ fused extras set by the caller on the descriptor)
static clipdlm_gemm_t linear_fwd(const Act& x, const Act& w, const float* bias, int T, int N, int K, const Act& out) {
  clipdlm_gemm_t g = gemm_desc(x, K, 0, w, K, 0, T, N, K);
  g.bias = bias; g.out_hi = out.hi; g.out_lo = out.lo; g.ldo = N;
  return g;
}
// dx[T, K] = dy[T, N] W[N, K]   (W stored [N][K]: reduction dim N is the slow one -> MN-major B)
static clipdlm_gemm_t linear_dgrad(const Act& dy, const Act& w, int T, int N, int K, const Act& out) {
  clipdlm_gemm_t g = gemm_desc(dy, N, 0, w, K, 1, T, K, N);
  g.out_hi = out.hi; g.out_lo = out.lo; g.ldo = K;
  return g;
}
// dW[N, K] += dy[T, N]^T x[T, K]
static clipdlm_gemm_t linear_wgrad(const Act& dy, const Act& x, int T, int N, int K, float* dw) {
  clipdlm_gemm_t g = gemm_desc(dy, N, 1, x, K, 1, N, K, T);
  g.epilogue = CLIPDLM_EPI_WGRAD; g.acc_f32 = dw; g.ldo = K;
  return g;
}

static int forward_impl(clipdlm_engine* e, const clipdlm_pass_t* p, cudaStream_t st) {
  const clipdlm_config_t& c = e->cfg;
  CLIPDLM_CHECK(p != nullptr, "null pass");
  CLIPDLM_CHECK(p->R > 0 && p->R <= e->max_rows, "pass rows %d out of range (max_rows %d)", p->R, e->max_rows);
  CLIPDLM_CHECK(p->B > 0 && p->B <= e->batch && p->R % p->B == 0, "pass batch %d invalid (engine batch %d, rows %d)", p->B, e->batch, p->R);
  CLIPDLM_CHECK(p->image_clip && p->text_clip, "pass without CLIP features");
  CLIPDLM_CHECK(!p->train || e->training, "engine was created for inference; train pass refused");
  const int R = p->R, B = p->B, L = e->L, Ltxt = e->Ltxt, D = c.dim, F = c.hidden_dim, NL = c.n_layers;
  const int T = R * L;
  const bool train = p->train != 0;
  const float pdrop = train ? c.dropout : 0.f, padrop = train ? c.attn_dropout : 0.f;

  if (p->reuse_proj) {
    CLIPDLM_CHECK(e->have_fwd && e->last.R == R && e->last.B == B && e->last.guided == p->guided, "reuse_proj: previous pass had a different shape");
  } else {
    RUN(small_linear_fwd_dispatch(p->image_clip, param(e, CLIPDLM_P_IMG_W), param(e, CLIPDLM_P_IMG_B), B, c.clip_dim, D, e->img_proj, st));
    RUN(small_linear_fwd_dispatch(p->text_clip, param(e, CLIPDLM_P_TXT_W), param(e, CLIPDLM_P_TXT_B), B, c.clip_dim, D, e->txt_proj, st));
    RUN(keymask_dispatch(p->attn_mask, R, B, Ltxt, L, c.fusion, p->guided, e->keymask, st));
  }

  clipdlm_embed_t em;
  memset(&em, 0, sizeof(em));
  em.R = R; em.B = B; em.Ltxt = Ltxt; em.L = L; em.D = D; em.fusion = c.fusion; em.mode = p->mode; em.guided = p->guided;
  em.x_in = p->x_in; em.x_in_stride = p->x_in_stride;
  em.emb_table = e->bufs.emb_table; em.ids = p->ids; em.noise = p->noise; em.coef_a = p->coef_a; em.coef_b = p->coef_b;
  em.img_proj = e->img_proj; em.txt_proj = e->txt_proj;
  em.seg = c.fusion == 0 ? param(e, CLIPDLM_P_SEG) : nullptr; em.pos = param(e, CLIPDLM_P_POS);
  em.ln_w = param(e, CLIPDLM_P_EMB_LN_W); em.ln_b = param(e, CLIPDLM_P_EMB_LN_B); em.ln_eps = c.ln_eps;
  em.z = e->z0; em.h = e->h[0];
  em.drop_seed = p->drop_seed; em.drop_site = 0; em.drop_p = pdrop;
  RUNP(CLIPDLM_PROF_EMBED, 0, (double)T * D * (e->pair ? 4.0 : 2.0) * (train ? 2 : 1) + (double)B * Ltxt * D * 8, embed_fwd_dispatch(&em, st));

  const bool gelu_deriv = e->gelu_deriv != 0 && !e->pair && e->training && e->lay[0].u.hi != nullptr;
  for (int l = 0; l < NL; ++l) {
    const LayerBufs& b = e->lay[l];
    const Act& hin = e->h[l];
    clipdlm_gemm_t g = linear_fwd(hin, shadow(e, lslot(l, CLIPDLM_PL_QKV_W)), param(e, lslot(l, CLIPDLM_PL_QKV_B)), T, 3 * D, D, b.qkv);
    RUNG(g);
    RUNP(CLIPDLM_PROF_ATTN_FWD, 4.0 * R * L * L * D, (double)T * 4 * D * (e->pair ? 4.0 : 2.0), attn_fwd_dispatch(&b.qkv, e->keymask, R, L, D, c.n_heads, &b.ctx, p->drop_seed, 1 + 2 * l, padrop, st));
    g = linear_fwd(b.ctx, shadow(e, lslot(l, CLIPDLM_PL_O_W)), param(e, lslot(l, CLIPDLM_PL_O_B)), T, D, D, b.z1);
    g.res_hi = hin.hi; g.res_lo = hin.lo; g.ldr = D;
    RUNG(g);
    RUNP(CLIPDLM_PROF_LN_FWD, 0, (double)T * 2 * D * (e->pair ? 4.0 : 2.0), layernorm_fwd_dispatch(&b.z1, param(e, lslot(l, CLIPDLM_PL_LN1_W)), param(e, lslot(l, CLIPDLM_PL_LN1_B)), c.ln_eps, T, D, &b.h1,
                               nullptr, 0, 0, 0.f, st));
    g = linear_fwd(b.h1, shadow(e, lslot(l, CLIPDLM_PL_FF1_W)), param(e, lslot(l, CLIPDLM_PL_FF1_B)), T, F, D, b.u);
    g.out2_hi = b.g.hi; g.out2_lo = b.g.lo;
    if (gelu_deriv) g.epilogue = CLIPDLM_EPI_STORE_GELU_DERIV;   // b.u receives gelu'(u): the backward multiplies instead of evaluating it
    RUNG(g);
    g = linear_fwd(b.g, shadow(e, lslot(l, CLIPDLM_PL_FF2_W)), param(e, lslot(l, CLIPDLM_PL_FF2_B)), T, D, F, b.z2);
    g.res_hi = b.h1.hi; g.res_lo = b.h1.lo; g.ldr = D;
    g.drop_seed = p->drop_seed; g.drop_site = 2 + 2 * l; g.drop_p = pdrop;
    RUNG(g);
    RUNP(CLIPDLM_PROF_LN_FWD, 0, (double)T * 2 * D * (e->pair ? 4.0 : 2.0), layernorm_fwd_dispatch(&b.z2, param(e, lslot(l, CLIPDLM_PL_LN2_W)), param(e, lslot(l, CLIPDLM_PL_LN2_B)), c.ln_eps, T, D, &e->h[l + 1],
                               nullptr, 0, 0, 0.f, st));
  }
  clipdlm_gemm_t g = linear_fwd(e->h[NL], shadow(e, CLIPDLM_P_VT_W), param(e, CLIPDLM_P_VT_B), T, D, D, e->uv);
  g.out2_hi = e->gv.hi; g.out2_lo = e->gv.lo;
  RUNG(g);
  RUNP(CLIPDLM_PROF_LN_FWD, 0, (double)T * 2 * D * (e->pair ? 4.0 : 2.0), layernorm_fwd_dispatch(&e->gv, param(e, CLIPDLM_P_VLN_W), param(e, CLIPDLM_P_VLN_B), c.ln_eps, T, D, &e->xo, p->x_out, 0, 0, 0.f, st));
  e->last = *p;
  e->last_gelu_deriv = gelu_deriv ? 1 : 0;
  e->have_fwd = 1;
  return 0;
}

// LSE-fused lm_head over x_out[:, :Ltxt]: partials -> combine.  targets may be NULL (argmax only).
// exp_mode (factored softmax gradient): keep_logits receives bf16(exp(logit - *exp_shift)) instead of the logits, see CLIPDLM_EPI_LSE_EXP.
static int lm_head_lse(clipdlm_engine* e, const int32_t* targets, int tgt_period, int32_t* argmax, double* loss_acc, double scale,
                       cudaStream_t st, void* keep_logits = nullptr, bool exp_mode = false) {
  const clipdlm_config_t& c = e->cfg;
  const int M = e->last.R * e->Ltxt;
  Act emb{e->bufs.emb_hi, e->pair ? e->bufs.emb_lo : nullptr};
  clipdlm_gemm_t g = gemm_desc(e->xo, c.dim, 0, emb, c.dim, 0, M, c.vocab, c.dim);
  g.gather_len = e->Ltxt; g.gather_stride = e->L;
  g.epilogue = CLIPDLM_EPI_LSE;
  g.part_max = e->part_max; g.part_sum = e->part_sum; g.part_arg = argmax ? e->part_arg : nullptr; g.tgt_logit = e->tgt_logit;
  g.targets = targets; g.tgt_period = tgt_period;
  if (keep_logits != nullptr) { g.out_hi = keep_logits; g.ldo = e->ldl; }   // bf16 logits for the in-place softmax gradient
  if (exp_mode) { g.epilogue = CLIPDLM_EPI_LSE_EXP; g.exp_shift = e->exp_shift; }
  RUNG(g);
  RUNP(CLIPDLM_PROF_LOSS, 0, 6.0 * e->n_vtiles * M * 4, lse_combine_dispatch(e->part_max, e->part_sum, argmax ? e->part_arg : nullptr, 2 * e->n_vtiles, M, targets ? e->tgt_logit : nullptr, e->lse, argmax, loss_acc,
                           scale, st));
  return 0;
}

// Backward of the last forward from the upstream gradient d(x_out) held in e->g0: transform head, blocks, embeddings.
static int backward_from_g0(clipdlm_engine* e, cudaStream_t st) {
  const clipdlm_config_t& c = e->cfg;
  const clipdlm_pass_t& p = e->last;
  const int R = p.R, B = p.B, L = e->L, Ltxt = e->Ltxt, D = c.dim, F = c.hidden_dim, NL = c.n_layers;
  const int T = R * L;
  const bool train = p.train != 0;
  const float pdrop = train ? c.dropout : 0.f, padrop = train ? c.attn_dropout : 0.f;
  (void)Ltxt;
  // 3. MLM transform head: x_out = LN_v(gelu(h W_t^T + b_t))
  RUNP(CLIPDLM_PROF_LN_BWD, 0, (double)T * 3 * D * (e->pair ? 4.0 : 2.0), layernorm_bwd_dispatch(&e->gv, &e->g0, param(e, CLIPDLM_P_VLN_W), c.ln_eps, T, D, &e->g1, grad(e, CLIPDLM_P_VLN_W),
                             grad(e, CLIPDLM_P_VLN_B), 0, 0, 0.f, nullptr, 0, 0.f, &e->uv, grad(e, CLIPDLM_P_VT_B), st));
  clipdlm_gemm_t g = linear_wgrad(e->g1, e->h[NL], T, D, D, grad(e, CLIPDLM_P_VT_W));
  RUNG(g);
  g = linear_dgrad(e->g1, shadow(e, CLIPDLM_P_VT_W), T, D, D, e->g0);
  RUNG(g);

  // 4. transformer blocks, last to first.  g0 holds d(block output).
  for (int l = NL - 1; l >= 0; --l) {
    const LayerBufs& b = e->lay[l];
    const Act& hin = e->h[l];
    const bool drop = pdrop > 0.f;
    // h_out = LN2(drop(ffn) + h1)
    RUNP(CLIPDLM_PROF_LN_BWD, 0, (double)T * 3 * D * (e->pair ? 4.0 : 2.0), layernorm_bwd_dispatch(&b.z2, &e->g0, param(e, lslot(l, CLIPDLM_PL_LN2_W)), c.ln_eps, T, D, &e->g1, grad(e, lslot(l, CLIPDLM_PL_LN2_W)),
                               grad(e, lslot(l, CLIPDLM_PL_LN2_B)), p.drop_seed, 0, 0.f, drop ? &e->g2 : nullptr, 2 + 2 * l, pdrop, nullptr,
                               grad(e, lslot(l, CLIPDLM_PL_FF2_B)), st));
    const Act& dffn = drop ? e->g2 : e->g1;  // gradient of the lin2 output
    g = linear_wgrad(dffn, b.g, T, D, F, grad(e, lslot(l, CLIPDLM_PL_FF2_W)));
    RUNG(g);
    g = linear_dgrad(dffn, shadow(e, lslot(l, CLIPDLM_PL_FF2_W)), T, D, F, e->gf);
    g.u_hi = b.u.hi; g.u_lo = b.u.lo; g.ldu = F;  // * gelu'(u)
    if (e->last_gelu_deriv) g.epilogue = CLIPDLM_EPI_STORE_MULAUX;   // ... which the forward already evaluated and stored
    const bool fused_bias = e->last_gelu_deriv && e->gelu_deriv >= 2;
    if (fused_bias) g.acc_f32 = grad(e, lslot(l, CLIPDLM_PL_FF1_B));   // d(lin1 bias) = column sums of gf, formed in that GEMM's epilogue
    RUNG(g);
    if (!fused_bias)
      RUNP(CLIPDLM_PROF_COLSUM, 0, (double)T * F * (e->pair ? 4.0 : 2.0), colsum_dispatch(&e->gf, T, F, grad(e, lslot(l, CLIPDLM_PL_FF1_B)), st));
    g = linear_wgrad(e->gf, b.h1, T, F, D, grad(e, lslot(l, CLIPDLM_PL_FF1_W)));
    RUNG(g);
    g = linear_dgrad(e->gf, shadow(e, lslot(l, CLIPDLM_PL_FF1_W)), T, F, D, e->g0);
    g.res_hi = e->g1.hi; g.res_lo = e->g1.lo; g.ldr = D;  // + residual branch d(h1)
    RUNG(g);
    // h1 = LN1(attn_out + h_in)
    RUNP(CLIPDLM_PROF_LN_BWD, 0, (double)T * 3 * D * (e->pair ? 4.0 : 2.0), layernorm_bwd_dispatch(&b.z1, &e->g0, param(e, lslot(l, CLIPDLM_PL_LN1_W)), c.ln_eps, T, D, &e->g1, grad(e, lslot(l, CLIPDLM_PL_LN1_W)),
                               grad(e, lslot(l, CLIPDLM_PL_LN1_B)), 0, 0, 0.f, nullptr, 0, 0.f, nullptr, grad(e, lslot(l, CLIPDLM_PL_O_B)), st));
    g = linear_wgrad(e->g1, b.ctx, T, D, D, grad(e, lslot(l, CLIPDLM_PL_O_W)));
    RUNG(g);
    g = linear_dgrad(e->g1, shadow(e, lslot(l, CLIPDLM_PL_O_W)), T, D, D, e->g0);
    RUNG(g);
    int bias_folded = 0;   // the packed tcgen05 backward (L = 16 / 18, plain bf16) forms d(q bias), d(v bias) in its epilogue; d(k bias) == 0
    RUNP(CLIPDLM_PROF_ATTN_BWD, 10.0 * R * L * L * D, (double)T * 7 * D * (e->pair ? 4.0 : 2.0),
         attn_bwd_dispatch(&b.qkv, e->keymask, &e->g0, R, L, D, c.n_heads, &e->gq, p.drop_seed, 1 + 2 * l, padrop, st, grad(e, lslot(l, CLIPDLM_PL_QKV_B)), &bias_folded));
    if (!bias_folded)
      RUNP(CLIPDLM_PROF_COLSUM, 0, (double)T * 3 * D * (e->pair ? 4.0 : 2.0), colsum_dispatch(&e->gq, T, 3 * D, grad(e, lslot(l, CLIPDLM_PL_QKV_B)), st));
    g = linear_wgrad(e->gq, hin, T, 3 * D, D, grad(e, lslot(l, CLIPDLM_PL_QKV_W)));
    RUNG(g);
    g = linear_dgrad(e->gq, shadow(e, lslot(l, CLIPDLM_PL_QKV_W)), T, 3 * D, D, e->g0);
    g.res_hi = e->g1.hi; g.res_lo = e->g1.lo; g.ldr = D;
    RUNG(g);
  }

  // 5. embeddings: h0 = drop(LN_e(z0)); z0 = fuse(x, CLIP projections) + segment + position
  RUNP(CLIPDLM_PROF_LN_BWD, 0, (double)T * 3 * D * (e->pair ? 4.0 : 2.0), layernorm_bwd_dispatch(&e->z0, &e->g0, param(e, CLIPDLM_P_EMB_LN_W), c.ln_eps, T, D, &e->g1, grad(e, CLIPDLM_P_EMB_LN_W),
                             grad(e, CLIPDLM_P_EMB_LN_B), p.drop_seed, 0, pdrop, nullptr, 0, 0.f, nullptr, nullptr, st));
  CLIPDLM_CUDA_OK(cudaMemsetAsync(e->d_img_proj, 0, (size_t)B * D * 4, st));
  CLIPDLM_CUDA_OK(cudaMemsetAsync(e->d_txt_proj, 0, (size_t)B * D * 4, st));
  RUN(embed_bwd_dispatch(&e->g1, R, B, Ltxt, L, D, c.fusion, p.guided, grad(e, CLIPDLM_P_POS), c.fusion == 0 ? grad(e, CLIPDLM_P_SEG) : nullptr,
                         e->d_img_proj, e->d_txt_proj, st));
  e->launches++;  // embed_bwd issues two kernels
  RUN(small_linear_bwd_dispatch(p.image_clip, e->d_img_proj, B, c.clip_dim, D, grad(e, CLIPDLM_P_IMG_W), grad(e, CLIPDLM_P_IMG_B), st));
  RUN(small_linear_bwd_dispatch(p.text_clip, e->d_txt_proj, B, c.clip_dim, D, grad(e, CLIPDLM_P_TXT_W), grad(e, CLIPDLM_P_TXT_B), st));
  return 0;
}

static int loss_backward_impl(clipdlm_engine* e, const clipdlm_loss_cfg_t* lc, double* losses, cudaStream_t st) {
  const clipdlm_config_t& c = e->cfg;
  CLIPDLM_CHECK(e->have_fwd, "loss_backward without a preceding forward");
  CLIPDLM_CHECK(lc != nullptr && losses != nullptr, "null loss config / output");
  CLIPDLM_CHECK(e->last.ids != nullptr || (lc->target && !lc->use_prob_loss), "loss needs the caption ids of the pass");
  CLIPDLM_CHECK(!lc->backward || (e->last.train || e->training), "backward needs a training engine");
  CLIPDLM_CHECK(!lc->backward || e->training, "engine was created without training buffers");
  const clipdlm_pass_t& p = e->last;
  const int R = p.R, B = p.B, L = e->L, Ltxt = e->Ltxt, D = c.dim;
  const int T = R * L, M16 = R * Ltxt;
  const bool bwd = lc->backward != 0;
  const long long R_total = lc->R_total > 0 ? lc->R_total : R;
  const bool mean_kind = lc->loss_kind == 0 || lc->loss_kind == 2;
  const double ce_scale = mean_kind ? 1.0 / (double)R_total : 1.0 / (double)lc->batch_size;  // CLIP-DDPM.py:437 vs :439-440

  // 1. embedding-space loss (+ its gradient written over every row of g0)
  if (lc->use_embed_loss || bwd)
    RUNP(CLIPDLM_PROF_LOSS, 0, ((double)M16 + (bwd ? T : 0)) * D * (e->pair ? 4.0 : 2.0) + (double)M16 * D * 4, embed_loss_dispatch(&e->xo, e->bufs.emb_table, p.ids, lc->target, lc->target_rows, R, B, Ltxt, L, D, lc->loss_kind, R_total, lc->batch_size,
                            lc->use_embed_loss ? 1.f : 0.f, lc->use_embed_loss ? &losses[0] : nullptr, bwd ? &e->g0 : nullptr, st));
  // 2. rounding cross-entropy through the frozen lm_head
  if (lc->use_prob_loss) {
    // plain bf16 + backward: the LSE pass keeps its logits (bf16, in the d(logits) buffer) and an HBM-bound pass turns them into the
    // softmax-CE gradient in place - 8 GB of traffic instead of recomputing the 3 TFLOP lm_head GEMM.  Split precision (parity
    // mode) recomputes the vocabulary tiles with fp32 accumulators instead (SMGRAD epilogue).
    const bool store_logits = bwd && !e->pair;
    // Option FUSED_SOFTMAX_GRAD: d x = sum_v softmax_v W_v - W_tgt with softmax_v = exp(s_v - c) / sum_u exp(s_u - c): the lm_head pass stores
    // exp(s - c), the 1 / sum factor scales the ROWS of the gradient GEMM's accumulator and the one-hot term is a gather of W rows - the
    // in-place pass over the 8 GB of stored logits (read + write) drops out of the step.
    const bool factored = store_logits && e->fused_smgrad != 0;
    int rc = lm_head_lse(e, p.ids, B * Ltxt, nullptr, &losses[1], ce_scale, st, store_logits ? e->dlog.hi : nullptr, factored);
    if (rc) return rc;
    if (bwd && factored) {
      const float gs = (float)(lc->rounding_weight * ce_scale);
      // row factors into the (now free) target-logit buffer; one-hot term folded into d(x_out) rows, which already hold the embedding-loss gradient
      RUNP(CLIPDLM_PROF_LOSS, 0, (double)M16 * D * 6.0,
           ce_row_terms_dispatch(e->lse, e->exp_shift, p.ids, B * Ltxt, gs, M16, e->bufs.emb_hi, D, e->g0.hi, D, Ltxt, L, D, e->tgt_logit, st,
                                 lc->rounding_weight_dev));
      Act emb{e->bufs.emb_hi, nullptr};
      clipdlm_gemm_t g = gemm_desc(e->dlog, e->ldl, 0, emb, D, 1, M16, D, c.vocab);
      g.epilogue = CLIPDLM_EPI_STORE_ROWSCALE;
      g.row_scale = e->tgt_logit;
      g.out_hi = e->g0.hi; g.ldo = D;
      g.res_hi = e->g0.hi; g.ldr = D;
      g.scatter_len = Ltxt; g.scatter_stride = L;
      RUNG(g);
    } else if (bwd) {
      Act emb{e->bufs.emb_hi, e->pair ? e->bufs.emb_lo : nullptr};
      clipdlm_gemm_t g;
      if (store_logits) {
        RUNP(CLIPDLM_PROF_GEMM_SMGRAD, 0, 4.0 * M16 * (double)e->ldl,
             softmax_grad_inplace_dispatch(e->dlog.hi, e->ldl, M16, c.vocab, e->lse, p.ids, B * Ltxt, (float)(lc->rounding_weight * ce_scale), st,
                                           lc->rounding_weight_dev));
      } else {
        g = gemm_desc(e->xo, D, 0, emb, D, 0, M16, c.vocab, D);
        g.gather_len = Ltxt; g.gather_stride = L;
        g.epilogue = CLIPDLM_EPI_SMGRAD;
        g.out_hi = e->dlog.hi; g.out_lo = e->dlog.lo; g.ldo = e->ldl;
        g.lse = e->lse; g.targets = p.ids; g.tgt_period = B * Ltxt;
        g.grad_scale = (float)(lc->rounding_weight * ce_scale);
        g.part_max = const_cast<float*>(lc->rounding_weight_dev);   // SMGRAD: optional device scalar multiplying grad_scale (clipdlm.h)
        RUNG(g);
      }
      // d x_out[:, :Ltxt] += dlogits[M16, V] E[V, D]
      g = gemm_desc(e->dlog, e->ldl, 0, emb, D, 1, M16, D, c.vocab);
      g.out_hi = e->g0.hi; g.out_lo = e->g0.lo; g.ldo = D;
      g.res_hi = e->g0.hi; g.res_lo = e->g0.lo; g.ldr = D;
      g.scatter_len = Ltxt; g.scatter_stride = L;
      RUNG(g);
    }
  }
  if (!bwd) return 0;
  // classifier-free guidance: d(x_out) belongs to the MIX of two passes - scale our share, hand the other engine its share
  if (lc->row_scale_self != nullptr || lc->export_engine != nullptr) {
    CLIPDLM_CHECK(lc->export_engine == nullptr || (lc->export_engine->training && lc->export_engine->have_fwd && lc->export_engine->last.R == R &&
                                                   lc->export_engine->pair == e->pair && lc->export_engine->L == L),
                  "loss_backward: export engine must hold a training forward of the same shape");
    RUNP(CLIPDLM_PROF_OTHER, 0, (double)T * D * (e->pair ? 4.0 : 2.0) * 3,
         row_scale_dispatch(&e->g0, lc->row_scale_self, lc->export_engine ? &lc->export_engine->g0 : nullptr, lc->row_scale_export, R, L, D, st));
  }
  return backward_from_g0(e, st);
}

static int lm_head_impl(clipdlm_engine* e, float* logits, int64_t ld_logits, int32_t* argmax, cudaStream_t st) {
  const clipdlm_config_t& c = e->cfg;
  CLIPDLM_CHECK(e->have_fwd, "lm_head without a preceding forward");
  const int M = e->last.R * e->Ltxt;
  if (argmax != nullptr) {
    int rc = lm_head_lse(e, nullptr, 1, argmax, nullptr, 0.0, st);
    if (rc) return rc;
  }
  if (logits != nullptr) {
    const int N32 = (c.vocab + 31) / 32 * 32;
    CLIPDLM_CHECK(ld_logits >= N32, "logits pitch %lld < vocab rounded up to 32 (%d)", (long long)ld_logits, N32);
    Act emb{e->bufs.emb_hi, e->pair ? e->bufs.emb_lo : nullptr};
    clipdlm_gemm_t g = gemm_desc(e->xo, c.dim, 0, emb, c.dim, 0, M, N32, c.dim);  // emb shadow is zero-padded to a 256-row multiple
    g.gather_len = e->Ltxt; g.gather_stride = e->L;
    g.out_f32 = logits; g.ldo = ld_logits;
    RUNG(g);
  }
  return 0;
}

}  // namespace clipdlm

extern "C" {

int64_t clipdlm_param_count(const clipdlm_config_t* cfg) { return cfg ? slot_offset(cfg, num_slots(cfg)) : -1; }
int64_t clipdlm_param_offset(const clipdlm_config_t* cfg, int32_t slot) {
  return (cfg && slot >= 0 && slot < num_slots(cfg)) ? slot_offset(cfg, slot) : -1;
}
int64_t clipdlm_param_size(const clipdlm_config_t* cfg, int32_t slot) { return (cfg && slot < num_slots(cfg)) ? slot_size(cfg, slot) : -1; }

size_t clipdlm_workspace_bytes(const clipdlm_config_t* cfg, int32_t max_rows, int32_t batch, int32_t training) {
  clipdlm_engine tmp;
  memset(&tmp, 0, sizeof(tmp));
  if (init_shape(&tmp, cfg, max_rows, batch, training)) return 0;
  tmp.h = new (std::nothrow) Act[cfg->n_layers + 1];
  tmp.lay = new (std::nothrow) LayerBufs[cfg->n_layers];
  size_t n = (tmp.h && tmp.lay) ? carve(&tmp, nullptr) : 0;
  delete[] tmp.h;
  delete[] tmp.lay;
  return n;
}

clipdlm_engine_t* clipdlm_engine_create(const clipdlm_config_t* cfg, const clipdlm_buffers_t* bufs, int32_t max_rows, int32_t batch,
                                        int32_t training) {
  if (bufs == nullptr) { set_last_error("null buffers"); return nullptr; }
  clipdlm_engine* e = new (std::nothrow) clipdlm_engine;
  if (!e) { set_last_error("out of host memory"); return nullptr; }
  memset(e, 0, sizeof(*e));
  if (init_shape(e, cfg, max_rows, batch, training)) { delete e; return nullptr; }
  e->bufs = *bufs;
  e->h = new (std::nothrow) Act[cfg->n_layers + 1];
  e->lay = new (std::nothrow) LayerBufs[cfg->n_layers];
  bool ok = e->h && e->lay;
  if (ok && !(bufs->params && bufs->shadow_hi && bufs->emb_table && bufs->emb_hi && bufs->workspace)) { set_last_error("engine buffers: null pointer"); ok = false; }
  if (ok && e->pair && !(bufs->shadow_lo && bufs->emb_lo)) { set_last_error("precision 1 (bf16x3) needs shadow_lo and emb_lo"); ok = false; }
  if (ok && training && !bufs->grads) { set_last_error("training engine needs a gradient buffer"); ok = false; }
  if (ok && (reinterpret_cast<uintptr_t>(bufs->workspace) & 255)) { set_last_error("workspace must be 256-byte aligned"); ok = false; }
  if (ok) {
    const size_t need = carve(e, (uint8_t*)bufs->workspace);
    if (need > bufs->workspace_bytes) { set_last_error("workspace too small: need %zu bytes, got %zu", need, bufs->workspace_bytes); ok = false; }
  }
  if (!ok) { delete[] e->h; delete[] e->lay; delete e; return nullptr; }
  e->prof = new (std::nothrow) Profiler;
  return e;
}

void clipdlm_engine_destroy(clipdlm_engine_t* e) {
  if (!e) return;
  if (e->prof) {
    for (auto& r : e->prof->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto ev : e->prof->pool) cudaEventDestroy(ev);
    delete e->prof;
  }
  delete[] e->h;
  delete[] e->lay;
  delete e;
}

int clipdlm_engine_forward(clipdlm_engine_t* e, const clipdlm_pass_t* p, clipdlm_stream stream) {
  CLIPDLM_CHECK(e != nullptr, "null engine");
  return forward_impl(e, p, (cudaStream_t)stream);
}
int clipdlm_engine_lm_head(clipdlm_engine_t* e, float* logits, int64_t ld_logits, int32_t* argmax, clipdlm_stream stream) {
  CLIPDLM_CHECK(e != nullptr, "null engine");
  return lm_head_impl(e, logits, ld_logits, argmax, (cudaStream_t)stream);
}
int clipdlm_engine_loss_backward(clipdlm_engine_t* e, const clipdlm_loss_cfg_t* lc, double* losses, clipdlm_stream stream) {
  CLIPDLM_CHECK(e != nullptr, "null engine");
  return loss_backward_impl(e, lc, losses, (cudaStream_t)stream);
}
int clipdlm_engine_cfg_mix(clipdlm_engine_t* eu, clipdlm_engine_t* eg, const int32_t* guided, float w, clipdlm_stream stream) {
  CLIPDLM_CHECK(eu != nullptr && eg != nullptr && guided != nullptr, "cfg_mix: null argument");
  CLIPDLM_CHECK(eu->have_fwd && eg->have_fwd && eu->last.R == eg->last.R && eu->L == eg->L && eu->cfg.dim == eg->cfg.dim && eu->pair == eg->pair,
                "cfg_mix: the two engines must hold forward passes of the same shape");
  clipdlm_engine* e = eu;
  cudaStream_t st = (cudaStream_t)stream;
  RUNP(CLIPDLM_PROF_OTHER, 0, 0, cfg_mix_dispatch(&eu->xo, &eg->xo, guided, w, eu->last.R, eu->L, eu->cfg.dim, st));
  return 0;
}
int clipdlm_engine_backward(clipdlm_engine_t* e, clipdlm_stream stream) {
  CLIPDLM_CHECK(e != nullptr && e->have_fwd && e->training, "engine_backward: needs a training engine with a forward pass");
  return backward_from_g0(e, (cudaStream_t)stream);
}
int clipdlm_engine_backward_from(clipdlm_engine_t* e, const float* dx_out, float* dx_in, clipdlm_stream stream) {
  CLIPDLM_CHECK(e != nullptr && e->have_fwd && e->training, "engine_backward_from: needs a training engine with a forward pass");
  CLIPDLM_CHECK(dx_out != nullptr, "engine_backward_from: null upstream gradient");
  cudaStream_t st = (cudaStream_t)stream;
  const long long T = (long long)e->last.R * e->L;
  RUN(to_bf16_dispatch(dx_out, e->g0.hi, e->pair ? e->g0.lo : nullptr, T * e->cfg.dim, st));
  int rc = backward_from_g0(e, st);
  if (rc) return rc;
  // g1 now holds d(z0), the gradient of the pre-LayerNorm embedding sum; the caller's x_in enters z0 additively at the text positions
  if (dx_in != nullptr) RUN(gather_rows_f32_dispatch(&e->g1, (long long)e->last.R * e->Ltxt, e->Ltxt, e->L, e->cfg.dim, dx_in, st));
  return 0;
}
int clipdlm_engine_set_option(clipdlm_engine_t* e, int32_t option, int64_t value) {
  CLIPDLM_CHECK(e != nullptr, "set_option: null engine");
  switch (option) {
    case CLIPDLM_OPT_FUSED_SOFTMAX_GRAD:
      CLIPDLM_CHECK(value == 0 || !e->pair, "FUSED_SOFTMAX_GRAD needs plain-bf16 precision (split precision recomputes the vocabulary tiles in fp32)");
      e->fused_smgrad = value != 0;
      return 0;
    case CLIPDLM_OPT_EXP_SHIFT_PTR:
      e->exp_shift = reinterpret_cast<const float*>(static_cast<uintptr_t>(value));
      return 0;
    case CLIPDLM_OPT_GELU_DERIV_STORE:
      CLIPDLM_CHECK(value == 0 || !e->pair, "GELU_DERIV_STORE needs plain-bf16 precision");
      CLIPDLM_CHECK(value >= 0 && value <= 2, "GELU_DERIV_STORE: value 0, 1 or 2");
      e->gelu_deriv = (int)value;
      return 0;
    default:
      CLIPDLM_CHECK(false, "set_option: unknown option %d", (int)option);
  }
  return 0;
}

int64_t clipdlm_engine_launch_count(const clipdlm_engine_t* e) { return e ? e->launches : -1; }

int clipdlm_engine_profile(clipdlm_engine_t* e, int32_t enable) {
  CLIPDLM_CHECK(e != nullptr && e->prof != nullptr, "null engine");
  e->prof->on = enable != 0;
  return 0;
}
int clipdlm_engine_profile_read(clipdlm_engine_t* e, clipdlm_prof_t* out, int32_t reset) {
  CLIPDLM_CHECK(e != nullptr && e->prof != nullptr && out != nullptr, "null engine / output");
  for (int c = 0; c < CLIPDLM_PROF_NCAT; ++c) { out[c].ms = 0; out[c].flops = 0; out[c].bytes = 0; out[c].launches = 0; }
  for (auto& r : e->prof->recs) {
    CLIPDLM_CUDA_OK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CLIPDLM_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
    out[r.cat].ms += ms; out[r.cat].flops += r.flops; out[r.cat].bytes += r.bytes; out[r.cat].launches += 1;
  }
  if (reset) {
    for (auto& r : e->prof->recs) { e->prof->pool.push_back(r.a); e->prof->pool.push_back(r.b); }
    e->prof->recs.clear();
  }
  return 0;
}

}  // extern "C"
