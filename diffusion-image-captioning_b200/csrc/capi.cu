// extern "C" surface of libclipdlm.so (declared in include/clipdlm.h): thin, exception-free forwarding
// to the kernel dispatchers. No torch types cross this boundary.
#include "common.cuh"
#include "../../include/clipdlm.h"
#include <stdarg.h>
#include <string.h>

namespace clipdlm {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int gemm_dispatch(const clipdlm_gemm_t* g, cudaStream_t st);
int lse_combine_dispatch(const float* pmax, const float* psum, const int* parg, int n_tiles, int M, const float* tgt_logit, float* lse,
                         int* argmax, double* loss_acc, double scale, cudaStream_t st);
int ce_row_terms_dispatch(const float* lse, const float* exp_shift, const int* targets, int tgt_period, float scale, int M, const void* w,
                          long long ldw, void* dx, long long ldx, int scatter_len, int scatter_stride, int D, float* row_scale, cudaStream_t st,
                          const float* scale_mul);
void gemm_debug_mn_desc(uint32_t lbo, uint32_t sbo);
void gemm_debug_flags(uint32_t flags);
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
int embed_loss_dispatch(const clipdlm_bf_t* x_out, const float* emb, const int* ids, const float* tgt, int tgt_rows, int R, int B, int Ltxt,
                        int L, int D, int kind, long long R_total, int batch_size, float weight, double* loss_acc, const clipdlm_bf_t* dx,
                        cudaStream_t st);
int small_linear_fwd_dispatch(const float* x, const float* w, const float* b, int B, int K, int N, float* y, cudaStream_t st);
int small_linear_bwd_dispatch(const float* x, const float* dy, int B, int K, int N, float* dw, float* db, cudaStream_t st);
int adamw_dispatch(float* p, float* g, float* m, float* v, void* sh_hi, void* sh_lo, long long n, float lr, float beta1, float beta2,
                   float eps, float wd, int step, float grad_scale, int zero_grad, cudaStream_t st);
int to_bf16_dispatch(const float* x, void* hi, void* lo, long long n, cudaStream_t st);
int to_f32_dispatch(const void* hi, const void* lo, float* y, long long n, cudaStream_t st);
int gather_rows_f32_dispatch(const clipdlm_bf_t* x, long long rows_out, int len, int stride, int D, float* y, cudaStream_t st);
int q_sample_dispatch(const float* x0, const float* noise, const float* ca, const float* cb, long long n, int S, float* out, cudaStream_t st);
int keymask_dispatch(const int* attn_mask, int R, int B, int Ltxt, int L, int fusion, int guided, uint32_t* km, cudaStream_t st);
void attn_force_simt(int on);
int attn_fwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, int R, int L, int D, int H, const clipdlm_bf_t* ctx,
                      unsigned long long seed, uint32_t site, float p, cudaStream_t st);
int attn_bwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int R, int L, int D, int H,
                      const clipdlm_bf_t* dqkv, unsigned long long seed, uint32_t site, float p, cudaStream_t st, float* dbias, int* folded);

}  // namespace clipdlm

using namespace clipdlm;

extern "C" {

const char* clipdlm_last_error(void) { return g_last_error; }
int clipdlm_version(void) { return 100; }

int clipdlm_device_ok(void) {
  int dev = 0, major = 0;
  CLIPDLM_CUDA_OK(cudaGetDevice(&dev));
  CLIPDLM_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}

int clipdlm_gemm(const clipdlm_gemm_t* g, clipdlm_stream stream) { return gemm_dispatch(g, (cudaStream_t)stream); }

void clipdlm_gemm_debug_mn_desc(uint32_t lbo_bytes, uint32_t sbo_bytes) { gemm_debug_mn_desc(lbo_bytes, sbo_bytes); }
void clipdlm_gemm_debug_flags(uint32_t flags) { gemm_debug_flags(flags); }

int clipdlm_lse_combine(const float* part_max, const float* part_sum, const int32_t* part_arg, int32_t n_tiles, int32_t M,
                        const float* tgt_logit, float* lse, int32_t* argmax, double* loss_acc, double scale, clipdlm_stream stream) {
  return lse_combine_dispatch(part_max, part_sum, part_arg, n_tiles, M, tgt_logit, lse, argmax, loss_acc, scale, (cudaStream_t)stream);
}

int clipdlm_ce_row_terms(const float* lse, const float* exp_shift, const int32_t* targets, int32_t tgt_period, float scale, int32_t M,
                         const void* w_bf16, int64_t ldw, void* dx_bf16, int64_t ldx, int32_t scatter_len, int32_t scatter_stride, int32_t D,
                         float* row_scale, clipdlm_stream stream) {
  return ce_row_terms_dispatch(lse, exp_shift, targets, tgt_period, scale, M, w_bf16, ldw, dx_bf16, ldx, scatter_len, scatter_stride, D, row_scale,
                               (cudaStream_t)stream, nullptr);
}

#define ST ((cudaStream_t)stream)
int clipdlm_embed_fwd(const clipdlm_embed_t* e, clipdlm_stream stream) { return embed_fwd_dispatch(e, ST); }
int clipdlm_embed_bwd(const clipdlm_bf_t* dz, int32_t R, int32_t B, int32_t Ltxt, int32_t L, int32_t D, int32_t fusion, int32_t guided,
                      float* d_pos, float* d_seg, float* d_img_proj, float* d_txt_proj, clipdlm_stream stream) {
  return embed_bwd_dispatch(dz, R, B, Ltxt, L, D, fusion, guided, d_pos, d_seg, d_img_proj, d_txt_proj, ST);
}
int clipdlm_layernorm_fwd(const clipdlm_bf_t* z, const float* w, const float* b, float eps, int64_t rows, int32_t D, const clipdlm_bf_t* y,
                          float* y_f32, uint64_t drop_seed, uint32_t drop_site, float drop_p, clipdlm_stream stream) {
  return layernorm_fwd_dispatch(z, w, b, eps, rows, D, y, y_f32, drop_seed, drop_site, drop_p, ST);
}
int clipdlm_layernorm_bwd(const clipdlm_bf_t* z, const clipdlm_bf_t* dy, const float* w, float eps, int64_t rows, int32_t D,
                          const clipdlm_bf_t* dz, float* dw, float* db, uint64_t drop_seed, uint32_t drop_site_out, float drop_p_out,
                          const clipdlm_bf_t* dz_drop, uint32_t drop_site_in, float drop_p_in, const clipdlm_bf_t* gelu_u, float* dbias,
                          clipdlm_stream stream) {
  return layernorm_bwd_dispatch(z, dy, w, eps, rows, D, dz, dw, db, drop_seed, drop_site_out, drop_p_out, dz_drop, drop_site_in, drop_p_in,
                                gelu_u, dbias, ST);
}
int clipdlm_attn_fwd(const clipdlm_bf_t* qkv, const uint32_t* keymask, int32_t R, int32_t L, int32_t D, int32_t H, const clipdlm_bf_t* ctx,
                     uint64_t drop_seed, uint32_t drop_site, float drop_p, clipdlm_stream stream) {
  return attn_fwd_dispatch(qkv, keymask, R, L, D, H, ctx, drop_seed, drop_site, drop_p, ST);
}
int clipdlm_attn_bwd(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int32_t R, int32_t L, int32_t D, int32_t H,
                     const clipdlm_bf_t* dqkv, uint64_t drop_seed, uint32_t drop_site, float drop_p, clipdlm_stream stream) {
  return attn_bwd_dispatch(qkv, keymask, dctx, R, L, D, H, dqkv, drop_seed, drop_site, drop_p, ST, nullptr, nullptr);
}
int clipdlm_attn_bwd_bias(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int32_t R, int32_t L, int32_t D, int32_t H,
                          const clipdlm_bf_t* dqkv, uint64_t drop_seed, uint32_t drop_site, float drop_p, float* dbias_qkv, int32_t* bias_folded,
                          clipdlm_stream stream) {
  int folded = 0;
  const int rc = attn_bwd_dispatch(qkv, keymask, dctx, R, L, D, H, dqkv, drop_seed, drop_site, drop_p, ST, dbias_qkv, &folded);
  if (bias_folded) *bias_folded = folded;
  return rc;
}
void clipdlm_attn_force_simt(int32_t on) { attn_force_simt(on); }
int clipdlm_colsum(const clipdlm_bf_t* x, int64_t rows, int32_t N, float* out, clipdlm_stream stream) { return colsum_dispatch(x, rows, N, out, ST); }
int clipdlm_embed_loss(const clipdlm_bf_t* x_out, const float* emb_table, const int32_t* ids, const float* target, int32_t target_rows,
                       int32_t R, int32_t B, int32_t Ltxt, int32_t L,
                       int32_t D, int32_t kind, int64_t R_total, int32_t batch_size, float weight, double* loss_acc, const clipdlm_bf_t* dx,
                       clipdlm_stream stream) {
  return embed_loss_dispatch(x_out, emb_table, ids, target, target_rows, R, B, Ltxt, L, D, kind, R_total, batch_size, weight, loss_acc, dx, ST);
}
int clipdlm_small_linear_fwd(const float* x, const float* w, const float* b, int32_t B, int32_t K, int32_t N, float* y, clipdlm_stream stream) {
  return small_linear_fwd_dispatch(x, w, b, B, K, N, y, ST);
}
int clipdlm_small_linear_bwd(const float* x, const float* dy, int32_t B, int32_t K, int32_t N, float* dw, float* db, clipdlm_stream stream) {
  return small_linear_bwd_dispatch(x, dy, B, K, N, dw, db, ST);
}
int clipdlm_adamw(float* p, float* g, float* m, float* v, void* shadow_hi, void* shadow_lo, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int32_t step, float grad_scale, int32_t zero_grad, clipdlm_stream stream) {
  return adamw_dispatch(p, g, m, v, shadow_hi, shadow_lo, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, zero_grad, ST);
}
int clipdlm_to_bf16(const float* x, void* hi, void* lo, int64_t n, clipdlm_stream stream) { return to_bf16_dispatch(x, hi, lo, n, ST); }
int clipdlm_to_f32(const void* hi, const void* lo, float* y, int64_t n, clipdlm_stream stream) { return to_f32_dispatch(hi, lo, y, n, ST); }
int clipdlm_gather_rows_f32(const clipdlm_bf_t* x, int64_t rows_out, int32_t len, int32_t stride, int32_t D, float* y, clipdlm_stream stream) {
  return gather_rows_f32_dispatch(x, rows_out, len, stride, D, y, ST);
}
int clipdlm_keymask(const int32_t* attn_mask, int32_t R, int32_t B, int32_t Ltxt, int32_t L, int32_t fusion, int32_t guided, uint32_t* keymask,
                    clipdlm_stream stream) {
  return keymask_dispatch(attn_mask, R, B, Ltxt, L, fusion, guided, keymask, ST);
}
int clipdlm_q_sample(const float* x0, const float* noise, const float* coef_a, const float* coef_b, int64_t n, int32_t S, float* out,
                     clipdlm_stream stream) {
  return q_sample_dispatch(x0, noise, coef_a, coef_b, n, S, out, ST);
}
#undef ST

}  // extern "C"
