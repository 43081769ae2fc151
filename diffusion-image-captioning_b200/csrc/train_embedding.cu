// TRAIN_EMBEDDING=True branch of the hot path (CLIP-DDPM.py:238-243,292-293,319-320): a learned IN_CHANNEL-wide (16) token
// embedding with trainable input / output projections around the encoder and a trainable lm_head nn.Linear(IN_CHANNEL, V).
// The encoder itself is the engine (engine.cu); the K = 16 projections run through the fp32 small-linear kernels; the lm_head
// (logits / LSE / softmax gradient / dgrad / wgrad) runs through the tcgen05 GEMM on operands zero-padded to 64 channels.
// This file holds the narrow-channel HBM kernels that glue those pieces:
//   feature_loss_f32     LOSS_FUNC (CLIP-DDPM.py:77-89) on fp32 [R, L, ch] features: value, d(features), d(target)
//   pack_rows_bf16       y[:, :Ltxt, :ch] fp32 -> compact bf16 (pair) [R * Ltxt, ld] zero-padded GEMM operand
//   embedding_bwd        scatter-add of the gradient that reaches embedding.weight through q_sample (x_t = a E[ids] + b eps) / x_0
#include "common.cuh"
#include "../../include/clipdlm.h"

namespace clipdlm {

int num_sms();

// One warp per sequence row r. Text elements of the row: e in [0, Ltxt * ch): position p = e / ch, channel c = e % ch.
//   y      [R, L, ch] fp32 (model feature_out, CLIP-DDPM.py:322)
//   target [target_rows, Ltxt, ch] fp32, row r uses target[r % target_rows]  (x_0.repeat / x_tgt, :418,421,428)
//   dce    optional [R * Ltxt, ld_dce] fp32: gradient arriving from the rounding cross-entropy, added to the text positions
//   dy     optional [R, L, ch]: written (text positions: loss gradient + dce; other positions: 0)
//   d_target optional [target_rows, Ltxt, ch]: -= loss gradient (atomic; the target carries gradient when the embedding is learned)
__global__ void __launch_bounds__(256) feature_loss_f32_kernel(const float* __restrict__ y, const float* __restrict__ target, int target_rows,
                                                               int R, int Ltxt, int L, int ch, int kind, double inv_div, float gscale,
                                                               double* loss_acc, const float* __restrict__ dce, int ld_dce,
                                                               float* __restrict__ dy, float* __restrict__ d_target) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int n = Ltxt * ch;
  const float* yr = y + (size_t)r * L * ch;
  const float* tr = target + (size_t)(r % target_rows) * n;
  float acc = 0.f;
  for (int e = lane; e < n; e += 32) {
    const float d = yr[e] - tr[e];
    acc += kind <= 1 ? fabsf(d) : d * d;
  }
  acc = warp_sum(acc);
  float gs = gscale;
  double val = (double)acc;
  if (kind >= 2) {
    const float norm = sqrtf(acc);
    gs = norm > 0.f ? gscale / norm : 0.f;
    val = (double)norm;
  }
  if (lane == 0 && loss_acc != nullptr) atomicAdd(loss_acc, val * inv_div);
  if (dy == nullptr && d_target == nullptr) return;
  for (int e = lane; e < n; e += 32) {
    const float d = yr[e] - tr[e];
    const float g = kind <= 1 ? (d > 0.f ? gs : (d < 0.f ? -gs : 0.f)) : d * gs;
    if (dy != nullptr) {
      const int p = e / ch, c = e % ch;
      dy[(size_t)r * L * ch + e] = g + (dce != nullptr ? dce[((size_t)r * Ltxt + p) * ld_dce + c] : 0.f);
    }
    if (d_target != nullptr && g != 0.f) atomicAdd(d_target + (size_t)(r % target_rows) * n + e, -g);
  }
  if (dy != nullptr)
    for (int e = n + lane; e < L * ch; e += 32) dy[(size_t)r * L * ch + e] = 0.f;
}

__global__ void pack_rows_bf16_kernel(const float* __restrict__ y, long long rows_out, int Ltxt, int L, int ch, int ld,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = rows_out * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / ld;
    const int c = (int)(i % ld);
    float v = 0.f;
    if (c < ch) v = y[((m / Ltxt) * L + (m % Ltxt)) * ch + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo != nullptr) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// dE[ids[j], c] += sum_s scale[s] * dx[s, j, c]   (j over the B * Ltxt tokens of the batch; dx [S, B * Ltxt, ch])
__global__ void embedding_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ scale, const int* __restrict__ ids, int S,
                                     long long tokens, int ch, float* __restrict__ dE) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tokens * ch) return;
  const long long j = i / ch;
  const int c = (int)(i % ch);
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc = fmaf(scale != nullptr ? scale[s] : 1.f, dx[((size_t)s * tokens + j) * ch + c], acc);
  if (acc != 0.f) atomicAdd(dE + (size_t)ids[j] * ch + c, acc);
}

// classifier-free guidance on fp32 encoder outputs (CLIP-DDPM.py:313-317): xu[r] <- guided[r] ? (1 + w) xg[r] - w xu[r] : xu[r]
__global__ void row_mix_f32_kernel(float* __restrict__ xu, const float* __restrict__ xg, const int* __restrict__ guided, float w,
                                   long long rowlen4, long long total4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    if (!guided[i / rowlen4]) continue;
    const float4 a = reinterpret_cast<const float4*>(xu)[i], b = reinterpret_cast<const float4*>(xg)[i];
    reinterpret_cast<float4*>(xu)[i] = make_float4((1.f + w) * b.x - w * a.x, (1.f + w) * b.y - w * a.y, (1.f + w) * b.z - w * a.z, (1.f + w) * b.w - w * a.w);
  }
}
// its backward: the gradient of the mixed output goes (1 + w) to the guided pass and -w (1 on plain rows) to the unguided pass
__global__ void row_split_f32_kernel(float* __restrict__ d_self, float* __restrict__ d_other, const int* __restrict__ guided, float w,
                                     long long rowlen4, long long total4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(d_self)[i];
    const bool gd = guided[i / rowlen4] != 0;
    const float so = gd ? 1.f + w : 0.f, ss = gd ? -w : 1.f;
    reinterpret_cast<float4*>(d_other)[i] = make_float4(so * g.x, so * g.y, so * g.z, so * g.w);
    reinterpret_cast<float4*>(d_self)[i] = make_float4(ss * g.x, ss * g.y, ss * g.z, ss * g.w);
  }
}
__global__ void add_f32_kernel(float* __restrict__ a, const float* __restrict__ b, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    reinterpret_cast<float4*>(a)[i] = x;
  }
}
static inline unsigned grid_for(long long n, int per_block = 256) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = 16LL * num_sms();
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int feature_loss_f32_dispatch(const float* y, const float* target, int target_rows, int R, int Ltxt, int L, int ch, int kind,
                                     long long R_total, int batch_size, float weight, double* loss_acc, const float* dce, int ld_dce, float* dy,
                                     float* d_target, cudaStream_t st) {
  CLIPDLM_CHECK(y && target && target_rows > 0 && R > 0 && Ltxt > 0 && L >= Ltxt && ch > 0 && kind >= 0 && kind <= 3,
                "feature_loss_f32: bad arguments");
  CLIPDLM_CHECK(dce == nullptr || ld_dce >= ch, "feature_loss_f32: dce pitch %d < channels %d", ld_dce, ch);
  double inv_div;
  switch (kind) {
    case 0: inv_div = 1.0 / ((double)R_total * ch); break;                 // .abs().sum(dim=1).mean() over [R, ch]      CLIP-DDPM.py:77-78
    case 1: inv_div = 1.0 / ((double)batch_size * 768.0 * 100.0); break;   // .abs().sum() / BATCH_SIZE / 768 / 100      :80-81 (literals)
    case 2: inv_div = 1.0 / (double)R_total; break;                        // per-row L2 norm, mean                      :83-84
    default: inv_div = 1.0 / (double)batch_size; break;                    // per-row L2 norm, sum / BATCH_SIZE          :86-87
  }
  feature_loss_f32_kernel<<<(R + 7) / 8, 256, 0, st>>>(y, target, target_rows, R, Ltxt, L, ch, kind, inv_div, (float)(inv_div * weight),
                                                       loss_acc, dce, ld_dce, dy, d_target);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace clipdlm

using namespace clipdlm;

extern "C" {

int clipdlm_feature_loss_f32(const float* y, const float* target, int32_t target_rows, int32_t R, int32_t Ltxt, int32_t L, int32_t ch,
                             int32_t kind, int64_t R_total, int32_t batch_size, float weight, double* loss_acc, const float* dce,
                             int32_t ld_dce, float* dy, float* d_target, clipdlm_stream stream) {
  return feature_loss_f32_dispatch(y, target, target_rows, R, Ltxt, L, ch, kind, R_total, batch_size, weight, loss_acc, dce, ld_dce, dy,
                                   d_target, (cudaStream_t)stream);
}

int clipdlm_pack_rows_bf16(const float* y, int64_t rows_out, int32_t Ltxt, int32_t L, int32_t ch, int32_t ld, void* hi, void* lo,
                           clipdlm_stream stream) {
  CLIPDLM_CHECK(y && hi && rows_out > 0 && Ltxt > 0 && L >= Ltxt && ch > 0 && ld >= ch && rows_out % Ltxt == 0, "pack_rows_bf16: bad arguments");
  const long long total = (long long)rows_out * ld;
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  pack_rows_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, rows_out, Ltxt, L, ch, ld, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int clipdlm_row_mix_f32(float* x_unguided, const float* x_guided, const int32_t* guided, float w, int32_t R, int64_t row_len,
                        clipdlm_stream stream) {
  CLIPDLM_CHECK(x_unguided && x_guided && guided && R > 0 && row_len > 0 && row_len % 4 == 0, "row_mix_f32: bad arguments");
  const long long total4 = (long long)R * row_len / 4;
  row_mix_f32_kernel<<<grid_for(total4), 256, 0, (cudaStream_t)stream>>>(x_unguided, x_guided, guided, w, row_len / 4, total4);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int clipdlm_row_split_f32(float* d_self, float* d_other, const int32_t* guided, float w, int32_t R, int64_t row_len, clipdlm_stream stream) {
  CLIPDLM_CHECK(d_self && d_other && guided && R > 0 && row_len > 0 && row_len % 4 == 0, "row_split_f32: bad arguments");
  const long long total4 = (long long)R * row_len / 4;
  row_split_f32_kernel<<<grid_for(total4), 256, 0, (cudaStream_t)stream>>>(d_self, d_other, guided, w, row_len / 4, total4);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int clipdlm_add_f32(float* a, const float* b, int64_t n, clipdlm_stream stream) {
  CLIPDLM_CHECK(a && b && n > 0 && n % 4 == 0, "add_f32: bad arguments");
  add_f32_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(a, b, n / 4);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int clipdlm_embedding_bwd(const float* dx, const float* scale, const int32_t* ids, int32_t S, int64_t tokens, int32_t ch, float* d_table,
                          clipdlm_stream stream) {
  CLIPDLM_CHECK(dx && ids && d_table && S > 0 && tokens > 0 && ch > 0, "embedding_bwd: bad arguments");
  const long long n = (long long)tokens * ch;
  embedding_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dx, scale, ids, S, tokens, ch, d_table);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
