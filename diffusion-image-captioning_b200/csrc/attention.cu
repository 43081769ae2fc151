// Multi-head self-attention for very short sequences (L = MAX_LENGTH + 2 = 18 on the reference path, <= 128 supported),
// head dim 64. Replaces DistilBertSelfAttention / SDPA (HF modeling_distilbert.py:126-151,177-207) and its autograd backward.
//
// One warp owns one (sequence row r, head h): q, k, v of that head (L x 64 each) are staged once in shared memory with
// coalesced 16-byte loads, scores are computed with lane j owning key j, softmax / dropout statistics by warp shuffles,
// and the P.V products with lane d owning output dims (d, d + 32).  Nothing but the context (fwd) or d(qkv) (bwd) is
// written to HBM: the backward recomputes the probabilities from q, k and the key mask, and regenerates the dropout mask
// from the counter-based RNG, so no [R, H, L, L] tensor ever exists in memory.
#include "common.cuh"
#include "../../include/clipdlm.h"

namespace clipdlm {

int num_sms();
DropoutCfg make_drop(unsigned long long seed, uint32_t site, float p);

constexpr int DH = 64;
constexpr int ROWP = 65;  // smem row pitch in floats: conflict-free for both "lane = key row" and "lane = column" access

// 128 random bits shared by the 8 query rows i0..i0+7 for key j of (row, head) pair rh; field (i & 7) is 16 bits wide.
__device__ __forceinline__ uint4 attn_rand(const DropoutCfg& d, unsigned long long rh, int igroup, int j) {
  return philox4x32_10(make_uint4((uint32_t)rh, (uint32_t)(rh >> 32), (uint32_t)igroup | ((uint32_t)j << 16), d.site ^ 0xa77e0000u),
                       make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
}
__device__ __forceinline__ bool attn_keep(const DropoutCfg& d, const uint4& rnd, int i) {
  const int f = i & 7;
  const uint32_t word = (f >> 1) == 0 ? rnd.x : ((f >> 1) == 1 ? rnd.y : ((f >> 1) == 2 ? rnd.z : rnd.w));
  const uint32_t v16 = (f & 1) ? (word >> 16) : (word & 0xffffu);
  return v16 >= d.thresh16;
}

__device__ __forceinline__ void stage_head(const CBfPtr& src, size_t elem_off, float* dst) {
  float t[8];
  load8(src, elem_off, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = t[i];
}

template <int KG>
__global__ void __launch_bounds__(128) attn_fwd_kernel(CBfPtr qkv, const uint32_t* __restrict__ keymask, int R, int L, int D, int H, BfPtr ctx,
                                                       DropoutCfg drop, float scale, int warps_per_block) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rh = (long long)blockIdx.x * warps_per_block + warp;
  if (rh >= (long long)R * H) return;
  const int r = (int)(rh / H), h = (int)(rh % H);
  float* q = smem + (size_t)warp * 3 * L * ROWP;
  float* k = q + L * ROWP;
  float* v = k + L * ROWP;
  for (int idx = lane; idx < L * 8; idx += 32) {
    const int i = idx >> 3, part = idx & 7;
    const size_t base = ((size_t)r * L + i) * 3 * D + h * DH + part * 8;
    stage_head(qkv, base, q + i * ROWP + part * 8);
    stage_head(qkv, base + D, k + i * ROWP + part * 8);
    stage_head(qkv, base + 2 * D, v + i * ROWP + part * 8);
  }
  __syncwarp();
  const int kw = (L + 31) / 32;
  bool valid[KG];
#pragma unroll
  for (int kg = 0; kg < KG; ++kg) {
    const int j = lane + 32 * kg;
    valid[kg] = j < L && ((keymask[(size_t)r * kw + kg] >> lane) & 1u);
  }
  uint4 rnd[KG];
  for (int i = 0; i < L; ++i) {
    float s[KG];
    float mx = -INFINITY;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      const int j = lane + 32 * kg;
      float acc = -INFINITY;
      if (valid[kg]) {
        acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) acc += q[i * ROWP + d] * k[j * ROWP + d];
        acc *= scale;
      }
      s[kg] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) { s[kg] = valid[kg] ? __expf(s[kg] - mx) : 0.f; sum += s[kg]; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    if (drop.thresh16 != 0 && (i & 7) == 0) {
#pragma unroll
      for (int kg = 0; kg < KG; ++kg) rnd[kg] = attn_rand(drop, (unsigned long long)rh, i >> 3, lane + 32 * kg);
    }
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      float a = s[kg] * inv;
      if (drop.thresh16 != 0) a = attn_keep(drop, rnd[kg], i) ? a * drop.scale : 0.f;
      s[kg] = a;
    }
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      const int nj = min(32, L - 32 * kg);
      for (int jj = 0; jj < nj; ++jj) {
        const float a = __shfl_sync(0xffffffffu, s[kg], jj);
        const float* vr = v + (jj + 32 * kg) * ROWP;
        o0 += a * vr[lane];
        o1 += a * vr[lane + 32];
      }
    }
    __syncwarp();             // every lane has finished reading q row i
    q[i * ROWP + lane] = o0;  // reuse q row i as the output staging row
    q[i * ROWP + lane + 32] = o1;
  }
  __syncwarp();
  for (int idx = lane; idx < L * 8; idx += 32) {
    const int i = idx >> 3, part = idx & 7;
    float t[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = q[i * ROWP + part * 8 + e];
    store8(ctx, ((size_t)r * L + i) * D + h * DH + part * 8, t);
  }
}

template <int KG>
__global__ void __launch_bounds__(128) attn_bwd_kernel(CBfPtr qkv, const uint32_t* __restrict__ keymask, CBfPtr dctx, int R, int L, int D, int H,
                                                       BfPtr dqkv, DropoutCfg drop, float scale, int warps_per_block) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rh = (long long)blockIdx.x * warps_per_block + warp;
  if (rh >= (long long)R * H) return;
  const int r = (int)(rh / H), h = (int)(rh % H);
  float* q = smem + (size_t)warp * 6 * L * ROWP;
  float* k = q + L * ROWP;
  float* v = k + L * ROWP;
  float* dO = v + L * ROWP;   // d(ctx) rows; row i is overwritten by dq_i once consumed
  float* dk = dO + L * ROWP;
  float* dv = dk + L * ROWP;
  for (int idx = lane; idx < L * 8; idx += 32) {
    const int i = idx >> 3, part = idx & 7;
    const size_t base = ((size_t)r * L + i) * 3 * D + h * DH + part * 8;
    stage_head(qkv, base, q + i * ROWP + part * 8);
    stage_head(qkv, base + D, k + i * ROWP + part * 8);
    stage_head(qkv, base + 2 * D, v + i * ROWP + part * 8);
    stage_head(dctx, ((size_t)r * L + i) * D + h * DH + part * 8, dO + i * ROWP + part * 8);
  }
  for (int idx = lane; idx < L * ROWP; idx += 32) { dk[idx] = 0.f; dv[idx] = 0.f; }
  __syncwarp();
  const int kw = (L + 31) / 32;
  bool valid[KG];
#pragma unroll
  for (int kg = 0; kg < KG; ++kg) {
    const int j = lane + 32 * kg;
    valid[kg] = j < L && ((keymask[(size_t)r * kw + kg] >> lane) & 1u);
  }
  uint4 rnd[KG];
  for (int i = 0; i < L; ++i) {
    float p[KG], dA[KG];
    float mx = -INFINITY;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      const int j = lane + 32 * kg;
      float acc = -INFINITY, acc2 = 0.f;
      if (valid[kg]) {
        acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) {
          acc += q[i * ROWP + d] * k[j * ROWP + d];
          acc2 += dO[i * ROWP + d] * v[j * ROWP + d];
        }
        acc *= scale;
      }
      p[kg] = acc; dA[kg] = acc2;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) { p[kg] = valid[kg] ? __expf(p[kg] - mx) : 0.f; sum += p[kg]; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    if (drop.thresh16 != 0 && (i & 7) == 0) {
#pragma unroll
      for (int kg = 0; kg < KG; ++kg) rnd[kg] = attn_rand(drop, (unsigned long long)rh, i >> 3, lane + 32 * kg);
    }
    float a[KG], dS[KG];
    float dot = 0.f;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      p[kg] *= inv;
      float keep_scale = 1.f;
      if (drop.thresh16 != 0) keep_scale = attn_keep(drop, rnd[kg], i) ? drop.scale : 0.f;
      a[kg] = p[kg] * keep_scale;       // dropped probabilities (what multiplied V in the forward)
      dA[kg] = dA[kg] * keep_scale;     // d/dp through the dropout
      dot += dA[kg] * p[kg];
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) dS[kg] = p[kg] * (dA[kg] - dot) * scale;
    const float qi0 = q[i * ROWP + lane], qi1 = q[i * ROWP + lane + 32];
    const float do0 = dO[i * ROWP + lane], do1 = dO[i * ROWP + lane + 32];
    float dq0 = 0.f, dq1 = 0.f;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      const int nj = min(32, L - 32 * kg);
      for (int jj = 0; jj < nj; ++jj) {
        const float ds = __shfl_sync(0xffffffffu, dS[kg], jj);
        const float aa = __shfl_sync(0xffffffffu, a[kg], jj);
        const int j = jj + 32 * kg;
        dq0 += ds * k[j * ROWP + lane];
        dq1 += ds * k[j * ROWP + lane + 32];
        dk[j * ROWP + lane] += ds * qi0;
        dk[j * ROWP + lane + 32] += ds * qi1;
        dv[j * ROWP + lane] += aa * do0;
        dv[j * ROWP + lane + 32] += aa * do1;
      }
    }
    __syncwarp();  // all lanes done reading dO row i
    dO[i * ROWP + lane] = dq0;
    dO[i * ROWP + lane + 32] = dq1;
  }
  __syncwarp();
  for (int idx = lane; idx < L * 8; idx += 32) {
    const int i = idx >> 3, part = idx & 7;
    const size_t base = ((size_t)r * L + i) * 3 * D + h * DH + part * 8;
    float t[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = dO[i * ROWP + part * 8 + e];
    store8(dqkv, base, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = dk[i * ROWP + part * 8 + e];
    store8(dqkv, base + D, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = dv[i * ROWP + part * 8 + e];
    store8(dqkv, base + 2 * D, t);
  }
}

static inline CBfPtr cbf(const clipdlm_bf_t* p) {
  CBfPtr r; r.hi = (const __nv_bfloat16*)p->hi; r.lo = (const __nv_bfloat16*)p->lo; return r;
}
static inline BfPtr mbf(const clipdlm_bf_t* p) {
  BfPtr r; r.hi = (__nv_bfloat16*)p->hi; r.lo = (__nv_bfloat16*)p->lo; return r;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  CLIPDLM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

static int warps_for(size_t per_warp_bytes) {
  int w = (int)((200 * 1024) / per_warp_bytes);
  if (w > 4) w = 4;
  return w;
}

int attn_fwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, int R, int L, int D, int H, const clipdlm_bf_t* ctx,
                      unsigned long long seed, uint32_t site, float p, cudaStream_t st) {
  CLIPDLM_CHECK(qkv && qkv->hi && ctx && ctx->hi && keymask, "attn_fwd: null pointer");
  CLIPDLM_CHECK(H > 0 && D == H * DH, "attn_fwd: head dim must be 64 (D %d, H %d)", D, H);
  CLIPDLM_CHECK(L >= 1 && L <= 128, "attn_fwd: L %d out of range (1..128)", L);
  const size_t per_warp = (size_t)3 * L * ROWP * sizeof(float);
  const int W = warps_for(per_warp);
  CLIPDLM_CHECK(W >= 1, "attn_fwd: L %d needs too much shared memory", L);
  const size_t smem = per_warp * W;
  const long long items = (long long)R * H;
  const int grid = (int)((items + W - 1) / W);
  const DropoutCfg d = make_drop(seed, site, p);
  const float scale = 0.125f;  // 1/sqrt(64)
  const int kg = (L + 31) / 32;
#define LAUNCH_FWD(KG)                                                                            \
  {                                                                                               \
    if (set_smem(attn_fwd_kernel<KG>, smem)) return -1;                                           \
    attn_fwd_kernel<KG><<<grid, W * 32, smem, st>>>(cbf(qkv), keymask, R, L, D, H, mbf(ctx), d, scale, W); \
  }
  if (kg == 1) LAUNCH_FWD(1) else if (kg == 2) LAUNCH_FWD(2) else if (kg == 3) LAUNCH_FWD(3) else LAUNCH_FWD(4)
#undef LAUNCH_FWD
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int attn_bwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int R, int L, int D, int H,
                      const clipdlm_bf_t* dqkv, unsigned long long seed, uint32_t site, float p, cudaStream_t st) {
  CLIPDLM_CHECK(qkv && qkv->hi && dctx && dctx->hi && dqkv && dqkv->hi && keymask, "attn_bwd: null pointer");
  CLIPDLM_CHECK(H > 0 && D == H * DH, "attn_bwd: head dim must be 64 (D %d, H %d)", D, H);
  CLIPDLM_CHECK(L >= 1 && L <= 128, "attn_bwd: L %d out of range (1..128)", L);
  const size_t per_warp = (size_t)6 * L * ROWP * sizeof(float);
  const int W = warps_for(per_warp);
  CLIPDLM_CHECK(W >= 1, "attn_bwd: L %d needs too much shared memory", L);
  const size_t smem = per_warp * W;
  const long long items = (long long)R * H;
  const int grid = (int)((items + W - 1) / W);
  const DropoutCfg d = make_drop(seed, site, p);
  const float scale = 0.125f;
  const int kg = (L + 31) / 32;
#define LAUNCH_BWD(KG)                                                                            \
  {                                                                                               \
    if (set_smem(attn_bwd_kernel<KG>, smem)) return -1;                                           \
    attn_bwd_kernel<KG><<<grid, W * 32, smem, st>>>(cbf(qkv), keymask, cbf(dctx), R, L, D, H, mbf(dqkv), d, scale, W); \
  }
  if (kg == 1) LAUNCH_BWD(1) else if (kg == 2) LAUNCH_BWD(2) else if (kg == 3) LAUNCH_BWD(3) else LAUNCH_BWD(4)
#undef LAUNCH_BWD
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace clipdlm
