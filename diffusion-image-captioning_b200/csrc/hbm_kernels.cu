// HBM-bound kernels of the clipdlm hot path (sm_100a): everything that is not a GEMM or attention.
//
//   embed_fwd / embed_bwd      K1-K4: embedding gather + q_sample + CLIP concat/add fusion + segment/position add + LayerNorm
//   layernorm_fwd / _bwd       post-LN blocks, MLM-transform LN; bwd fuses dropout masks, gelu', and the bias column sums
//   colsum                     bias gradients of the wide Linears
//   embed_loss                 LOSS_FUNC (L1 / L2-norm variants) forward value + gradient in one pass
//   small_linear fwd/bwd       image_linear / text_linear (512 -> D), evaluated once per caption instead of once per row
//   adamw                      flat multi-tensor AdamW, refreshes the bf16 (pair) shadow and zeroes the gradient
//   to_bf16 / to_f32 / gather_rows_f32 / keymask
//
// Design rule for all of them: one warp owns one row of D (= NV * 256) elements, every lane moves 16-byte vectors
// (8 bf16) that are contiguous across the warp (512 B per request), reductions are warp shuffles, cross-row
// reductions (dw/db/bias grads) are register partials -> shared memory -> one fp32 atomic per column per block.
#include "common.cuh"
#include "../../include/clipdlm.h"

namespace clipdlm {

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

DropoutCfg make_drop(unsigned long long seed, uint32_t site, float p) {
  DropoutCfg d;
  d.seed = seed; d.site = site;
  d.thresh16 = p > 0.f ? (uint32_t)(p * 65536.f + 0.5f) : 0u;
  d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  return d;
}

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// Coherent variant for TRAINABLE parameters read by kernels launched with programmatic dependent launch (the fast LayerNorm kernels): __ldg /
// ld.global.nc is outside the acquire griddepcontrol.wait performs, and a parameter is rewritten by the optimizer between two launches of the
// same kernel - the GEMM epilogues' bias served stale values that way. Plain ld.global, immune to const / __restrict__ inference.
__device__ __forceinline__ void load8_f32_coherent(const float* p, float (&v)[8]) {
  float4 a, b;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void apply_drop8(const DropoutCfg& d, unsigned long long elem_idx, float (&v)[8]) {
  const uint32_t keep = dropout_keep8(d, elem_idx >> 3);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ((keep >> i) & 1u) ? v[i] * d.scale : 0.f;
}

// mean / rstd of a row held as x[NV][8] per lane (two-pass, fp32)
template <int NV>
__device__ __forceinline__ void row_stats(const float (&x)[NV][8], int D, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[v][i];
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = x[v][i] - mean; q += d * d; }
  rstd = rsqrtf(warp_sum(q) / (float)D + eps);
}

// ------------------------------------------------------------------------------------------------------------------
// embed_fwd
// ------------------------------------------------------------------------------------------------------------------
struct EmbedArgs {
  int R, B, Ltxt, L, D, fusion, mode, guided;
  const float* x_in; long long x_in_stride;
  const float* emb_table; const int* ids; const float* noise; const float* coef_a; const float* coef_b;
  const float* img_proj; const float* txt_proj; const float* seg; const float* pos;
  const float* ln_w; const float* ln_b; float ln_eps;
  BfPtr z, h;
  DropoutCfg drop;
};

template <int NV>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const EmbedArgs e) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)e.R * e.L) return;
  const int r = (int)(row / e.L), p = (int)(row % e.L);
  const int s = r / e.B, b = r % e.B;
  float x[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * 8;
    float t[8];
    if (p < e.Ltxt) {
      if (e.mode == 0) {
        load8_f32(e.x_in + (long long)r * e.x_in_stride + (long long)p * e.D + c, t);
      } else {
        const int id = e.ids[b * e.Ltxt + p];
        float ev[8], nv[8];
        load8_f32(e.emb_table + (long long)id * e.D + c, ev);
        load8_f32(e.noise + ((long long)b * e.Ltxt + p) * e.D + c, nv);
        const float ca = e.coef_a[s], cb = e.coef_b[s];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = ca * ev[i] + cb * nv[i];
      }
      if (e.fusion == 1) {
        float q[8];
        load8_f32(e.img_proj + (long long)b * e.D + c, q);
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] += q[i];
        if (e.guided) {
          load8_f32(e.txt_proj + (long long)b * e.D + c, q);
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] += q[i];
        }
      }
    } else {
      load8_f32((p == e.Ltxt ? e.img_proj : e.txt_proj) + (long long)b * e.D + c, t);
    }
    float a[8];
    if (e.fusion == 0) {
      load8_f32(e.seg + (p >= e.Ltxt ? e.D : 0) + c, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] += a[i];
    }
    load8_f32(e.pos + (long long)p * e.D + c, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[v][i] = t[i] + a[i];
  }
  if (e.z.hi != nullptr) {
#pragma unroll
    for (int v = 0; v < NV; ++v) store8(e.z, (size_t)row * e.D + (v * 32 + lane) * 8, x[v]);
  }
  float mean, rstd;
  row_stats<NV>(x, e.D, e.ln_eps, mean, rstd);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * 8;
    float w[8], bb[8], y[8];
    load8_f32(e.ln_w + c, w);
    load8_f32(e.ln_b + c, bb);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (x[v][i] - mean) * rstd * w[i] + bb[i];
    if (e.drop.thresh16 != 0) apply_drop8(e.drop, (unsigned long long)row * e.D + c, y);
    store8(e.h, (size_t)row * e.D + c, y);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// embed_bwd: (1) position / segment gradients, (2) CLIP projection gradients
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_bwd_posseg_kernel(CBfPtr dz, int R, int L, int Ltxt, int D, int fusion, int rows_per_slab,
                                                               float* d_pos, float* d_seg) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int p = blockIdx.y;
  const int r0 = blockIdx.z * rows_per_slab;
  const int r1 = min(R, r0 + rows_per_slab);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = r0 + warp; r < r1; r += 8) {
    float v[8];
    load8(dz, ((size_t)r * L + p) * D + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int t = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][t];
  const int col = blockIdx.x * 256 + t;
  atomicAdd(d_pos + (size_t)p * D + col, s);
  if (fusion == 0) atomicAdd(d_seg + (p >= Ltxt ? D : 0) + col, s);
}

// thread per (b, 8-column vector): sums over the samples s (rows r = s*B + b)
__global__ void embed_bwd_proj_kernel(CBfPtr dz, int R, int B, int L, int Ltxt, int D, int fusion, int guided, float* d_img, float* d_txt) {
  const int nvec = D / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * nvec) return;
  const int b = (int)(idx / nvec), c = (int)(idx % nvec) * 8;
  const int S = R / B;
  float ai[8] = {0, 0, 0, 0, 0, 0, 0, 0}, at[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = 0; s < S; ++s) {
    const size_t r = (size_t)s * B + b;
    float v[8];
    if (fusion == 0) {
      load8(dz, (r * L + Ltxt) * D + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) ai[i] += v[i];
      load8(dz, (r * L + Ltxt + 1) * D + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) at[i] += v[i];
    } else {
      for (int p = 0; p < L; ++p) {
        load8(dz, (r * L + p) * D + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) ai[i] += v[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    d_img[(size_t)b * D + c + i] += ai[i];
    if (fusion == 0) d_txt[(size_t)b * D + c + i] += at[i];
    else if (guided) d_txt[(size_t)b * D + c + i] += ai[i];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm forward
// ------------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(CBfPtr z, const float* __restrict__ w, const float* __restrict__ b, float eps,
                                                            long long rows, int D, BfPtr y, float* y_f32, DropoutCfg drop) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float x[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) load8(z, (size_t)row * D + (v * 32 + lane) * 8, x[v]);
  float mean, rstd;
  row_stats<NV>(x, D, eps, mean, rstd);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * 8;
    float ww[8], bb[8], o[8];
    load8_f32(w + c, ww);
    load8_f32(b + c, bb);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (x[v][i] - mean) * rstd * ww[i] + bb[i];
    if (drop.thresh16 != 0) apply_drop8(drop, (unsigned long long)row * D + c, o);
    if (y.hi != nullptr) store8(y, (size_t)row * D + c, o);
    if (y_f32 != nullptr) store8_f32(y_f32 + (size_t)row * D + c, o);
  }
}

// Plain-bf16, D = 768 specialisation of the forward: persistent warps, rows streamed through a per-warp bulk-copy ring, packed
// fp32 math, w / b held in registers.  (The generic kernel: one row per warp, 319 instructions per row, latency-exposed loads.)
constexpr int LNF_STAGES = 4;
template <bool F32OUT, int NV = 3>
__global__ void __launch_bounds__(256, 2) layernorm_fwd_fast_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ w,
                                                                    const float* __restrict__ b, float eps, long long rows,
                                                                    __nv_bfloat16* __restrict__ y, float* __restrict__ y_f32) {
  constexpr int D = NV * 256, NW = 8;   // NV = 3: D = 768 (the reference model); NV = 4: D = 1024 (bert-large geometry)
  constexpr uint32_t row_bytes = D * 2;
  extern __shared__ __align__(16) uint8_t lnf_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* ring = lnf_smem + (size_t)warp * LNF_STAGES * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lnf_smem + (size_t)NW * LNF_STAGES * row_bytes) + warp * LNF_STAGES;
  const long long row0 = (long long)blockIdx.x * NW + warp, row_step = (long long)gridDim.x * NW;
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (lane == 0) {
    for (int s = 0; s < LNF_STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  pdl_wait();
  if (lane == 0) {
    for (int s = 0; s < LNF_STAGES; ++s)
      if (row0 + s * row_step < rows) {
        mbar_arrive_expect_tx(&bars[s], row_bytes);
        bulk_load_1d(ring + (size_t)s * row_bytes, z + (size_t)(row0 + s * row_step) * D, row_bytes, &bars[s]);
      }
  }
  __syncwarp();
  f32x2 W[NV][4], B[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float wv[8], bv[8];
    load8_f32_coherent(w + (v * 32 + lane) * 8, wv);
    load8_f32_coherent(b + (v * 32 + lane) * 8, bv);
#pragma unroll
    for (int i = 0; i < 4; ++i) { W[v][i] = pk2(wv[2 * i], wv[2 * i + 1]); B[v][i] = pk2(bv[2 * i], bv[2 * i + 1]); }
  }
  constexpr float inv_d = 1.0f / (float)D;
  int stage = 0; uint32_t phase = 0;
  for (long long row = row0; row < rows; row += row_step) {
    f32x2 X[NV][4];
    mbar_wait(&bars[stage], phase);
    {
      const uint8_t* sz = ring + (size_t)stage * row_bytes;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const uint4 xr = *reinterpret_cast<const uint4*>(sz + (v * 32 + lane) * 16);
        X[v][0] = bf2_to_f2(xr.x); X[v][1] = bf2_to_f2(xr.y); X[v][2] = bf2_to_f2(xr.z); X[v][3] = bf2_to_f2(xr.w);
      }
    }
    __syncwarp();
    if (row + LNF_STAGES * row_step < rows && elect_one()) {   // (elect.sync, not lane == 0: no ELECT / R2UR waterfall around the bulk copy)
      mbar_arrive_expect_tx(&bars[stage], row_bytes);
      bulk_load_1d(ring + (size_t)stage * row_bytes, z + (size_t)(row + LNF_STAGES * row_step) * D, row_bytes, &bars[stage]);
    }
    if (++stage == LNF_STAGES) { stage = 0; phase ^= 1; }
    f32x2 acc = 0ull;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc = add2(acc, X[v][i]);
    const float mean = warp_sum(hsum2(acc)) * inv_d;
    const f32x2 nmean = pk2(-mean, -mean);
    acc = 0ull;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) { X[v][i] = add2(X[v][i], nmean); acc = fma2(X[v][i], X[v][i], acc); }
    const float rstd = rsqrtf(warp_sum(hsum2(acc)) * inv_d + eps);
    const f32x2 rstd2 = pk2(rstd, rstd);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const size_t off = (size_t)row * D + (v * 32 + lane) * 8;
      f32x2 o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fma2(mul2(X[v][i], rstd2), W[v][i], B[v][i]);   // same rounding order as the generic kernel
      if (y != nullptr) *reinterpret_cast<uint4*>(y + off) = make_uint4(f2_to_bf2(o[0]), f2_to_bf2(o[1]), f2_to_bf2(o[2]), f2_to_bf2(o[3]));
      if (F32OUT) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) upk2(o[i], f[2 * i], f[2 * i + 1]);
        store8_f32(y_f32 + off, f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm backward (+ fused dropout masks, gelu', bias column sums)
// ------------------------------------------------------------------------------------------------------------------
struct LnBwdArgs {
  CBfPtr z, dy;
  const float* w; float eps; long long rows; int D;
  BfPtr dz;
  float* dw; float* db;
  DropoutCfg drop_out;   // dropout that followed this LN's output (mask applied to dy)
  BfPtr dz_drop;         // optional second output dz * mask_in / (1 - p_in)
  DropoutCfg drop_in;
  CBfPtr gelu_u;         // optional: dz *= gelu'(u)   (MLM transform head: z = gelu(u))
  float* dbias;          // optional: += column sums of (dz_drop if set else dz)   -> bias grad of the producing Linear
};

// Rows are streamed through a per-warp ring of shared-memory stages filled by 1-D bulk copies (cp.async.bulk, one elected lane,
// completion on an mbarrier): the 72 column-partial accumulators per lane leave room for only 12 warps per SM, far too few to
// cover HBM latency with register loads (measured 2.1 TB/s); with LNB_STAGES rows in flight per warp the kernel keeps
// > 100 KB per SM outstanding at zero register cost.
constexpr int LNB_STAGES = 3;
template <int NV>
__global__ void __launch_bounds__(128, 3) layernorm_bwd_kernel(const LnBwdArgs a) {
  extern __shared__ __align__(16) uint8_t lnb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const bool pair = a.z.lo != nullptr;
  const int D = a.D;
  const int narr = pair ? 4 : 2;                       // z.hi, dy.hi (, z.lo, dy.lo)
  const uint32_t row_bytes = (uint32_t)D * 2;
  float* red = reinterpret_cast<float*>(lnb_smem);      // [nwarps][D] reused for dw, db, dbias
  uint8_t* ring = lnb_smem + (size_t)nwarps * D * sizeof(float) + (size_t)warp * LNB_STAGES * narr * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lnb_smem + (size_t)nwarps * D * sizeof(float) + (size_t)nwarps * LNB_STAGES * narr * row_bytes) +
                   warp * LNB_STAGES;
  const long long row0 = (long long)blockIdx.x * nwarps + warp, row_step = (long long)gridDim.x * nwarps;
  if (lane == 0) {
    for (int s = 0; s < LNB_STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto issue = [&](int s, long long row) {   // lane 0 only
    uint8_t* dst = ring + (size_t)s * narr * row_bytes;
    mbar_arrive_expect_tx(&bars[s], narr * row_bytes);
    bulk_load_1d(dst, a.z.hi + (size_t)row * D, row_bytes, &bars[s]);
    bulk_load_1d(dst + row_bytes, a.dy.hi + (size_t)row * D, row_bytes, &bars[s]);
    if (pair) {
      bulk_load_1d(dst + 2 * row_bytes, a.z.lo + (size_t)row * D, row_bytes, &bars[s]);
      bulk_load_1d(dst + 3 * row_bytes, a.dy.lo + (size_t)row * D, row_bytes, &bars[s]);
    }
  };
  if (lane == 0) {
    for (int s = 0; s < LNB_STAGES; ++s)
      if (row0 + s * row_step < a.rows) issue(s, row0 + s * row_step);
  }
  float pdw[NV][8], pdb[NV][8], pbias[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) { pdw[v][i] = 0.f; pdb[v][i] = 0.f; pbias[v][i] = 0.f; }
  int stage = 0; uint32_t phase = 0;
  for (long long row = row0; row < a.rows; row += row_step) {
    float x[NV][8], g[NV][8];
    mbar_wait(&bars[stage], phase);
    {
      const __nv_bfloat16* sz = reinterpret_cast<const __nv_bfloat16*>(ring + (size_t)stage * narr * row_bytes);
      const __nv_bfloat16* sdy = sz + D;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = (v * 32 + lane) * 8;
        lds8(sz + c, pair ? sz + 2 * D + c : nullptr, x[v]);
        lds8(sdy + c, pair ? sdy + 2 * D + c : nullptr, g[v]);
      }
    }
    __syncwarp();   // every lane has consumed the stage: refill it with the row LNB_STAGES ahead
    if (row + LNB_STAGES * row_step < a.rows && elect_one()) issue(stage, row + LNB_STAGES * row_step);
    if (++stage == LNB_STAGES) { stage = 0; phase ^= 1; }
    if (a.drop_out.thresh16 != 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v) apply_drop8(a.drop_out, (unsigned long long)row * D + (v * 32 + lane) * 8, g[v]);
    }
    float mean, rstd;
    row_stats<NV>(x, a.D, a.eps, mean, rstd);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float wv[8];  // re-read per row (L1-resident): keeping w in registers costs occupancy on this HBM-bound kernel
      load8_f32(a.w + (v * 32 + lane) * 8, wv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (x[v][i] - mean) * rstd;
        const float gw = g[v][i] * wv[i];
        pdw[v][i] += g[v][i] * xh;
        pdb[v][i] += g[v][i];
        s1 += gw; s2 += gw * xh;
        x[v][i] = xh; g[v][i] = gw;
      }
    }
    s1 = warp_sum(s1) / (float)a.D;
    s2 = warp_sum(s2) / (float)a.D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const size_t off = (size_t)row * a.D + (v * 32 + lane) * 8;
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = (g[v][i] - s1 - x[v][i] * s2) * rstd;
      if (a.gelu_u.hi != nullptr) {
        float u[8];
        load8(a.gelu_u, off, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] *= dgelu_f(u[i]);
      }
      store8(a.dz, off, d);
      if (a.dz_drop.hi != nullptr) {
        apply_drop8(a.drop_in, (unsigned long long)off, d);
        store8(a.dz_drop, off, d);
      }
      if (a.dbias != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pbias[v][i] += d[i];
      }
    }
  }
  // block reduction: three passes through the same [nwarps][D] shared buffer
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? a.dw : (which == 1 ? a.db : a.dbias);
    if (dst == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        red[warp * a.D + (v * 32 + lane) * 8 + i] = which == 0 ? pdw[v][i] : (which == 1 ? pdb[v][i] : pbias[v][i]);
    __syncthreads();
    for (int c = threadIdx.x; c < a.D; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += red[w * a.D + c];
      atomicAdd(dst + c, s);
    }
  }
}

// Plain-bf16 specialisation of the kernel above: same row ring, the per-element arithmetic in packed fp32 pairs (FFMA2 / FADD2 /
// FMUL2), the optional features as template flags (no per-row branches), reciprocal-multiplies instead of divisions.  The generic
// kernel spends 742 warp instructions per 768-wide row (357 of them scalar fp32 math) and is issue-bound at 12 warps per SM.
// NV = 3 (D = 768): 168 registers, 3 CTAs per SM; NV = 4 (D = 1024, the bert-large geometry): 32 more packed accumulators, 2 CTAs per SM.
template <int NV, bool DROP_OUT, bool DZ_DROP, bool GELU, bool DBIAS>
__global__ void __launch_bounds__(128, NV <= 3 ? 3 : 2) layernorm_bwd_fast_kernel(const LnBwdArgs a) {
  extern __shared__ __align__(16) uint8_t lnb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) pdl_launch_dependents();
  constexpr int nwarps = 4;
  const int D = a.D;
  const uint32_t row_bytes = (uint32_t)D * 2;
  float* red = reinterpret_cast<float*>(lnb_smem);
  uint8_t* ring = lnb_smem + (size_t)nwarps * D * sizeof(float) + (size_t)warp * LNB_STAGES * 2 * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lnb_smem + (size_t)nwarps * D * sizeof(float) + (size_t)nwarps * LNB_STAGES * 2 * row_bytes) +
                   warp * LNB_STAGES;
  const long long row0 = (long long)blockIdx.x * nwarps + warp, row_step = (long long)gridDim.x * nwarps;
  if (lane == 0) {
    for (int s = 0; s < LNB_STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto issue = [&](int s, long long row) {   // lane 0 only
    uint8_t* dst = ring + (size_t)s * 2 * row_bytes;
    mbar_arrive_expect_tx(&bars[s], 2 * row_bytes);
    bulk_load_1d(dst, a.z.hi + (size_t)row * D, row_bytes, &bars[s]);
    bulk_load_1d(dst + row_bytes, a.dy.hi + (size_t)row * D, row_bytes, &bars[s]);
  };
  pdl_wait();
  if (lane == 0) {
    for (int s = 0; s < LNB_STAGES; ++s)
      if (row0 + s * row_step < a.rows) issue(s, row0 + s * row_step);
  }
  // LayerNorm weight of this lane's columns, packed (the same 24 columns for every row)
  f32x2 W[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float wv[8];
    load8_f32_coherent(a.w + (v * 32 + lane) * 8, wv);
#pragma unroll
    for (int i = 0; i < 4; ++i) W[v][i] = pk2(wv[2 * i], wv[2 * i + 1]);
  }
  f32x2 pdw[NV][4], pdb[NV][4], pbias[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i) { pdw[v][i] = 0ull; pdb[v][i] = 0ull; pbias[v][i] = 0ull; }
  const float inv_d = 1.0f / (float)D;
  int stage = 0; uint32_t phase = 0;
  for (long long row = row0; row < a.rows; row += row_step) {
    f32x2 X[NV][4], G[NV][4];
    mbar_wait(&bars[stage], phase);
    {
      const uint8_t* sz = ring + (size_t)stage * 2 * row_bytes;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const uint4 xr = *reinterpret_cast<const uint4*>(sz + (v * 32 + lane) * 16);
        const uint4 gr = *reinterpret_cast<const uint4*>(sz + row_bytes + (v * 32 + lane) * 16);
        X[v][0] = bf2_to_f2(xr.x); X[v][1] = bf2_to_f2(xr.y); X[v][2] = bf2_to_f2(xr.z); X[v][3] = bf2_to_f2(xr.w);
        G[v][0] = bf2_to_f2(gr.x); G[v][1] = bf2_to_f2(gr.y); G[v][2] = bf2_to_f2(gr.z); G[v][3] = bf2_to_f2(gr.w);
      }
    }
    __syncwarp();
    if (row + LNB_STAGES * row_step < a.rows && elect_one()) issue(stage, row + LNB_STAGES * row_step);
    if (++stage == LNB_STAGES) { stage = 0; phase ^= 1; }
    if (DROP_OUT) {   // the dropout that followed this LN's output masks dy
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const uint32_t keep = dropout_keep8(a.drop_out, ((unsigned long long)row * D + (v * 32 + lane) * 8) >> 3);
        const float sc = a.drop_out.scale;
#pragma unroll
        for (int i = 0; i < 4; ++i) G[v][i] = mul2(G[v][i], pk2(((keep >> (2 * i)) & 1u) ? sc : 0.f, ((keep >> (2 * i + 1)) & 1u) ? sc : 0.f));
      }
    }
    // two-pass statistics (as the forward): mean, then centred second moment; X becomes x - mean
    f32x2 acc = 0ull;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc = add2(acc, X[v][i]);
    const float mean = warp_sum(hsum2(acc)) * inv_d;
    const f32x2 nmean = pk2(-mean, -mean);
    acc = 0ull;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) { X[v][i] = add2(X[v][i], nmean); acc = fma2(X[v][i], X[v][i], acc); }
    const float rstd = rsqrtf(warp_sum(hsum2(acc)) * inv_d + a.eps);
    const f32x2 rstd2 = pk2(rstd, rstd);
    f32x2 s1 = 0ull, s2 = 0ull;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2 xh = mul2(X[v][i], rstd2);
        const f32x2 gw = mul2(G[v][i], W[v][i]);
        pdw[v][i] = fma2(G[v][i], xh, pdw[v][i]);
        pdb[v][i] = add2(pdb[v][i], G[v][i]);
        s1 = add2(s1, gw);
        s2 = fma2(gw, xh, s2);
        X[v][i] = xh; G[v][i] = gw;
      }
    const float s1r = warp_sum(hsum2(s1)) * inv_d * rstd, s2r = warp_sum(hsum2(s2)) * inv_d * rstd;
    const f32x2 ns1 = pk2(-s1r, -s1r), ns2 = pk2(-s2r, -s2r);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const size_t off = (size_t)row * D + (v * 32 + lane) * 8;
      f32x2 d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = fma2(X[v][i], ns2, fma2(G[v][i], rstd2, ns1));   // (gw - s1 - xh s2) rstd
      if (GELU) {
        const uint4 ur = *reinterpret_cast<const uint4*>(a.gelu_u.hi + off);
        const uint32_t uw[4] = {ur.x, ur.y, ur.z, ur.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 u = unpack_bf16x2(uw[i]);
          d[i] = mul2(d[i], pk2(dgelu_f(u.x), dgelu_f(u.y)));
        }
      }
      *reinterpret_cast<uint4*>(a.dz.hi + off) = make_uint4(f2_to_bf2(d[0]), f2_to_bf2(d[1]), f2_to_bf2(d[2]), f2_to_bf2(d[3]));
      if (DZ_DROP) {
        const uint32_t keep = dropout_keep8(a.drop_in, (unsigned long long)off >> 3);
        const float sc = a.drop_in.scale;
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = mul2(d[i], pk2(((keep >> (2 * i)) & 1u) ? sc : 0.f, ((keep >> (2 * i + 1)) & 1u) ? sc : 0.f));
        *reinterpret_cast<uint4*>(a.dz_drop.hi + off) = make_uint4(f2_to_bf2(d[0]), f2_to_bf2(d[1]), f2_to_bf2(d[2]), f2_to_bf2(d[3]));
      }
      if (DBIAS) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pbias[v][i] = add2(pbias[v][i], d[i]);
      }
    }
  }
  // block reduction: three passes through the same [nwarps][D] shared buffer
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? a.dw : (which == 1 ? a.db : (DBIAS ? a.dbias : nullptr));
    if (dst == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        upk2(which == 0 ? pdw[v][i] : (which == 1 ? pdb[v][i] : pbias[v][i]), lo, hi);
        *reinterpret_cast<float2*>(&red[warp * D + (v * 32 + lane) * 8 + 2 * i]) = make_float2(lo, hi);
      }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += red[w * D + c];
      atomicAdd(dst + c, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// colsum
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(CBfPtr x, long long rows, int N, long long rows_per_slab, float* out) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  const long long r1 = min(rows, r0 + rows_per_slab);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = r0 + warp; r < r1; r += 8) {
    float v[8];
    load8(x, (size_t)r * N + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
  atomicAdd(out + blockIdx.x * 256 + threadIdx.x, s);
}

// ------------------------------------------------------------------------------------------------------------------
// embed_loss: one block per sequence row r
// ------------------------------------------------------------------------------------------------------------------
// target row of (r, p): tgt != NULL ? tgt[(r % tgt_rows), p, :] : emb[ids[r % B, p], :]
__device__ __forceinline__ const float* loss_target_row(const float* emb, const int* ids, const float* tgt, int tgt_rows, int r, int b, int p,
                                                        int Ltxt, int D) {
  if (tgt != nullptr) return tgt + ((size_t)(r % tgt_rows) * Ltxt + p) * D;
  return emb + (size_t)ids[b * Ltxt + p] * D;
}
__global__ void __launch_bounds__(256) embed_loss_kernel(CBfPtr x_out, const float* __restrict__ emb, const int* __restrict__ ids,
                                                         const float* __restrict__ tgt, int tgt_rows, int R,
                                                         int B, int Ltxt, int L, int D, int kind, double inv_div, float gscale,
                                                         double* loss_acc, BfPtr dx) {
  __shared__ float wred[8];
  __shared__ float total_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x;
  const int b = r % B;
  const int nvec = D / 8;
  float acc = 0.f;
  for (int p = warp; p < Ltxt; p += 8) {
    const size_t row = (size_t)r * L + p;
    const float* e = loss_target_row(emb, ids, tgt, tgt_rows, r, b, p, Ltxt, D);
    for (int v = lane; v < nvec; v += 32) {
      float x[8], t[8];
      load8(x_out, row * D + v * 8, x);
      load8_f32(e + v * 8, t);
      if (kind <= 1) {
        float g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = x[i] - t[i];
          acc += fabsf(d);
          g[i] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
        }
        if (dx.hi != nullptr) store8(dx, row * D + v * 8, g);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = x[i] - t[i]; acc += d * d; }
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) wred[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += wred[i];
    total_s = t;
  }
  __syncthreads();
  const float total = total_s;
  if (kind >= 2) {
    const float norm = sqrtf(total);
    const float gs = norm > 0.f ? gscale / norm : 0.f;
    if (dx.hi != nullptr) {
      for (int p = warp; p < Ltxt; p += 8) {
        const size_t row = (size_t)r * L + p;
        const float* e = loss_target_row(emb, ids, tgt, tgt_rows, r, b, p, Ltxt, D);
        for (int v = lane; v < nvec; v += 32) {
          float x[8], t[8], g[8];
          load8(x_out, row * D + v * 8, x);
          load8_f32(e + v * 8, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = (x[i] - t[i]) * gs;
          store8(dx, row * D + v * 8, g);
        }
      }
    }
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(loss_acc, (double)norm * inv_div);
  } else {
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(loss_acc, (double)total * inv_div);
  }
  if (dx.hi != nullptr) {  // rows that do not enter the loss (CLIP positions) get a zero gradient
    const float zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = Ltxt + warp; p < L; p += 8)
      for (int v = lane; v < nvec; v += 32) store8(dx, ((size_t)r * L + p) * D + v * 8, zero);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// small linear (CLIP projections), fp32
// ------------------------------------------------------------------------------------------------------------------
// Shared-memory tiled fp32 SIMT GEMM, 64 x 64 output tile per block of 256 threads (4 x 4 per thread), k-step 16.
//   KCONTIG = true : C[m, n] = sum_k A[m, k] Bm[n, k] + bias[n]          (A [M][K], Bm [N][K]; forward  y = x W^T + b)
//   KCONTIG = false: C[m, n] += sum_k A[k, m] Bm[k, n]; csum[m] += sum_k A[k, m]   (A [K][M], Bm [K][N]; backward dW += dy^T x, db += colsum dy)
// fp32 on purpose: the CLIP projections feed the parity-mode (fp32-class) path and are 0.4 GFLOP per call.
template <bool KCONTIG>
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ bias,
                                                         int M, int N, int K, float* __restrict__ C, float* __restrict__ csum,
                                                         int k_per_split) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  float acc[4][4] = {};
  float asum[4] = {0.f, 0.f, 0.f, 0.f};
  // KCONTIG = false only: gridDim.z > 1 splits the reduction (rows of a long token dimension) across blocks, partial sums are
  // combined with fp32 atomics (the destination is a gradient accumulator anyway)
  const int k_begin = KCONTIG ? 0 : blockIdx.z * k_per_split;
  const int k_end = KCONTIG ? K : min(K, k_begin + k_per_split);
  const bool split = !KCONTIG && gridDim.z > 1;
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    if (KCONTIG) {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (m0 + r < M && k0 + kq < K) a = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * K + k0 + kq));
      if (n0 + r < N && k0 + kq < K) b = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)(n0 + r) * K + k0 + kq));
      As[kq][r] = a.x; As[kq + 1][r] = a.y; As[kq + 2][r] = a.z; As[kq + 3][r] = a.w;
      Bs[kq][r] = b.x; Bs[kq + 1][r] = b.y; Bs[kq + 2][r] = b.z; Bs[kq + 3][r] = b.w;
    } else {
      const int kk = tid >> 4, q = (tid & 15) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (k0 + kk < k_end && m0 + q < M) a = __ldg(reinterpret_cast<const float4*>(A + (size_t)(k0 + kk) * M + m0 + q));
      if (k0 + kk < k_end && n0 + q < N) b = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)(k0 + kk) * N + n0 + q));
      *reinterpret_cast<float4*>(&As[kk][q]) = a;
      *reinterpret_cast<float4*>(&Bs[kk][q]) = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!KCONTIG) asum[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (KCONTIG) C[(size_t)m * N + n] = acc[i][j] + (bias != nullptr ? bias[n] : 0.f);
      else if (split) atomicAdd(C + (size_t)m * N + n, acc[i][j]);
      else C[(size_t)m * N + n] += acc[i][j];
    }
    if (!KCONTIG && csum != nullptr && blockIdx.y == 0 && tx == 0) {
      if (split) atomicAdd(csum + m, asum[i]);
      else csum[m] += asum[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW single-tensor semantics applied to the flat buffer)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    __nv_bfloat16* __restrict__ sh_hi, __nv_bfloat16* __restrict__ sh_lo, long long n4,
                                                    float decay, float beta1, float beta2, float eps, float step_size, float inv_sqrt_bc2,
                                                    float grad_scale, int zero_grad) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = reinterpret_cast<float*>(&pp); float* G = reinterpret_cast<float*>(&gg);
    float* M = reinterpret_cast<float*>(&mm); float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j] * grad_scale;
      P[j] *= decay;
      M[j] = beta1 * M[j] + (1.f - beta1) * gr;
      V[j] = beta2 * V[j] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(V[j]) * inv_sqrt_bc2 + eps;
      P[j] -= step_size * (M[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sh_hi != nullptr) {
      uint2 h;
      h.x = pack_bf16x2(P[0], P[1]); h.y = pack_bf16x2(P[2], P[3]);
      reinterpret_cast<uint2*>(sh_hi)[i] = h;
      if (sh_lo != nullptr) {
        uint2 l;
        l.x = pack_bf16x2(P[0] - bf16_round(P[0]), P[1] - bf16_round(P[1]));
        l.y = pack_bf16x2(P[2] - bf16_round(P[2]), P[3] - bf16_round(P[3]));
        reinterpret_cast<uint2*>(sh_lo)[i] = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// conversions
// ------------------------------------------------------------------------------------------------------------------
__global__ void to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo != nullptr) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
__global__ void to_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = __bfloat162float(hi[i]);
    if (lo != nullptr) v += __bfloat162float(lo[i]);
    y[i] = v;
  }
}
__global__ void gather_rows_f32_kernel(CBfPtr x, long long rows_out, int len, int stride, int D, float* __restrict__ y) {
  const int nvec = D / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows_out * nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i % nvec) * 8;
    const long long src = (r / len) * stride + (r % len);
    float v[8];
    load8(x, (size_t)src * D + c, v);
    store8_f32(y + (size_t)r * D + c, v);
  }
}
// q_sample (diffuse_t, CLIP-DDPM.py:347-362): out[s, b, :] = ca[s] * x0[b, :] + cb[s] * noise[b, :]; n = elements per sample
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const float* __restrict__ ca,
                                const float* __restrict__ cb, long long n4, int S, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4 * S; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i / n4);
    const long long j = i % n4;
    const float4 x = __ldg(reinterpret_cast<const float4*>(x0) + j);
    const float4 e = __ldg(reinterpret_cast<const float4*>(noise) + j);
    const float a = ca[s], b = cb[s];
    reinterpret_cast<float4*>(out)[i] = make_float4(a * x.x + b * e.x, a * x.y + b * e.y, a * x.z + b * e.z, a * x.w + b * e.w);
  }
}
// keymask[r] (one word per 32 keys): text keys from attn_mask[b], image key visible, text-CLIP key visible iff guided
__global__ void keymask_kernel(const int* __restrict__ attn_mask, int R, int B, int Ltxt, int L, int fusion, int guided, int kw,
                               uint32_t* __restrict__ km) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * kw) return;
  const int r = idx / kw, w = idx % kw;
  const int b = r % B;
  uint32_t bits = 0;
  for (int j = w * 32; j < min(L, w * 32 + 32); ++j) {
    bool vis;
    if (j < Ltxt) vis = attn_mask == nullptr ? true : attn_mask[b * Ltxt + j] != 0;
    else if (fusion == 0) vis = (j == Ltxt) ? true : (guided != 0);
    else vis = true;
    if (vis) bits |= 1u << (j & 31);
  }
  km[idx] = bits;
}

// classifier-free guidance: mix of the guided / unguided encoder outputs, and per-row scaling of the upstream gradient
__global__ void __launch_bounds__(256) cfg_mix_kernel(BfPtr xu, CBfPtr xg, const int* __restrict__ guided, float w, int L, int D, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int vec_per_row = L * (D >> 3);
  const int r = (int)(i / vec_per_row);
  if (!guided[r]) return;
  float a[8], b[8];
  CBfPtr cu; cu.hi = xu.hi; cu.lo = xu.lo;
  load8(cu, (size_t)i * 8, a);
  load8(xg, (size_t)i * 8, b);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = (1.f + w) * b[k] - w * a[k];
  store8(xu, (size_t)i * 8, a);
}
__global__ void __launch_bounds__(256) row_scale_kernel(BfPtr g, const float* __restrict__ s_self, BfPtr ex, const float* __restrict__ s_ex,
                                                        int L, int D, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int vec_per_row = L * (D >> 3);
  const int r = (int)(i / vec_per_row);
  float a[8], b[8];
  CBfPtr cg; cg.hi = g.hi; cg.lo = g.lo;
  load8(cg, (size_t)i * 8, a);
  if (ex.hi != nullptr) {
    const float se = s_ex != nullptr ? s_ex[r] : 1.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = a[k] * se;
    store8(ex, (size_t)i * 8, b);
  }
  if (s_self != nullptr) {
    const float ss = s_self[r];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= ss;
    store8(g, (size_t)i * 8, a);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------------------------
static inline CBfPtr cbf(const clipdlm_bf_t* p) {
  CBfPtr r; r.hi = p ? (const __nv_bfloat16*)p->hi : nullptr; r.lo = p ? (const __nv_bfloat16*)p->lo : nullptr; return r;
}
static inline BfPtr mbf(const clipdlm_bf_t* p) {
  BfPtr r; r.hi = p ? (__nv_bfloat16*)p->hi : nullptr; r.lo = p ? (__nv_bfloat16*)p->lo : nullptr; return r;
}
static inline int grid_1d(long long n, int threads, int max_blocks) {
  long long b = (n + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}

#define DISPATCH_NV(D, ...)                                                                       \
  switch ((D) / 256) {                                                                            \
    case 1: { constexpr int NV = 1; __VA_ARGS__; break; }                                         \
    case 2: { constexpr int NV = 2; __VA_ARGS__; break; }                                         \
    case 3: { constexpr int NV = 3; __VA_ARGS__; break; }                                         \
    case 4: { constexpr int NV = 4; __VA_ARGS__; break; }                                         \
    default: CLIPDLM_CHECK(false, "unsupported model dim %d (need a multiple of 256, <= 1024)", (D)); \
  }

int cfg_mix_dispatch(const clipdlm_bf_t* xu, const clipdlm_bf_t* xg, const int* guided, float w, int R, int L, int D, cudaStream_t st) {
  CLIPDLM_CHECK(xu && xu->hi && xg && xg->hi && guided && R > 0 && D % 8 == 0, "cfg_mix: bad arguments");
  const long long nvec = (long long)R * L * (D / 8);
  cfg_mix_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>(mbf(xu), cbf(xg), guided, w, L, D, nvec);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int row_scale_dispatch(const clipdlm_bf_t* g, const float* s_self, const clipdlm_bf_t* ex, const float* s_ex, int R, int L, int D,
                       cudaStream_t st) {
  CLIPDLM_CHECK(g && g->hi && R > 0 && D % 8 == 0, "row_scale: bad arguments");
  const long long nvec = (long long)R * L * (D / 8);
  row_scale_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>(mbf(g), s_self, ex ? mbf(ex) : mbf(nullptr), s_ex, L, D, nvec);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int embed_fwd_dispatch(const clipdlm_embed_t* e, cudaStream_t st) {
  CLIPDLM_CHECK(e != nullptr && e->D % 256 == 0, "embed_fwd: bad descriptor / dim");
  CLIPDLM_CHECK(e->R > 0 && e->B > 0 && e->R % e->B == 0, "embed_fwd: R %d must be a positive multiple of B %d", e->R, e->B);
  CLIPDLM_CHECK(e->fusion == 0 ? e->L == e->Ltxt + 2 : e->L == e->Ltxt, "embed_fwd: L %d inconsistent with Ltxt %d for fusion %d", e->L,
                e->Ltxt, e->fusion);
  CLIPDLM_CHECK(e->mode == 0 ? e->x_in != nullptr : (e->emb_table && e->ids && e->noise && e->coef_a && e->coef_b),
                "embed_fwd: missing inputs for mode %d", e->mode);
  CLIPDLM_CHECK(e->img_proj && e->txt_proj && e->pos && e->ln_w && e->ln_b && e->h.hi, "embed_fwd: null pointer");
  EmbedArgs a;
  a.R = e->R; a.B = e->B; a.Ltxt = e->Ltxt; a.L = e->L; a.D = e->D; a.fusion = e->fusion; a.mode = e->mode; a.guided = e->guided;
  a.x_in = e->x_in; a.x_in_stride = e->x_in_stride > 0 ? e->x_in_stride : (long long)e->Ltxt * e->D;
  a.emb_table = e->emb_table; a.ids = e->ids; a.noise = e->noise; a.coef_a = e->coef_a; a.coef_b = e->coef_b;
  a.img_proj = e->img_proj; a.txt_proj = e->txt_proj; a.seg = e->seg; a.pos = e->pos;
  a.ln_w = e->ln_w; a.ln_b = e->ln_b; a.ln_eps = e->ln_eps;
  a.z = mbf(&e->z); a.h = mbf(&e->h);
  a.drop = make_drop(e->drop_seed, e->drop_site, e->drop_p);
  const long long rows = (long long)e->R * e->L;
  const int grid = (int)((rows + 7) / 8);
  DISPATCH_NV(e->D, embed_fwd_kernel<NV><<<grid, 256, 0, st>>>(a));
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int embed_bwd_dispatch(const clipdlm_bf_t* dz, int R, int B, int Ltxt, int L, int D, int fusion, int guided, float* d_pos, float* d_seg,
                       float* d_img, float* d_txt, cudaStream_t st) {
  CLIPDLM_CHECK(dz && dz->hi && D % 256 == 0 && R % B == 0, "embed_bwd: bad arguments");
  CLIPDLM_CHECK(d_pos && d_img && d_txt && (fusion != 0 || d_seg), "embed_bwd: null gradient buffer");
  int slabs = (4 * num_sms()) / ((D / 256) * L);
  if (slabs < 1) slabs = 1;
  if (slabs > (R + 7) / 8) slabs = (R + 7) / 8;
  const int rps = (R + slabs - 1) / slabs;
  dim3 grid(D / 256, L, (R + rps - 1) / rps);
  embed_bwd_posseg_kernel<<<grid, 256, 0, st>>>(cbf(dz), R, L, Ltxt, D, fusion, rps, d_pos, d_seg);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  const long long n = (long long)B * (D / 8);
  embed_bwd_proj_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(cbf(dz), R, B, L, Ltxt, D, fusion, guided, d_img, d_txt);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_fwd_dispatch(const clipdlm_bf_t* z, const float* w, const float* b, float eps, long long rows, int D, const clipdlm_bf_t* y,
                           float* y_f32, unsigned long long seed, uint32_t site, float p, cudaStream_t st) {
  CLIPDLM_CHECK(z && z->hi && w && b && rows > 0 && ((y && y->hi) || y_f32), "layernorm_fwd: bad arguments");
  const DropoutCfg d = make_drop(seed, site, p);
  if (z->lo == nullptr && (!y || y->lo == nullptr) && D == 1024 && d.thresh16 == 0 && rows >= 64) {
    const size_t smem = (size_t)8 * LNF_STAGES * 1024 * 2 + 8 * LNF_STAGES * sizeof(uint64_t);
    static bool set4 = false;
    if (!set4) {
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_fwd_fast_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_fwd_fast_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set4 = true;
    }
    const long long want = (rows + 7) / 8;
    const int fgrid = (int)(want < 2LL * num_sms() ? want : 2LL * num_sms());
    __nv_bfloat16* yh = y ? (__nv_bfloat16*)y->hi : nullptr;
    if (y_f32 != nullptr) CLIPDLM_CUDA_OK(launch_pdl(layernorm_fwd_fast_kernel<true, 4>, dim3(fgrid), dim3(256), smem, st, (const __nv_bfloat16*)z->hi, w, b, eps, rows, yh, y_f32));
    else CLIPDLM_CUDA_OK(launch_pdl(layernorm_fwd_fast_kernel<false, 4>, dim3(fgrid), dim3(256), smem, st, (const __nv_bfloat16*)z->hi, w, b, eps, rows, yh, (float*)nullptr));
    return 0;
  }
  if (z->lo == nullptr && (!y || y->lo == nullptr) && D == 768 && d.thresh16 == 0 && rows >= 64) {
    const size_t smem = (size_t)8 * LNF_STAGES * 768 * 2 + 8 * LNF_STAGES * sizeof(uint64_t);
    static bool set = false;
    if (!set) {
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_fwd_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_fwd_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set = true;
    }
    const long long want = (rows + 7) / 8;
    const int fgrid = (int)(want < 2LL * num_sms() ? want : 2LL * num_sms());
    __nv_bfloat16* yh = y ? (__nv_bfloat16*)y->hi : nullptr;
    if (y_f32 != nullptr) CLIPDLM_CUDA_OK(launch_pdl(layernorm_fwd_fast_kernel<true>, dim3(fgrid), dim3(256), smem, st, (const __nv_bfloat16*)z->hi, w, b, eps, rows, yh, y_f32));
    else CLIPDLM_CUDA_OK(launch_pdl(layernorm_fwd_fast_kernel<false>, dim3(fgrid), dim3(256), smem, st, (const __nv_bfloat16*)z->hi, w, b, eps, rows, yh, (float*)nullptr));
    return 0;
  }
  const int grid = (int)((rows + 7) / 8);
  DISPATCH_NV(D, layernorm_fwd_kernel<NV><<<grid, 256, 0, st>>>(cbf(z), w, b, eps, rows, D, mbf(y), y_f32, d));
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_bwd_dispatch(const clipdlm_bf_t* z, const clipdlm_bf_t* dy, const float* w, float eps, long long rows, int D,
                           const clipdlm_bf_t* dz, float* dw, float* db, unsigned long long seed, uint32_t site_out, float p_out,
                           const clipdlm_bf_t* dz_drop, uint32_t site_in, float p_in, const clipdlm_bf_t* gelu_u, float* dbias,
                           cudaStream_t st) {
  CLIPDLM_CHECK(z && z->hi && dy && dy->hi && w && dz && dz->hi && rows > 0, "layernorm_bwd: bad arguments");
  LnBwdArgs a;
  a.z = cbf(z); a.dy = cbf(dy); a.w = w; a.eps = eps; a.rows = rows; a.D = D;
  a.dz = mbf(dz); a.dw = dw; a.db = db;
  a.drop_out = make_drop(seed, site_out, p_out);
  a.dz_drop = (dz_drop && p_in > 0.f) ? mbf(dz_drop) : mbf(nullptr);
  a.drop_in = make_drop(seed, site_in, p_in);
  a.gelu_u = cbf(gelu_u);
  a.dbias = dbias;
  long long want = (rows + 3) / 4;
  int grid = (int)(want < 3LL * num_sms() ? want : 3LL * num_sms());  // 3 resident CTAs of 4 warps per SM (register-limited)
  const int narr = a.z.lo != nullptr ? 4 : 2;
  CLIPDLM_CHECK((a.z.lo != nullptr) == (a.dy.lo != nullptr), "layernorm_bwd: z and dy must use the same storage mode");
  const size_t smem = (size_t)4 * D * sizeof(float) + (size_t)4 * LNB_STAGES * narr * D * 2 + 4 * LNB_STAGES * sizeof(uint64_t);
  {
    static size_t smem_set = 0;
    if (smem > smem_set) {
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      smem_set = 120 * 1024;
    }
  }
  // plain bf16, D = 768 (the reference model) or 1024 (bert-large geometry): packed-math specialisations for the feature combinations the engine uses
  if (a.z.lo == nullptr && a.dz.lo == nullptr && (D == 768 || D == 1024) && a.dw != nullptr && a.db != nullptr) {
    if (D == 1024) grid = (int)(want < 2LL * num_sms() ? want : 2LL * num_sms());
    const bool f_do = a.drop_out.thresh16 != 0, f_dd = a.dz_drop.hi != nullptr, f_g = a.gelu_u.hi != nullptr, f_b = a.dbias != nullptr;
    const int combo = (f_do ? 1 : 0) | (f_dd ? 2 : 0) | (f_g ? 4 : 0) | (f_b ? 8 : 0);
    const bool plain_aux = (!f_dd || a.dz_drop.lo == nullptr) && (!f_g || a.gelu_u.lo == nullptr);
#define LNB_FAST_NV(NVV, DO, DD, GE, DB)                                                                                      \
  {                                                                                                                           \
    static bool set = false;                                                                                                  \
    if (!set) {                                                                                                               \
      CLIPDLM_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_fast_kernel<NVV, DO, DD, GE, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024)); \
      set = true;                                                                                                             \
    }                                                                                                                         \
    CLIPDLM_CUDA_OK(launch_pdl(layernorm_bwd_fast_kernel<NVV, DO, DD, GE, DB>, dim3(grid), dim3(128), smem, st, a));          \
    return 0;                                                                                                                 \
  }
#define LNB_FAST(DO, DD, GE, DB) { if (D == 768) LNB_FAST_NV(3, DO, DD, GE, DB) else LNB_FAST_NV(4, DO, DD, GE, DB) }
    if (plain_aux) {
      switch (combo) {
        case 8: LNB_FAST(false, false, false, true)     // sa_layer_norm: + out_lin bias gradient
        case 10: LNB_FAST(false, true, false, true)     // output_layer_norm (train): dropped copy + lin2 bias gradient
        case 12: LNB_FAST(false, false, true, true)     // vocab_layer_norm: * gelu'(u) + vocab_transform bias gradient
        case 1: LNB_FAST(true, false, false, false)     // embedding LayerNorm (train)
        case 0: LNB_FAST(false, false, false, false)    // embedding LayerNorm (eval / p = 0)
        default: break;
      }
    }
#undef LNB_FAST
#undef LNB_FAST_NV
  }
  DISPATCH_NV(D, layernorm_bwd_kernel<NV><<<grid, 128, smem, st>>>(a));
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int colsum_dispatch(const clipdlm_bf_t* x, long long rows, int N, float* out, cudaStream_t st) {
  CLIPDLM_CHECK(x && x->hi && out && N % 256 == 0 && rows > 0, "colsum: bad arguments (N must be a multiple of 256)");
  long long slabs = (4LL * num_sms()) / (N / 256);
  if (slabs < 1) slabs = 1;
  if (slabs > (rows + 31) / 32) slabs = (rows + 31) / 32;
  const long long rps = (rows + slabs - 1) / slabs;
  dim3 grid(N / 256, (unsigned)((rows + rps - 1) / rps));
  colsum_kernel<<<grid, 256, 0, st>>>(cbf(x), rows, N, rps, out);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int embed_loss_dispatch(const clipdlm_bf_t* x_out, const float* emb, const int* ids, const float* tgt, int tgt_rows, int R, int B, int Ltxt,
                        int L, int D, int kind, long long R_total, int batch_size, float weight, double* loss_acc, const clipdlm_bf_t* dx,
                        cudaStream_t st) {
  CLIPDLM_CHECK(x_out && x_out->hi && ((emb && ids) || (tgt && tgt_rows > 0)) && R > 0 && D % 8 == 0 && kind >= 0 && kind <= 3,
                "embed_loss: bad arguments");
  double inv_div;
  switch (kind) {
    case 0: inv_div = 1.0 / ((double)R_total * D); break;            // (..).abs().sum(dim=1).mean()           CLIP-DDPM.py:77-78
    case 1: inv_div = 1.0 / ((double)batch_size * 768.0 * 100.0); break;  // .abs().sum()/BATCH_SIZE/768/100  :80-81 (literals)
    case 2: inv_div = 1.0 / (double)R_total; break;                   // L2 norm per row, mean                  :83-84
    default: inv_div = 1.0 / (double)batch_size; break;               // L2 norm per row, sum / BATCH_SIZE      :86-87
  }
  embed_loss_kernel<<<R, 256, 0, st>>>(cbf(x_out), emb, ids, tgt, tgt_rows, R, B, Ltxt, L, D, kind, inv_div, (float)(inv_div * weight), loss_acc, mbf(dx));
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int small_linear_fwd_dispatch(const float* x, const float* w, const float* b, int B, int K, int N, float* y, cudaStream_t st) {
  CLIPDLM_CHECK(x && w && y && B > 0 && K % 4 == 0, "small_linear_fwd: bad arguments (K must be a multiple of 4)");
  dim3 grid((B + 63) / 64, (N + 63) / 64);
  small_gemm_kernel<true><<<grid, 256, 0, st>>>(x, w, b, B, N, K, y, nullptr, 0);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int small_linear_bwd_dispatch(const float* x, const float* dy, int B, int K, int N, float* dw, float* db, cudaStream_t st) {
  CLIPDLM_CHECK(x && dy && dw && B > 0 && K % 4 == 0 && N % 4 == 0, "small_linear_bwd: bad arguments (K, N must be multiples of 4)");
  dim3 grid((N + 63) / 64, (K + 63) / 64);   // dW [N, K]: rows = output features, reduction over the B rows
  // a long reduction (the TRAIN_EMBEDDING projections reduce over every token of a chunk) is split across blocks; the CLIP
  // projections (B = captions of the batch) keep one block per tile and plain stores
  int splits = 1;
  if (B > 4096) {
    const int tiles = (int)(grid.x * grid.y);
    splits = (4 * num_sms() + tiles - 1) / tiles;
    if (splits > (B + 1023) / 1024) splits = (B + 1023) / 1024;
    if (splits < 1) splits = 1;
  }
  const int k_per_split = ((B + splits - 1) / splits + 15) / 16 * 16;
  grid.z = (unsigned)((B + k_per_split - 1) / k_per_split);
  small_gemm_kernel<false><<<grid, 256, 0, st>>>(dy, x, nullptr, N, K, B, dw, db, k_per_split);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int adamw_dispatch(float* p, float* g, float* m, float* v, void* sh_hi, void* sh_lo, long long n, float lr, float beta1, float beta2,
                   float eps, float wd, int step, float grad_scale, int zero_grad, cudaStream_t st) {
  CLIPDLM_CHECK(p && g && m && v && n > 0 && n % 4 == 0 && step >= 1, "adamw: bad arguments (n must be a multiple of 4, step >= 1)");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float decay = (float)(1.0 - (double)lr * (double)wd);
  const int grid = grid_1d(n / 4, 256, 8 * num_sms());
  adamw_kernel<<<grid, 256, 0, st>>>(p, g, m, v, (__nv_bfloat16*)sh_hi, (__nv_bfloat16*)sh_lo, n / 4, decay, beta1, beta2, eps, step_size,
                                     inv_sqrt_bc2, grad_scale, zero_grad);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

int to_bf16_dispatch(const float* x, void* hi, void* lo, long long n, cudaStream_t st) {
  CLIPDLM_CHECK(x && hi && n > 0, "to_bf16: bad arguments");
  to_bf16_kernel<<<grid_1d(n, 256, 16 * num_sms()), 256, 0, st>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int to_f32_dispatch(const void* hi, const void* lo, float* y, long long n, cudaStream_t st) {
  CLIPDLM_CHECK(hi && y && n > 0, "to_f32: bad arguments");
  to_f32_kernel<<<grid_1d(n, 256, 16 * num_sms()), 256, 0, st>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, y, n);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int gather_rows_f32_dispatch(const clipdlm_bf_t* x, long long rows_out, int len, int stride, int D, float* y, cudaStream_t st) {
  CLIPDLM_CHECK(x && x->hi && y && rows_out > 0 && D % 8 == 0 && len > 0 && stride >= len, "gather_rows_f32: bad arguments");
  gather_rows_f32_kernel<<<grid_1d(rows_out * (D / 8), 256, 16 * num_sms()), 256, 0, st>>>(cbf(x), rows_out, len, stride, D, y);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int q_sample_dispatch(const float* x0, const float* noise, const float* ca, const float* cb, long long n, int S, float* out, cudaStream_t st) {
  CLIPDLM_CHECK(x0 && noise && ca && cb && out && n > 0 && n % 4 == 0 && S > 0, "q_sample: bad arguments");
  q_sample_kernel<<<grid_1d(n / 4 * S, 256, 16 * num_sms()), 256, 0, st>>>(x0, noise, ca, cb, n / 4, S, out);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}
int keymask_dispatch(const int* attn_mask, int R, int B, int Ltxt, int L, int fusion, int guided, uint32_t* km, cudaStream_t st) {
  CLIPDLM_CHECK(km && R > 0 && B > 0, "keymask: bad arguments");
  const int kw = (L + 31) / 32;
  keymask_kernel<<<(R * kw + 255) / 256, 256, 0, st>>>(attn_mask, R, B, Ltxt, L, fusion, guided, kw, km);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace clipdlm
