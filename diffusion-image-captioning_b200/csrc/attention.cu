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

// Dropout RNG for the attention probabilities. Element (query i, key j) of (row, head) pair rh draws 16 bits from the
// Philox block addressed by the *tensor-core fragment coordinates* of (i, j): lane' = (i % 8) * 4 + (j % 8) / 2 is the lane
// that owns the element in an m16n8 accumulator tile, q = (i / 8) * 4 + j / 32 the block index, field = ((j / 8) % 4) * 2 + j % 2.
// The mma kernel therefore needs 4 Philox calls per thread for a 32 x 32 score tile; the SIMT kernel evaluates the same
// function per element, so both produce identical masks (forward, backward, bf16 and split precision).
__device__ __forceinline__ uint4 attn_rand_block(const DropoutCfg& d, unsigned long long rh, int lane_p, int q) {
  return philox4x32_10(make_uint4((uint32_t)rh, (uint32_t)(rh >> 32), (uint32_t)lane_p | ((uint32_t)q << 8), d.site ^ 0xa77e0000u),
                       make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
}
__device__ __forceinline__ uint32_t attn_field16(const uint4& rnd, int f) {
  const uint32_t word = (f >> 1) == 0 ? rnd.x : ((f >> 1) == 1 ? rnd.y : ((f >> 1) == 2 ? rnd.z : rnd.w));
  return (f & 1) ? (word >> 16) : (word & 0xffffu);
}
__device__ __forceinline__ bool attn_keep_ij(const DropoutCfg& d, unsigned long long rh, int i, int j) {
  const uint4 rnd = attn_rand_block(d, rh, (i & 7) * 4 + ((j & 7) >> 1), (i >> 3) * 4 + (j >> 5));
  return attn_field16(rnd, ((j >> 3) & 3) * 2 + (j & 1)) >= d.thresh16;
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
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      float a = s[kg] * inv;
      if (drop.thresh16 != 0) a = attn_keep_ij(drop, (unsigned long long)rh, i, lane + 32 * kg) ? a * drop.scale : 0.f;
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
    float a[KG], dS[KG];
    float dot = 0.f;
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) {
      p[kg] *= inv;
      float keep_scale = 1.f;
      if (drop.thresh16 != 0) keep_scale = attn_keep_ij(drop, (unsigned long long)rh, i, lane + 32 * kg) ? drop.scale : 0.f;
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


// ==================================================================================================================
// Tensor-core path for L <= 32, plain bf16 storage: mma.sync.m16n8k16 (bf16 in, fp32 accumulate) on a padded 32 x 32
// score tile per (row, head).  The per-(row, head) problems are far too small for a 128-row tcgen05 tile (packing 7
// sequences per tile wastes 7x on the block-diagonal), so the warp-level MMA is the right tensor-core instruction here:
// ~64 (fwd) / ~160 (bwd) MMAs per warp instead of ~5 k / ~14 k FMA-issue slots in the SIMT kernels above.
// ==================================================================================================================
constexpr int QP = 72;  // bf16 row pitch of the q / k / v / dO tiles (144 B: ldmatrix rows hit distinct 16-byte bank groups)
constexpr int PP = 40;  // bf16 row pitch of the P / dS tiles (80 B)
constexpr int TILE_E = 32 * QP;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(const __nv_bfloat16* p, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(const __nv_bfloat16* p, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// ldmatrix lane addresses (l = lane, mi = l >> 3 selects the 8x8 matrix the lane supplies a row address for)
//  A  (16 x 16 at (m0, k0)) from row-major X[m][k]            : X[m0 + (l & 7) + (mi & 1) * 8][k0 + (mi >> 1) * 8]
//  A^T(16 x 16 at (m0, k0)) from X[k][m] (use .trans)          : X[k0 + (l & 7) + (mi >> 1) * 8][m0 + (mi & 1) * 8]
//  B  (k16 x two n8 tiles at (k0, n0)) from Bt[n][k]           : Bt[n0 + (l & 7) + (mi >> 1) * 8][k0 + (mi & 1) * 8]   -> {b0, b1 | b0', b1'}
//  B  (k16 x two n8 tiles at (k0, n0)) from B[k][n] (.trans)   : B[k0 + (l & 7) + (mi & 1) * 8][n0 + (mi >> 1) * 8]    -> {b0, b1 | b0', b1'}
__device__ __forceinline__ const __nv_bfloat16* addr_a(const __nv_bfloat16* x, int pitch, int m0, int k0, int l) {
  const int mi = l >> 3;
  return x + (m0 + (l & 7) + (mi & 1) * 8) * pitch + k0 + (mi >> 1) * 8;
}
__device__ __forceinline__ const __nv_bfloat16* addr_at(const __nv_bfloat16* x, int pitch, int m0, int k0, int l) {
  const int mi = l >> 3;
  return x + (k0 + (l & 7) + (mi >> 1) * 8) * pitch + m0 + (mi & 1) * 8;
}
__device__ __forceinline__ const __nv_bfloat16* addr_bt(const __nv_bfloat16* bt, int pitch, int k0, int n0, int l) {
  const int mi = l >> 3;
  return bt + (n0 + (l & 7) + (mi >> 1) * 8) * pitch + k0 + (mi & 1) * 8;
}
__device__ __forceinline__ const __nv_bfloat16* addr_b(const __nv_bfloat16* b, int pitch, int k0, int n0, int l) {
  const int mi = l >> 3;
  return b + (k0 + (l & 7) + (mi & 1) * 8) * pitch + n0 + (mi >> 1) * 8;
}

// stage rows [0, L) of one head slice (64 bf16 per row) into a [32][QP] tile, zero-filling rows >= L
__device__ __forceinline__ void stage_tile(const __nv_bfloat16* __restrict__ src, size_t row_elems, int L, __nv_bfloat16* tile, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 32 + lane;
    const int i = idx >> 3, part = idx & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i < L) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)i * row_elems + part * 8));
    *reinterpret_cast<uint4*>(tile + i * QP + part * 8) = v;
  }
}
__device__ __forceinline__ void unstage_tile(const __nv_bfloat16* tile, __nv_bfloat16* __restrict__ dst, size_t row_elems, int L, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 32 + lane;
    const int i = idx >> 3, part = idx & 7;
    if (i < L) *reinterpret_cast<uint4*>(dst + (size_t)i * row_elems + part * 8) = *reinterpret_cast<const uint4*>(tile + i * QP + part * 8);
  }
}
// C-fragment accumulators [2 m-tiles][8 n-tiles][4] (32 x 64 fp32) -> bf16 tile [32][QP]
__device__ __forceinline__ void acc_to_tile(const float (&o)[2][8][4], __nv_bfloat16* tile, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(tile + (mt * 16 + g) * QP + nt * 8 + 2 * t) = pack_bf16x2(o[mt][nt][0], o[mt][nt][1]);
      *reinterpret_cast<uint32_t*>(tile + (mt * 16 + g + 8) * QP + nt * 8 + 2 * t) = pack_bf16x2(o[mt][nt][2], o[mt][nt][3]);
    }
}

// S = scale * Q K^T on the padded 32 x 32 tile, masked softmax in registers. Returns probabilities in p (C-fragment layout:
// p[mt][nt][e] is row mt*16 + g + (e >> 1) * 8, column nt*8 + 2t + (e & 1)).
__device__ __forceinline__ void scores_softmax(const __nv_bfloat16* q, const __nv_bfloat16* k, uint32_t keybits, int L, float scale,
                                               int lane, float (&p)[2][4][4]) {
  const int t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) p[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a0[4], a1[4], b01[4], b23[4];
    ldsm_x4(addr_a(q, QP, 0, ks * 16, lane), a0);
    ldsm_x4(addr_a(q, QP, 16, ks * 16, lane), a1);
    ldsm_x4(addr_bt(k, QP, ks * 16, 0, lane), b01);
    ldsm_x4(addr_bt(k, QP, ks * 16, 16, lane), b23);
    mma16816(p[0][0], a0, b01[0], b01[1]); mma16816(p[0][1], a0, b01[2], b01[3]);
    mma16816(p[0][2], a0, b23[0], b23[1]); mma16816(p[0][3], a0, b23[2], b23[3]);
    mma16816(p[1][0], a1, b01[0], b01[1]); mma16816(p[1][1], a1, b01[2], b01[3]);
    mma16816(p[1][2], a1, b23[0], b23[1]); mma16816(p[1][3], a1, b23[2], b23[3]);
  }
  const float sl = scale * LOG2E;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int j = nt * 8 + 2 * t + cc;
          const bool ok = j < L && ((keybits >> j) & 1u);
          const float v = ok ? p[mt][nt][hh * 2 + cc] * sl : -INFINITY;
          p[mt][nt][hh * 2 + cc] = v;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const float ev = exp2f(p[mt][nt][hh * 2 + cc] - mx);
          p[mt][nt][hh * 2 + cc] = ev;
          sum += ev;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) p[mt][nt][hh * 2 + cc] *= inv;
    }
}
// keep-scale (0 or 1 / (1 - p)) of every element of the thread's C fragments
__device__ __forceinline__ void dropout_scales(const DropoutCfg& d, unsigned long long rh, int lane, float (&ks)[2][4][4]) {
#pragma unroll
  for (int ib = 0; ib < 4; ++ib) {  // ib = i / 8 = mt * 2 + hh
    const uint4 rnd = attn_rand_block(d, rh, lane, ib * 4);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
        ks[ib >> 1][nt][(ib & 1) * 2 + cc] = attn_field16(rnd, nt * 2 + cc) >= d.thresh16 ? d.scale : 0.f;
  }
}
// C fragments of a 32 x 32 fp32 tile -> A fragments (bf16) for the k16 block kb (keys 16*kb .. 16*kb + 15) of m-tile mt
__device__ __forceinline__ void c_to_a(const float (&c)[2][4][4], int mt, int kb, uint32_t (&a)[4]) {
  a[0] = pack_bf16x2(c[mt][2 * kb][0], c[mt][2 * kb][1]);
  a[1] = pack_bf16x2(c[mt][2 * kb][2], c[mt][2 * kb][3]);
  a[2] = pack_bf16x2(c[mt][2 * kb + 1][0], c[mt][2 * kb + 1][1]);
  a[3] = pack_bf16x2(c[mt][2 * kb + 1][2], c[mt][2 * kb + 1][3]);
}

__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const uint32_t* __restrict__ keymask, int R, int L,
                                                           int D, int H, __nv_bfloat16* __restrict__ ctx, DropoutCfg drop, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rh = (long long)blockIdx.x * 4 + warp;
  if (rh >= (long long)R * H) return;
  const int r = (int)(rh / H), h = (int)(rh % H);
  __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(smem_raw) + (size_t)warp * 3 * TILE_E;
  __nv_bfloat16* k = q + TILE_E;
  __nv_bfloat16* v = k + TILE_E;
  const __nv_bfloat16* base = qkv + (size_t)r * L * 3 * D + h * DH;
  stage_tile(base, 3 * (size_t)D, L, q, lane);
  stage_tile(base + D, 3 * (size_t)D, L, k, lane);
  stage_tile(base + 2 * D, 3 * (size_t)D, L, v, lane);
  __syncwarp();
  float p[2][4][4];
  scores_softmax(q, k, keymask[r], L, scale, lane, p);
  if (drop.thresh16 != 0) {
    float ks[2][4][4];
    dropout_scales(drop, (unsigned long long)rh, lane, ks);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) p[mt][nt][e] *= ks[mt][nt][e];
  }
  float o[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    uint32_t a0[4], a1[4];
    c_to_a(p, 0, kb, a0);
    c_to_a(p, 1, kb, a1);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(addr_b(v, QP, kb * 16, dp * 16, lane), b);
      mma16816(o[0][2 * dp], a0, b[0], b[1]); mma16816(o[0][2 * dp + 1], a0, b[2], b[3]);
      mma16816(o[1][2 * dp], a1, b[0], b[1]); mma16816(o[1][2 * dp + 1], a1, b[2], b[3]);
    }
  }
  __syncwarp();  // all ldmatrix reads of q are long done; reuse its tile as the output staging buffer
  acc_to_tile(o, q, lane);
  __syncwarp();
  unstage_tile(q, ctx + (size_t)r * L * D + h * DH, (size_t)D, L, lane);
}

__global__ void __launch_bounds__(128) attn_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const uint32_t* __restrict__ keymask,
                                                           const __nv_bfloat16* __restrict__ dctx, int R, int L, int D, int H,
                                                           __nv_bfloat16* __restrict__ dqkv, DropoutCfg drop, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rh = (long long)blockIdx.x * 4 + warp;
  if (rh >= (long long)R * H) return;
  const int r = (int)(rh / H), h = (int)(rh % H);
  const int g = lane >> 2, t = lane & 3;
  constexpr int WARP_E = 4 * TILE_E + 2 * 32 * PP;
  __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(smem_raw) + (size_t)warp * WARP_E;
  __nv_bfloat16* k = q + TILE_E;
  __nv_bfloat16* v = k + TILE_E;
  __nv_bfloat16* dO = v + TILE_E;
  __nv_bfloat16* pd = dO + TILE_E;   // dropped probabilities  [32][PP]
  __nv_bfloat16* ds = pd + 32 * PP;  // d(scores)              [32][PP]
  const __nv_bfloat16* base = qkv + (size_t)r * L * 3 * D + h * DH;
  stage_tile(base, 3 * (size_t)D, L, q, lane);
  stage_tile(base + D, 3 * (size_t)D, L, k, lane);
  stage_tile(base + 2 * D, 3 * (size_t)D, L, v, lane);
  stage_tile(dctx + (size_t)r * L * D + h * DH, (size_t)D, L, dO, lane);
  __syncwarp();
  float p[2][4][4];
  scores_softmax(q, k, keymask[r], L, scale, lane, p);
  // dP = dO V^T
  float dp[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dp[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a0[4], a1[4], b01[4], b23[4];
    ldsm_x4(addr_a(dO, QP, 0, ks * 16, lane), a0);
    ldsm_x4(addr_a(dO, QP, 16, ks * 16, lane), a1);
    ldsm_x4(addr_bt(v, QP, ks * 16, 0, lane), b01);
    ldsm_x4(addr_bt(v, QP, ks * 16, 16, lane), b23);
    mma16816(dp[0][0], a0, b01[0], b01[1]); mma16816(dp[0][1], a0, b01[2], b01[3]);
    mma16816(dp[0][2], a0, b23[0], b23[1]); mma16816(dp[0][3], a0, b23[2], b23[3]);
    mma16816(dp[1][0], a1, b01[0], b01[1]); mma16816(dp[1][1], a1, b01[2], b01[3]);
    mma16816(dp[1][2], a1, b23[0], b23[1]); mma16816(dp[1][3], a1, b23[2], b23[3]);
  }
  if (drop.thresh16 != 0) {
    float ksc[2][4][4];
    dropout_scales(drop, (unsigned long long)rh, lane, ksc);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dp[mt][nt][e] *= ksc[mt][nt][e];  // gradient through the dropout
    // pd = p * keep_scale is what multiplied V in the forward
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<uint32_t*>(pd + (mt * 16 + g) * PP + nt * 8 + 2 * t) = pack_bf16x2(p[mt][nt][0] * ksc[mt][nt][0], p[mt][nt][1] * ksc[mt][nt][1]);
        *reinterpret_cast<uint32_t*>(pd + (mt * 16 + g + 8) * PP + nt * 8 + 2 * t) = pack_bf16x2(p[mt][nt][2] * ksc[mt][nt][2], p[mt][nt][3] * ksc[mt][nt][3]);
      }
  } else {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<uint32_t*>(pd + (mt * 16 + g) * PP + nt * 8 + 2 * t) = pack_bf16x2(p[mt][nt][0], p[mt][nt][1]);
        *reinterpret_cast<uint32_t*>(pd + (mt * 16 + g + 8) * PP + nt * 8 + 2 * t) = pack_bf16x2(p[mt][nt][2], p[mt][nt][3]);
      }
  }
  // dS = P o (dP - rowsum(dP o P)) * scale
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float dot = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) dot += dp[mt][nt][hh * 2 + cc] * p[mt][nt][hh * 2 + cc];
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        *reinterpret_cast<uint32_t*>(ds + (mt * 16 + g + hh * 8) * PP + nt * 8 + 2 * t) =
            pack_bf16x2(p[mt][nt][hh * 2] * (dp[mt][nt][hh * 2] - dot) * scale, p[mt][nt][hh * 2 + 1] * (dp[mt][nt][hh * 2 + 1] - dot) * scale);
    }
  __syncwarp();
  float acc[2][8][4];
  // ---- dV[j][d] = sum_i pd[i][j] dO[i][d]  -> staged into the v tile (v is dead after dP)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {  // reduction over queries i in blocks of 16
    uint32_t a0[4], a1[4];
    ldsm_x4_t(addr_at(pd, PP, 0, kb * 16, lane), a0);
    ldsm_x4_t(addr_at(pd, PP, 16, kb * 16, lane), a1);
#pragma unroll
    for (int dpi = 0; dpi < 4; ++dpi) {
      uint32_t b[4];
      ldsm_x4_t(addr_b(dO, QP, kb * 16, dpi * 16, lane), b);
      mma16816(acc[0][2 * dpi], a0, b[0], b[1]); mma16816(acc[0][2 * dpi + 1], a0, b[2], b[3]);
      mma16816(acc[1][2 * dpi], a1, b[0], b[1]); mma16816(acc[1][2 * dpi + 1], a1, b[2], b[3]);
    }
  }
  __syncwarp();
  acc_to_tile(acc, v, lane);
  // ---- dK[j][d] = sum_i dS[i][j] Q[i][d]  -> staged into the dO tile (dead after dV; the syncwarp below orders the reads)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    uint32_t a0[4], a1[4];
    ldsm_x4_t(addr_at(ds, PP, 0, kb * 16, lane), a0);
    ldsm_x4_t(addr_at(ds, PP, 16, kb * 16, lane), a1);
#pragma unroll
    for (int dpi = 0; dpi < 4; ++dpi) {
      uint32_t b[4];
      ldsm_x4_t(addr_b(q, QP, kb * 16, dpi * 16, lane), b);
      mma16816(acc[0][2 * dpi], a0, b[0], b[1]); mma16816(acc[0][2 * dpi + 1], a0, b[2], b[3]);
      mma16816(acc[1][2 * dpi], a1, b[0], b[1]); mma16816(acc[1][2 * dpi + 1], a1, b[2], b[3]);
    }
  }
  __syncwarp();
  acc_to_tile(acc, dO, lane);
  // ---- dQ[i][d] = sum_j dS[i][j] K[j][d]  -> staged into the q tile (dead after dK)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {  // reduction over keys j
    uint32_t a0[4], a1[4];
    ldsm_x4(addr_a(ds, PP, 0, kb * 16, lane), a0);
    ldsm_x4(addr_a(ds, PP, 16, kb * 16, lane), a1);
#pragma unroll
    for (int dpi = 0; dpi < 4; ++dpi) {
      uint32_t b[4];
      ldsm_x4_t(addr_b(k, QP, kb * 16, dpi * 16, lane), b);
      mma16816(acc[0][2 * dpi], a0, b[0], b[1]); mma16816(acc[0][2 * dpi + 1], a0, b[2], b[3]);
      mma16816(acc[1][2 * dpi], a1, b[0], b[1]); mma16816(acc[1][2 * dpi + 1], a1, b[2], b[3]);
    }
  }
  __syncwarp();
  acc_to_tile(acc, q, lane);
  __syncwarp();
  __nv_bfloat16* out = dqkv + (size_t)r * L * 3 * D + h * DH;
  unstage_tile(q, out, 3 * (size_t)D, L, lane);
  unstage_tile(dO, out + D, 3 * (size_t)D, L, lane);
  unstage_tile(v, out + 2 * D, 3 * (size_t)D, L, lane);
}

static inline CBfPtr cbf(const clipdlm_bf_t* p) {
  CBfPtr r; r.hi = (const __nv_bfloat16*)p->hi; r.lo = (const __nv_bfloat16*)p->lo; return r;
}
static inline BfPtr mbf(const clipdlm_bf_t* p) {
  BfPtr r; r.hi = (__nv_bfloat16*)p->hi; r.lo = (__nv_bfloat16*)p->lo; return r;
}

static bool g_force_simt = false;
void attn_force_simt(int on) { g_force_simt = on != 0; }

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
  if (qkv->lo == nullptr && ctx->lo == nullptr && L <= 32 && !g_force_simt) {
    const size_t smem = (size_t)4 * 3 * TILE_E * sizeof(__nv_bfloat16);
    if (set_smem(attn_fwd_mma_kernel, smem)) return -1;
    const long long items = (long long)R * H;
    attn_fwd_mma_kernel<<<(unsigned)((items + 3) / 4), 128, smem, st>>>((const __nv_bfloat16*)qkv->hi, keymask, R, L, D, H, (__nv_bfloat16*)ctx->hi,
                                                                       make_drop(seed, site, p), 0.125f);
    CLIPDLM_CUDA_OK(cudaGetLastError());
    return 0;
  }
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
  if (qkv->lo == nullptr && dctx->lo == nullptr && dqkv->lo == nullptr && L <= 32 && !g_force_simt) {
    const size_t smem = (size_t)4 * (4 * TILE_E + 2 * 32 * PP) * sizeof(__nv_bfloat16);
    if (set_smem(attn_bwd_mma_kernel, smem)) return -1;
    const long long items = (long long)R * H;
    attn_bwd_mma_kernel<<<(unsigned)((items + 3) / 4), 128, smem, st>>>((const __nv_bfloat16*)qkv->hi, keymask, (const __nv_bfloat16*)dctx->hi, R, L, D,
                                                                       H, (__nv_bfloat16*)dqkv->hi, make_drop(seed, site, p), 0.125f);
    CLIPDLM_CUDA_OK(cudaGetLastError());
    return 0;
  }
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
