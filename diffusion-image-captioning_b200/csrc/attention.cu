// Multi-head self-attention for very short sequences (L = MAX_LENGTH + 2 = 18 on the reference path, <= 128 supported),
// head dim 64. Replaces DistilBertSelfAttention / SDPA (HF modeling_distilbert.py:126-151,177-207) and its autograd backward.
//
// One warp owns one (sequence row r, head h): q, k, v of that head (L x 64 each) are staged once in shared memory with
// coalesced 16-byte loads, scores are computed with lane j owning key j, softmax / dropout statistics by warp shuffles,
// and the P.V products with lane d owning output dims (d, d + 32).  Nothing but the context (fwd) or d(qkv) (bwd) is
// written to HBM: the backward recomputes the probabilities from q, k and the key mask, and regenerates the dropout mask
// from the counter-based RNG, so no [R, H, L, L] tensor ever exists in memory.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/clipdlm.h"
#include <cudaTypedefs.h>

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
  return philox4x32(make_uint4((uint32_t)rh, (uint32_t)(rh >> 32), (uint32_t)lane_p | ((uint32_t)q << 8), d.site ^ 0xa77e0000u),
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
// Tensor-core path for L <= 32, plain bf16 storage.
//
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate) on a padded 32 x 32 score tile per (row, head): the per-(row, head) problems are
// far too small for a 128-row tcgen05 tile, so the warp-level MMA is the right tensor-core instruction here.  What bounds this
// kernel is HBM latency, not math, hence the structure: ONE persistent CTA per SM = 1 TMA producer warp + W consumer warps
// around a ring of S shared-memory stages (S > W).  The producer streams the q / k / v (/ dO) head slices of task after task
// into the ring with cp.async.bulk.tensor (128-byte swizzle, one 18 x 128 B box per tensor, completion on an mbarrier), always
// S - W tasks ahead of the math; a consumer warp waits for its stage, computes straight out of it with ldmatrix on the swizzled
// rows (conflict-free, no padding), stages its result in tile space that is already dead, writes it out with coalesced 16-byte
// stores and hands the stage back.  P^T / dS^T for the backward products come from movmatrix (register transposes): no score
// tile ever touches shared memory.  The first version (one warp per task, per-lane LDG/STS staging, 4-8 warps per SM because
// of the staging tiles) sat at 2 TB/s with the load latency fully exposed.
// ==================================================================================================================
constexpr float LOG2E = 1.4426950408889634f;

int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, unsigned long long inner, unsigned long long outer, unsigned long long pitch_bytes,
                      uint32_t box_inner, uint32_t box_outer);

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {   // 8 x 8 b16 block held in the fragment layout -> its transpose
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Tile = rows of 128 B (one head slice row), 16-byte chunk c of row i stored at chunk c ^ (i & 7) (what TMA SWIZZLE_128B writes
// into a 1024-byte aligned tile).
__device__ __forceinline__ uint32_t sw_addr(uint32_t tile, int row, int chunk) { return tile + row * 128 + ((chunk ^ (row & 7)) << 4); }
// ldmatrix lane addresses (l = lane; matrix index mi = l >> 3 selects which 8 x 8 block the lane addresses a row of)
//  A  (16 x 16 at (m0, k0)) from row-major X[m][k]           : row m0 + (l & 7) + (mi & 1) * 8, chunk k0 / 8 + (mi >> 1)
//  B  (k16 x two n8 tiles at (k0, n0)) from Bt[n][k]          : row n0 + (l & 7) + (mi >> 1) * 8, chunk k0 / 8 + (mi & 1)
//  B  (k16 x two n8 tiles at (k0, n0)) from B[k][n] (.trans)  : row k0 + (l & 7) + (mi & 1) * 8, chunk n0 / 8 + (mi >> 1)
__device__ __forceinline__ uint32_t addr_a(uint32_t tile, int m0, int k0, int l) {
  return sw_addr(tile, m0 + (l & 7) + ((l >> 3) & 1) * 8, (k0 >> 3) + (l >> 4));
}
__device__ __forceinline__ uint32_t addr_bt(uint32_t tile, int k0, int n0, int l) {
  return sw_addr(tile, n0 + (l & 7) + (l >> 4) * 8, (k0 >> 3) + ((l >> 3) & 1));
}
__device__ __forceinline__ uint32_t addr_b(uint32_t tile, int k0, int n0, int l) {
  return sw_addr(tile, k0 + (l & 7) + ((l >> 3) & 1) * 8, (n0 >> 3) + (l >> 4));
}

struct AttGeo {
  int L, MT, NT, KB;      // sequence length; 16-row query tiles, 8-column key tiles, 16-key reduction blocks that hold real data
  int tile_rows;          // rows a tile occupies in shared memory (24 or 32); rows [L, tile_rows) are never written by TMA
  uint32_t tile_bytes;
};
__device__ __forceinline__ AttGeo att_geo(int L) {
  AttGeo g;
  g.L = L; g.MT = (L + 15) >> 4; g.NT = (L + 7) >> 3; g.KB = (L + 15) >> 4;
  g.tile_rows = L <= 24 ? 24 : 32;
  g.tile_bytes = (uint32_t)g.tile_rows * 128u;
  return g;
}

// acc (+)= X[m][:] . Y[n][:]^T over the 64-wide head dim: A = X tile (queries), B = Y tile (keys) - S = Q K^T and dP = dO V^T
__device__ __forceinline__ void qk_product(uint32_t x, uint32_t y, const AttGeo& ge, int lane, float (&p)[2][4][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) p[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a0[4], a1[4] = {0u, 0u, 0u, 0u}, b01[4], b23[4] = {0u, 0u, 0u, 0u};
    ldsm_x4(addr_a(x, 0, ks * 16, lane), a0);
    if (ge.MT > 1) ldsm_x4(addr_a(x, 16, ks * 16, lane), a1);
    ldsm_x4(addr_bt(y, ks * 16, 0, lane), b01);
    if (ge.NT > 2) ldsm_x4(addr_bt(y, ks * 16, 16, lane), b23);   // keys 24..31 of a 24-row tile alias the next tile: never used (NT <= 3)
    mma16816(p[0][0], a0, b01[0], b01[1]);
    if (ge.NT > 1) mma16816(p[0][1], a0, b01[2], b01[3]);
    if (ge.NT > 2) mma16816(p[0][2], a0, b23[0], b23[1]);
    if (ge.NT > 3) mma16816(p[0][3], a0, b23[2], b23[3]);
    if (ge.MT > 1) {
      mma16816(p[1][0], a1, b01[0], b01[1]);
      if (ge.NT > 1) mma16816(p[1][1], a1, b01[2], b01[3]);
      if (ge.NT > 2) mma16816(p[1][2], a1, b23[0], b23[1]);
      if (ge.NT > 3) mma16816(p[1][3], a1, b23[2], b23[3]);
    }
  }
}
// masked softmax of scale * S in registers (C-fragment layout: p[mt][nt][e] is row mt*16 + g + (e >> 1) * 8, column nt*8 + 2t + (e & 1))
__device__ __forceinline__ void softmax_rows(float (&p)[2][4][4], uint32_t keybits, int L, float scale, int lane) {
  const int t = lane & 3;
  const float sl = scale * LOG2E;
  // visible keys clipped to the length, shifted so that bit (nt * 8 + cc) is this lane's column nt * 8 + 2 t + cc: one constant-position
  // bit test per score
  const uint32_t kb = (keybits & (L >= 32 ? 0xffffffffu : ((1u << L) - 1u))) >> (2 * t);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const bool ok = (kb & (1u << (nt * 8 + cc))) != 0u;
          const float v = ok ? p[mt][nt][hh * 2 + cc] * sl : -INFINITY;
          p[mt][nt][hh * 2 + cc] = v;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      if (mx == -INFINITY) mx = 0.f;   // garbage query rows (>= L) of a fully NaN/-inf tile: keep the arithmetic finite
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const float ev = ex2_ftz(p[mt][nt][hh * 2 + cc] - mx);
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
// keep-scale (0 or 1 / (1 - p)) of every element of the thread's C fragments; row blocks of 8 beyond L draw nothing
__device__ __forceinline__ void dropout_scales(const DropoutCfg& d, unsigned long long rh, int lane, int L, float (&ks)[2][4][4]) {
#pragma unroll
  for (int ib = 0; ib < 4; ++ib) {  // ib = i / 8 = mt * 2 + hh
    uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
    if (ib * 8 < L) rnd = attn_rand_block(d, rh, lane, ib * 4);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
        ks[ib >> 1][nt][(ib & 1) * 2 + cc] = attn_field16(rnd, nt * 2 + cc) >= d.thresh16 ? d.scale : 0.f;
  }
}
// C fragments of a 32 x 32 fp32 tile -> A fragments (bf16) for the k16 block kb (columns 16*kb .. 16*kb + 15) of m-tile mt
__device__ __forceinline__ void c_to_a(const float (&c)[2][4][4], int mt, int kb, uint32_t (&a)[4]) {
  a[0] = pack_bf16x2(c[mt][2 * kb][0], c[mt][2 * kb][1]);
  a[1] = pack_bf16x2(c[mt][2 * kb][2], c[mt][2 * kb][3]);
  a[2] = pack_bf16x2(c[mt][2 * kb + 1][0], c[mt][2 * kb + 1][1]);
  a[3] = pack_bf16x2(c[mt][2 * kb + 1][2], c[mt][2 * kb + 1][3]);
}
// acc[mt][0..7] (+)= A (given as C fragments of a [32 x 32] matrix, reduction over its columns) . B tile [k][d] (64 wide)
__device__ __forceinline__ void av_product(const float (&c)[2][4][4], uint32_t btile, const AttGeo& ge, int lane, float (&o)[2][8][4]) {
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    if (kb >= ge.KB) break;
    const bool hi_ok = kb * 16 + 16 <= ge.tile_rows;   // rows 24..31 of a 24-row tile are not ours: feed zeros instead
    uint32_t a0[4], a1[4];
    c_to_a(c, 0, kb, a0);
    c_to_a(c, 1, kb, a1);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(addr_b(btile, kb * 16, dp * 16, lane), b);
      if (!hi_ok) { b[1] = 0u; b[3] = 0u; }
      mma16816(o[0][2 * dp], a0, b[0], b[1]); mma16816(o[0][2 * dp + 1], a0, b[2], b[3]);
      if (ge.MT > 1) { mma16816(o[1][2 * dp], a1, b[0], b[1]); mma16816(o[1][2 * dp + 1], a1, b[2], b[3]); }
    }
  }
}
// acc[mj][0..7] (+)= X^T . B tile: xt[jb][ib] is the transposed 8 x 8 block (rows ib, columns jb) of X in fragment layout, the
// reduction runs over X's rows (queries), the result rows are X's columns (keys)
__device__ __forceinline__ void atv_product(const uint32_t (&xt)[4][4], uint32_t btile, const AttGeo& ge, int lane, float (&o)[2][8][4]) {
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    if (kb >= ge.KB) break;
    const bool hi_ok = kb * 16 + 16 <= ge.tile_rows;
    const uint32_t a0[4] = {xt[0][2 * kb], xt[1][2 * kb], xt[0][2 * kb + 1], xt[1][2 * kb + 1]};
    const uint32_t a1[4] = {xt[2][2 * kb], xt[3][2 * kb], xt[2][2 * kb + 1], xt[3][2 * kb + 1]};
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(addr_b(btile, kb * 16, dp * 16, lane), b);
      if (!hi_ok) { b[1] = 0u; b[3] = 0u; }
      mma16816(o[0][2 * dp], a0, b[0], b[1]); mma16816(o[0][2 * dp + 1], a0, b[2], b[3]);
      if (ge.MT > 1) { mma16816(o[1][2 * dp], a1, b[0], b[1]); mma16816(o[1][2 * dp + 1], a1, b[2], b[3]); }
    }
  }
}
__device__ __forceinline__ void zero_acc(float (&o)[2][8][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
}
// C-fragment accumulators (32 x 64 fp32) -> bf16 rows [0, L) of a swizzled tile
__device__ __forceinline__ void acc_to_tile(const float (&o)[2][8][4], uint32_t tile, int L, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = mt * 16 + g + hh * 8;
      if (row < L) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw_addr(tile, row, nt) + 4 * t), "r"(pack_bf16x2(o[mt][nt][hh * 2], o[mt][nt][hh * 2 + 1])) : "memory");
      }
    }
}
// rows [0, L) of a swizzled tile -> global rows of 64 bf16 (row pitch row_elems), coalesced 16-byte stores
__device__ __forceinline__ void unstage_tile(uint32_t tile, __nv_bfloat16* __restrict__ dst, size_t row_elems, int L, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 32 + lane;
    const int i = idx >> 3, part = idx & 7;
    if (i < L) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw_addr(tile, i, part)));
      *reinterpret_cast<uint4*>(dst + (size_t)i * row_elems + part * 8) = v;
    }
  }
}

struct AttArgs {
  const uint32_t* keymask;
  __nv_bfloat16* out;      // ctx [T, D] (forward) or dqkv [T, 3D] (backward)
  int R, L, D, H, stages, consumers;
  DropoutCfg drop;
  float scale;
};

template <bool BWD>
__global__ void __launch_bounds__(BWD ? 384 : 512, 1) attn_ring_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                                                           const AttArgs a) {
  constexpr int NTILES = BWD ? 4 : 3;
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~uintptr_t(1023));
  const AttGeo ge = att_geo(a.L);
  const uint32_t stage_bytes = NTILES * ge.tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * stage_bytes + 1024);   // 1 KB pad: aliased ldmatrix rows of the last tile
  uint64_t* empty_bar = full_bar + a.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tasks = (long long)a.R * a.H;
  const long long my_tasks = tasks > blockIdx.x ? (tasks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // zero the ring once: rows [L, tile_rows) are never written by TMA and must stay finite (they meet exact zeros in the MMAs)
  for (uint32_t off = threadIdx.x * 16; off < (uint32_t)a.stages * stage_bytes + 1024; off += blockDim.x * 16)
    *reinterpret_cast<uint4*>(ring + off) = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tm_qkv);
    if (BWD) tma_prefetch_desc(&tm_do);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zero fill before async-proxy (TMA) writes
  __syncthreads();
  pdl_wait();

  if (warp == a.consumers) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (long long n = 0; n < my_tasks; ++n) {
        const long long task = blockIdx.x + n * gridDim.x;
        const int r = (int)(task / a.H), h = (int)(task % a.H);
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = ring + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], NTILES * a.L * 128);
        tma_load_2d(st, &tm_qkv, &full_bar[s], h * DH, r * a.L);                              // q
        tma_load_2d(st + ge.tile_bytes, &tm_qkv, &full_bar[s], a.D + h * DH, r * a.L);        // k
        tma_load_2d(st + 2 * ge.tile_bytes, &tm_qkv, &full_bar[s], 2 * a.D + h * DH, r * a.L);  // v
        if (BWD) tma_load_2d(st + 3 * ge.tile_bytes, &tm_do, &full_bar[s], h * DH, r * a.L);  // dO
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < a.consumers) {
    // ===================================== consumers =====================================
    const int g = lane >> 2;
    for (long long n = warp; n < my_tasks; n += a.consumers) {
      const long long task = blockIdx.x + n * gridDim.x;
      const int r = (int)(task / a.H), h = (int)(task % a.H);
      const int s = (int)(n % a.stages);
      const uint32_t ph = (uint32_t)((n / a.stages) & 1);
      const uint32_t keybits = a.keymask[r];
      mbar_wait(&full_bar[s], ph);
      const uint32_t q = smem_u32(ring + (size_t)s * stage_bytes), k = q + ge.tile_bytes, v = k + ge.tile_bytes, dO = v + ge.tile_bytes;
      float p[2][4][4];
      qk_product(q, k, ge, lane, p);
      softmax_rows(p, keybits, a.L, a.scale, lane);
      if (!BWD) {
        if (a.drop.thresh16 != 0) {
          float ks[2][4][4];
          dropout_scales(a.drop, (unsigned long long)task, lane, a.L, ks);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) p[mt][nt][e] *= ks[mt][nt][e];
        }
        float o[2][8][4];
        zero_acc(o);
        av_product(p, v, ge, lane, o);
        __syncwarp();                      // every lane's ldmatrix reads of q are done: reuse its tile as the output staging buffer
        acc_to_tile(o, q, a.L, lane);
        __syncwarp();
        unstage_tile(q, a.out + (size_t)r * a.L * a.D + h * DH, (size_t)a.D, a.L, lane);
      } else {
        float dp[2][4][4];
        qk_product(dO, v, ge, lane, dp);   // dP = dO V^T
        if (a.drop.thresh16 != 0) {
          float ksc[2][4][4];
          dropout_scales(a.drop, (unsigned long long)task, lane, a.L, ksc);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) { dp[mt][nt][e] *= ksc[mt][nt][e]; ksc[mt][nt][e] *= p[mt][nt][e]; }  // ksc := dropped probabilities
          // dV[j][d] = sum_i pd[i][j] dO[i][d]   (pd = dropped probabilities, what multiplied V in the forward)
          uint32_t xt[4][4];
#pragma unroll
          for (int ib = 0; ib < 4; ++ib)
#pragma unroll
            for (int jb = 0; jb < 4; ++jb) {
              const bool live = ib * 8 + g < a.L;   // query rows >= L carry garbage: they must not enter the reduction over queries
              const uint32_t blk = live ? pack_bf16x2(ksc[ib >> 1][jb][(ib & 1) * 2], ksc[ib >> 1][jb][(ib & 1) * 2 + 1]) : 0u;
              xt[jb][ib] = movm_t(blk);
            }
          float acc[2][8][4];
          zero_acc(acc);
          atv_product(xt, dO, ge, lane, acc);
          __syncwarp();                    // v is dead since dP: stage dV there
          acc_to_tile(acc, v, a.L, lane);
        } else {
          uint32_t xt[4][4];
#pragma unroll
          for (int ib = 0; ib < 4; ++ib)
#pragma unroll
            for (int jb = 0; jb < 4; ++jb) {
              const bool live = ib * 8 + g < a.L;
              const uint32_t blk = live ? pack_bf16x2(p[ib >> 1][jb][(ib & 1) * 2], p[ib >> 1][jb][(ib & 1) * 2 + 1]) : 0u;
              xt[jb][ib] = movm_t(blk);
            }
          float acc[2][8][4];
          zero_acc(acc);
          atv_product(xt, dO, ge, lane, acc);
          __syncwarp();
          acc_to_tile(acc, v, a.L, lane);
        }
        // dS = P o (dP - rowsum(dP o P)) * scale, in place in dp (rows >= L zeroed)
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
            const bool live = mt * 16 + hh * 8 + g < a.L;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int cc = 0; cc < 2; ++cc)
                dp[mt][nt][hh * 2 + cc] = live ? p[mt][nt][hh * 2 + cc] * (dp[mt][nt][hh * 2 + cc] - dot) * a.scale : 0.f;
          }
        // dK[j][d] = sum_i dS[i][j] Q[i][d]
        {
          uint32_t xt[4][4];
#pragma unroll
          for (int ib = 0; ib < 4; ++ib)
#pragma unroll
            for (int jb = 0; jb < 4; ++jb)
              xt[jb][ib] = movm_t(pack_bf16x2(dp[ib >> 1][jb][(ib & 1) * 2], dp[ib >> 1][jb][(ib & 1) * 2 + 1]));
          float acc[2][8][4];
          zero_acc(acc);
          atv_product(xt, q, ge, lane, acc);
          __syncwarp();                    // dO is dead since dV: stage dK there
          acc_to_tile(acc, dO, a.L, lane);
        }
        // dQ[i][d] = sum_j dS[i][j] K[j][d]
        {
          float acc[2][8][4];
          zero_acc(acc);
          av_product(dp, k, ge, lane, acc);
          __syncwarp();                    // q is dead since dK: stage dQ there
          acc_to_tile(acc, q, a.L, lane);
        }
        __syncwarp();
        __nv_bfloat16* out = a.out + (size_t)r * a.L * 3 * a.D + h * DH;
        unstage_tile(q, out, 3 * (size_t)a.D, a.L, lane);
        unstage_tile(dO, out + a.D, 3 * (size_t)a.D, a.L, lane);
        unstage_tile(v, out + 2 * a.D, 3 * (size_t)a.D, a.L, lane);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our generic-proxy accesses to the stage before TMA overwrites it
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
}

static inline CBfPtr cbf(const clipdlm_bf_t* p) {
  CBfPtr r; r.hi = (const __nv_bfloat16*)p->hi; r.lo = (const __nv_bfloat16*)p->lo; return r;
}
static inline BfPtr mbf(const clipdlm_bf_t* p) {
  BfPtr r; r.hi = (__nv_bfloat16*)p->hi; r.lo = (__nv_bfloat16*)p->lo; return r;
}

// 0 = auto (L = 16 / 18: back-to-back packed tcgen05 tiles, attention_packed.cu; other L <= 32: forward tcgen05 32-row slots, backward mma.sync
// TMA ring), 1 = fp32 SIMT, 2 = ring, 3 = tcgen05 with 32-row slots (attention_umma.cu), 4 = as 0 but the backward without software pipelining
static int g_force_path = 0;
bool attn_packed_supported(int L, int D, int H);
void attn_packed_bwd_pipeline(int on);
template <bool BWD>
int launch_attn_packed(const __nv_bfloat16* qkv, const __nv_bfloat16* qkv_lo, const __nv_bfloat16* dctx, const __nv_bfloat16* dctx_lo,
                       const uint32_t* keymask, int R, int L, int D, int H, __nv_bfloat16* out, __nv_bfloat16* out_lo, float* dbias,
                       const DropoutCfg& drop, cudaStream_t st);
void attn_force_simt(int on) { g_force_path = on; }
template <bool BWD>
int launch_attn_umma(const __nv_bfloat16* qkv, const __nv_bfloat16* dctx, const uint32_t* keymask, int R, int L, int D, int H,
                     __nv_bfloat16* out, const DropoutCfg& drop, cudaStream_t st);

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

// Persistent TMA-ring launch (plain bf16, L <= 32): one CTA per SM, ring depth from the shared-memory budget.
template <bool BWD>
static int launch_ring(const __nv_bfloat16* qkv, const __nv_bfloat16* dctx, const uint32_t* keymask, int R, int L, int D, int H,
                       __nv_bfloat16* out, const DropoutCfg& drop, cudaStream_t st) {
  const long long T = (long long)R * L;
  CUtensorMap tm_qkv, tm_do;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tm_qkv, qkv, 3ull * D, (unsigned long long)T, 3ull * D * 2, DH, (uint32_t)L))) return rc;
  tm_do = tm_qkv;
  if (BWD && (rc = make_tmap_2d_bf16(&tm_do, dctx, (unsigned long long)D, (unsigned long long)T, (unsigned long long)D * 2, DH, (uint32_t)L))) return rc;
  const uint32_t tile_bytes = (L <= 24 ? 24u : 32u) * 128u;
  const uint32_t stage_bytes = (BWD ? 4u : 3u) * tile_bytes;
  AttArgs a;
  a.keymask = keymask; a.out = out; a.R = R; a.L = L; a.D = D; a.H = H; a.drop = drop; a.scale = 0.125f;  // 1 / sqrt(64)
  a.stages = (int)((216u * 1024u) / stage_bytes);
  if (a.stages > 32) a.stages = 32;
  a.consumers = BWD ? 10 : 14;
  if (BWD) {   // tuning knob (tools/attn_perf.py): the backward's launch bound of 384 threads leaves room for 11 consumer warps + the producer
    static const int env_consumers = [] { const char* v = getenv("CLIPDLM_ATTN_BWD_CONSUMERS"); return v ? atoi(v) : 0; }();
    if (env_consumers >= 1 && env_consumers <= 11) a.consumers = env_consumers;
  }
  if (a.consumers > a.stages - 4) a.consumers = a.stages - 4;
  const size_t smem = 1024 + (size_t)a.stages * stage_bytes + 1024 + (size_t)2 * a.stages * sizeof(uint64_t);
  static bool attr_set = false;
  if (!attr_set) {
    if (set_smem(attn_ring_kernel<BWD>, 227 * 1024)) return -1;
    attr_set = true;
  }
  const long long tasks = (long long)R * H;
  const int grid = (int)(tasks < num_sms() ? tasks : num_sms());
  CLIPDLM_CUDA_OK(launch_pdl(attn_ring_kernel<BWD>, dim3(grid), dim3((a.consumers + 1) * 32), smem, st, tm_qkv, tm_do, a));
  return 0;
}

int attn_fwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, int R, int L, int D, int H, const clipdlm_bf_t* ctx,
                      unsigned long long seed, uint32_t site, float p, cudaStream_t st) {
  CLIPDLM_CHECK(qkv && qkv->hi && ctx && ctx->hi && keymask, "attn_fwd: null pointer");
  CLIPDLM_CHECK(H > 0 && D == H * DH, "attn_fwd: head dim must be 64 (D %d, H %d)", D, H);
  CLIPDLM_CHECK(L >= 1 && L <= 128, "attn_fwd: L %d out of range (1..128)", L);
  if ((qkv->lo == nullptr) == (ctx->lo == nullptr) && (g_force_path == 0 || g_force_path == 4) && attn_packed_supported(L, D, H))
    return launch_attn_packed<false>((const __nv_bfloat16*)qkv->hi, (const __nv_bfloat16*)qkv->lo, nullptr, nullptr, keymask, R, L, D, H,
                                     (__nv_bfloat16*)ctx->hi, (__nv_bfloat16*)ctx->lo, nullptr, make_drop(seed, site, p), st);
  if (qkv->lo == nullptr && ctx->lo == nullptr && ((L <= 32 && (g_force_path == 0 || g_force_path == 3)) || (L > 32 && g_force_path != 1)))
    return launch_attn_umma<false>((const __nv_bfloat16*)qkv->hi, nullptr, keymask, R, L, D, H, (__nv_bfloat16*)ctx->hi, make_drop(seed, site, p), st);
  if (qkv->lo == nullptr && ctx->lo == nullptr && L <= 32 && g_force_path == 2)
    return launch_ring<false>((const __nv_bfloat16*)qkv->hi, nullptr, keymask, R, L, D, H, (__nv_bfloat16*)ctx->hi, make_drop(seed, site, p), st);
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

// dbias (nullable): d(qkv bias) [3D] fp32. When the kernel that runs can fold the bias gradients into its epilogue it accumulates d(q bias) and
// d(v bias) there (d(k bias) is analytically zero) and sets *folded = 1; otherwise *folded = 0 and the caller runs its column-sum pass over dqkv.
int attn_bwd_dispatch(const clipdlm_bf_t* qkv, const uint32_t* keymask, const clipdlm_bf_t* dctx, int R, int L, int D, int H,
                      const clipdlm_bf_t* dqkv, unsigned long long seed, uint32_t site, float p, cudaStream_t st, float* dbias, int* folded) {
  CLIPDLM_CHECK(qkv && qkv->hi && dctx && dctx->hi && dqkv && dqkv->hi && keymask, "attn_bwd: null pointer");
  CLIPDLM_CHECK(H > 0 && D == H * DH, "attn_bwd: head dim must be 64 (D %d, H %d)", D, H);
  CLIPDLM_CHECK(L >= 1 && L <= 128, "attn_bwd: L %d out of range (1..128)", L);
  if (folded) *folded = 0;
  const bool all_pair = qkv->lo != nullptr && dctx->lo != nullptr && dqkv->lo != nullptr;
  const bool all_plain = qkv->lo == nullptr && dctx->lo == nullptr && dqkv->lo == nullptr;
  if ((all_plain || all_pair) && (g_force_path == 0 || g_force_path == 4) && attn_packed_supported(L, D, H)) {
    const bool fold = all_plain && folded && dbias;
    if (fold) *folded = 1;
    attn_packed_bwd_pipeline(g_force_path == 4 ? 0 : 1);   // 4: the non-pipelined packed backward (two CTAs per SM), kept for A/B
    return launch_attn_packed<true>((const __nv_bfloat16*)qkv->hi, (const __nv_bfloat16*)qkv->lo, (const __nv_bfloat16*)dctx->hi, (const __nv_bfloat16*)dctx->lo,
                                    keymask, R, L, D, H, (__nv_bfloat16*)dqkv->hi, (__nv_bfloat16*)dqkv->lo, fold ? dbias : nullptr, make_drop(seed, site, p), st);
  }
  // 32 < L <= 128 (bert-large / seq_len 64 shapes): tcgen05 tiles with one or two sequences per tile, both directions
  if (qkv->lo == nullptr && dctx->lo == nullptr && dqkv->lo == nullptr && ((L <= 32 && g_force_path == 3) || (L > 32 && g_force_path != 1)))
    return launch_attn_umma<true>((const __nv_bfloat16*)qkv->hi, (const __nv_bfloat16*)dctx->hi, keymask, R, L, D, H, (__nv_bfloat16*)dqkv->hi,
                                  make_drop(seed, site, p), st);
  if (qkv->lo == nullptr && dctx->lo == nullptr && dqkv->lo == nullptr && L <= 32 && (g_force_path == 0 || g_force_path == 2))
    return launch_ring<true>((const __nv_bfloat16*)qkv->hi, (const __nv_bfloat16*)dctx->hi, keymask, R, L, D, H, (__nv_bfloat16*)dqkv->hi,
                             make_drop(seed, site, p), st);
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
