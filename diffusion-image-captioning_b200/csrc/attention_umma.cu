// Short-sequence multi-head attention on the tcgen05 tensor cores (plain bf16 storage, L <= 128, head dim 64).
// Replaces DistilBertSelfAttention / SDPA (HF modeling_distilbert.py:126-151,177-207) and its autograd backward for the shapes
// of the reference path (L = MAX_LENGTH + 2 = 18).
//
// Packing.  Four sequences of one head form one 128-row UMMA tile: sequence slot s owns rows [32 s, 32 s + L) (TMA zero-fills
// rows L..31 of each slot through an out-of-bounds box, so padding never has to be written or checked).  S = Q K^T is ONE
// 128 x 128 x 64 tcgen05.mma group whose four diagonal 32 x 32 blocks are the four score matrices; the off-diagonal blocks are
// wasted tensor-core work, which is free here (0.4 % of the model FLOPs) and buys the layout that matters: TMEM lane = query
// row, so each softmax thread owns one complete row of 32 scores in registers (one tcgen05.ld) - max / sum / dropout / dS need
// no shuffles, no shared-memory round trips and no fragment bookkeeping (the mma.sync version spent ~1500 (fwd) / ~2300 (bwd)
// instructions per (sequence, head); this one ~150 / ~300 per row-thread, i.e. per 32 rows of a warp).
// The probabilities are written once, as bf16, into a block-diagonal [128 x 128] shared-memory tile that then serves as
//   * K-major  A operand:  O  = P V,    dQ = dS K        (contraction over keys)
//   * MN-major A operand:  dV = P^T dO, dK = dS^T Q      (contraction over queries - the transpose is a descriptor, not a copy)
// and the q / k / v / dO tiles are likewise used both K-major (contraction over the head dim) and MN-major (contraction over rows).
//
// Longer rows (32 < L <= 128, BASELINE.json's bert-large / seq_len 64 configuration has L = 66): the slot size SL is a template
// parameter - 64 rows (two sequences per tile) or 128 rows (one sequence per tile). The MMA schedule is unchanged (P is still a
// block-diagonal [128 x 128] tile, with one or two blocks); a softmax thread then walks its row in 32-column chunks straight out of
// TMEM (max pass, sum pass, write pass - TMEM re-reads are cheap, registers stay at 32 scores per thread).
//
// One CTA = TMA producer warp + MMA-issuing warp + 4 softmax/epilogue warps (one per TMEM lane quadrant), persistent over groups;
// four (forward: 48 KB shared memory, 128 TMEM columns) or two (backward: 112 KB, 256 columns) CTAs per SM overlap one group's
// softmax with the others' loads and MMAs.
#include "common.cuh"
#include "../../include/clipdlm.h"
#include <cudaTypedefs.h>

namespace clipdlm {

int num_sms();
int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, const unsigned long long* dims, const unsigned long long* strides_bytes,
                      const uint32_t* box);

constexpr int AU_DH = 64;
constexpr int AU_THREADS = 192;          // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 softmax / epilogue
constexpr uint32_t AU_TILE = 16384;      // [128 rows][64 bf16] = 128 rows x 128 B, 128-byte swizzle
constexpr float AU_LOG2E = 1.4426950408889634f;

// same counter layout as attention.cu (attn_rand_block / attn_keep_ij): element (i, j) of pair rh draws field ((j / 8) % 4) * 2 + j % 2
// of the Philox block (rh, lane' = (i % 8) * 4 + (j % 8) / 2, q = (i / 8) * 4 + j / 32) - the SIMT and ring kernels see the same masks.
__device__ __forceinline__ uint4 au_rand_block(const DropoutCfg& d, unsigned long long rh, int lane_p, int q) {
  return philox4x32(make_uint4((uint32_t)rh, (uint32_t)(rh >> 32), (uint32_t)lane_p | ((uint32_t)q << 8), d.site ^ 0xa77e0000u),
                       make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
}
// keep-scale (0 or 1 / (1 - p)) for the 32 keys of query row i
__device__ __forceinline__ void au_row_keep(const DropoutCfg& d, unsigned long long rh, int i, int L, float (&ks)[32], int kc = 0) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // c = (j % 8) / 2; kc = j / 32 (key chunk)
    const uint4 rnd = au_rand_block(d, rh, (i & 7) * 4 + c, (i >> 3) * 4 + kc);
    const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
      if (kc * 32 + jb * 8 >= L) continue;
      ks[jb * 8 + 2 * c] = (w[jb] & 0xffffu) >= d.thresh16 ? d.scale : 0.f;
      ks[jb * 8 + 2 * c + 1] = (w[jb] >> 16) >= d.thresh16 ? d.scale : 0.f;
    }
  }
}

// keep bits (bit j = key 32 kc + j survives the dropout) of query row i: one Philox evaluation serves several passes over the row
__device__ __forceinline__ uint32_t au_row_keep_bits(const DropoutCfg& d, unsigned long long rh, int i, int L, int kc) {
  uint32_t bits = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 rnd = au_rand_block(d, rh, (i & 7) * 4 + c, (i >> 3) * 4 + kc);
    const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
      if (kc * 32 + jb * 8 >= L) continue;
      bits |= ((w[jb] & 0xffffu) >= d.thresh16 ? 1u : 0u) << (jb * 8 + 2 * c);
      bits |= ((w[jb] >> 16) >= d.thresh16 ? 1u : 0u) << (jb * 8 + 2 * c + 1);
    }
  }
  return bits;
}

// x[j] *= keep-scale, in place (forward)
__device__ __forceinline__ void au_row_drop(const DropoutCfg& d, unsigned long long rh, int i, int L, float (&x)[32], int kc = 0) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 rnd = au_rand_block(d, rh, (i & 7) * 4 + c, (i >> 3) * 4 + kc);
    const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
      if (kc * 32 + jb * 8 >= L) continue;
      x[jb * 8 + 2 * c] = (w[jb] & 0xffffu) >= d.thresh16 ? x[jb * 8 + 2 * c] * d.scale : 0.f;
      x[jb * 8 + 2 * c + 1] = (w[jb] >> 16) >= d.thresh16 ? x[jb * 8 + 2 * c + 1] * d.scale : 0.f;
    }
  }
}

struct AttUArgs {
  const uint32_t* keymask;
  __nv_bfloat16* out;       // ctx [T, D] (forward) or dqkv [T, 3D] (backward)
  int R, L, D, H;
  int kw;                   // keymask words per sequence, ceil(L / 32)
  int groups_per_head;      // ceil(R / sequences per tile)
  long long groups;         // groups_per_head * H
  DropoutCfg drop;
  float scale;
};

// 32 fp32 values of one row -> bf16, into columns [32 * slot, 32 * slot + 32) of row `row` of a [2 chunks][128 rows][128 B] K-major tile
__device__ __forceinline__ void au_store_row32(uint32_t tile, int row, int slot, const float (&x)[32]) {
  const uint32_t base = tile + (uint32_t)(slot >> 1) * AU_TILE + (uint32_t)row * 128u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t addr = base + ((((slot & 1) * 4 + c) ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(x[8 * c], x[8 * c + 1])),
                 "r"(pack_bf16x2(x[8 * c + 2], x[8 * c + 3])), "r"(pack_bf16x2(x[8 * c + 4], x[8 * c + 5])),
                 "r"(pack_bf16x2(x[8 * c + 6], x[8 * c + 7]))
                 : "memory");
  }
}
// 64 fp32 accumulator columns of this thread's TMEM lane -> 64 bf16 (128 contiguous bytes) in global memory
__device__ __forceinline__ void au_store_out64(uint32_t taddr, __nv_bfloat16* dst, bool valid, float row_scale = 1.f) {
  float v[32];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(taddr + half * 32, v);
    if (valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= row_scale;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(dst + half * 32 + c * 8) =
            make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]), pack_bf16x2(v[8 * c + 4], v[8 * c + 5]),
                       pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
    }
  }
}

// masked, scaled (base-2) scores of one 32-key chunk: s[j] <- bit j of `valid` ? s[j] * sl : -inf
// (`valid` = visible keys of the chunk, already clipped to the sequence length: one bit test per score)
__device__ __forceinline__ void au_mask_scale(float (&s)[32], uint32_t valid, float sl) {
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = (valid & (1u << j)) ? s[j] * sl : -INFINITY;
}
__device__ __forceinline__ uint32_t au_len_mask(int L, int key0) {   // bits of the keys key0 .. key0 + 31 that are < L
  const int n = L - key0;
  return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

template <bool BWD, int SL>
__global__ void __launch_bounds__(AU_THREADS, BWD ? 2 : 4) attn_umma_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                                                                  const AttUArgs a) {
  extern __shared__ __align__(1024) uint8_t au_smem[];
  // forward : Q | K | V, P (2 tiles) overlays Q | K once S is done      = 3 tiles (48 KB), 128 TMEM columns: 4 CTAs per SM
  // backward: Q | K | dO | V + P (P overlays V) | dS(2)                   = 7 tiles (112 KB), 256 TMEM columns: 2 CTAs per SM
  constexpr int OFF_Q = 0, OFF_K = 1, OFF_DO = 2, OFF_V = BWD ? 3 : 2, OFF_P = BWD ? 3 : 0, OFF_DS = 5;
  constexpr int NTILES_SMEM = BWD ? 7 : 3;
  constexpr uint32_t AU_TMEM_COLS = BWD ? 256 : 128;
  constexpr uint32_t COL_O = BWD ? 128 : 0;     // forward: O reuses the score columns (S has been consumed when P is ready)
  constexpr int NLOADS = BWD ? 4 : 3;
  constexpr int NS = 128 / SL;      // sequences per 128-row tile
  constexpr int NC = SL / 32;       // 32-key chunks per sequence
  static_assert(SL == 32 || SL == 64 || SL == 128, "slot size");
  uint8_t* smem = au_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NTILES_SMEM * AU_TILE);
  uint64_t* in_full = bars + 0; uint64_t* in_empty = bars + 1; uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3; uint64_t* o_full = bars + 4; uint64_t* t_empty = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0u) { if (threadIdx.x == 0) printf("clipdlm: attention smem base not 1024-byte aligned\n"); __trap(); }

  // backward: the block-diagonal dS tile and the upper key chunk of P keep their off-diagonal zeros for the life of the CTA
  // (the forward P tile lives where TMA keeps landing Q and K, so its rows are rewritten completely for every group)
  if (BWD) {
    uint8_t* z0 = smem + OFF_P * AU_TILE;
    const uint32_t zbytes = 4u * AU_TILE;
    for (uint32_t off = threadIdx.x * 16; off < zbytes; off += AU_THREADS * 16) *reinterpret_cast<uint4*>(z0 + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    mbar_init(in_full, 1); mbar_init(in_empty, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(t_empty, 4);
    fence_mbar_init();
    tma_prefetch_desc(&tm_qkv);
    if (BWD) tma_prefetch_desc(&tm_do);
  }
  if (warp == 1) tmem_alloc(tmem_slot, AU_TMEM_COLS);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const long long my_groups = a.groups > blockIdx.x ? (a.groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // (the whole warp walks the loop, one elect.sync-elected lane issues: no ELECT / BRA.U.ANY waterfall per TMA / tcgen05 instruction)
    for (long long n = 0; n < my_groups; ++n) {
      const long long grp = blockIdx.x + n * gridDim.x;
      const int h = (int)(grp % a.H), r0 = (int)(grp / a.H) * NS;
      mbar_wait(in_empty, ((uint32_t)n & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(in_full, NLOADS * AU_TILE);   // full boxes: rows >= L and sequences >= R arrive as zeros
        tma_load_3d(smem + OFF_Q * AU_TILE, &tm_qkv, in_full, h * AU_DH, 0, r0);
        tma_load_3d(smem + OFF_K * AU_TILE, &tm_qkv, in_full, a.D + h * AU_DH, 0, r0);
        tma_load_3d(smem + OFF_V * AU_TILE, &tm_qkv, in_full, 2 * a.D + h * AU_DH, 0, r0);
        if (BWD) tma_load_3d(smem + OFF_DO * AU_TILE, &tm_do, in_full, h * AU_DH, 0, r0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // Whole warp in the loop; each [MMAs + commits] group goes out from one elected lane (a commit sits with the MMAs it tracks).
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);    // S, dP: both operands K-major (contraction over the head dim)
      constexpr uint32_t idesc_kv = make_idesc_bf16(128, 64, 0, 1);    // O = P V, dQ = dS K: A K-major, B MN-major
      constexpr uint32_t idesc_tv = make_idesc_bf16(128, 64, 1, 1);    // dV = P^T dO, dK = dS^T Q: both MN-major
      const uint32_t sq = smem_u32(smem + OFF_Q * AU_TILE), sk = smem_u32(smem + OFF_K * AU_TILE), sv = smem_u32(smem + OFF_V * AU_TILE);
      const uint32_t sdo = smem_u32(smem + OFF_DO * AU_TILE), sp = smem_u32(smem + OFF_P * AU_TILE), sds = smem_u32(smem + OFF_DS * AU_TILE);
      for (long long n = 0; n < my_groups; ++n) {
        const uint32_t par = (uint32_t)n & 1u;
        mbar_wait(in_full, par);
        mbar_wait(t_empty, par ^ 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)   // S = Q K^T
          umma_bf16(tmem_base, make_smem_desc_sw128(sq, 16, 1024) + (uint64_t)(k * 2), make_smem_desc_sw128(sk, 16, 1024) + (uint64_t)(k * 2),
                    idesc_s, k > 0 ? 1u : 0u);
        if (BWD) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // dP = dO V^T
            umma_bf16(tmem_base + 128, make_smem_desc_sw128(sdo, 16, 1024) + (uint64_t)(k * 2),
                      make_smem_desc_sw128(sv, 16, 1024) + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_full, par);
        tc_fence_after();
        if (elect_one()) {
        if (!BWD) {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // O = P V  (keys in blocks of 16: P chunk k / 4, 32 B per step; V rows 16 k)
            umma_bf16(tmem_base + COL_O, make_smem_desc_sw128(sp + (k >> 2) * AU_TILE, 16, 1024) + (uint64_t)((k & 3) * 2),
                      make_smem_desc_sw128(sv, 8192, 1024) + (uint64_t)(k * 128), idesc_kv, k > 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dV = P^T dO   (queries in blocks of 16: rows 16 k of P and dO)
            umma_bf16(tmem_base, make_smem_desc_sw128(sp, AU_TILE, 1024) + (uint64_t)(k * 128),
                      make_smem_desc_sw128(sdo, 8192, 1024) + (uint64_t)(k * 128), idesc_tv, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dK = dS^T Q
            umma_bf16(tmem_base + 64, make_smem_desc_sw128(sds, AU_TILE, 1024) + (uint64_t)(k * 128),
                      make_smem_desc_sw128(sq, 8192, 1024) + (uint64_t)(k * 128), idesc_tv, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dQ = dS K
            umma_bf16(tmem_base + 128, make_smem_desc_sw128(sds + (k >> 2) * AU_TILE, 16, 1024) + (uint64_t)((k & 3) * 2),
                      make_smem_desc_sw128(sk, 8192, 1024) + (uint64_t)(k * 128), idesc_kv, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(in_empty);   // every operand tile of this group has been consumed
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== softmax / epilogue =====================================
    const int row = (warp & 3) * 32 + lane;   // TMEM lane (warp w may access the lane quadrant w % 4) == row of the 128-row tile
    const int slot = row / SL;                // sequence slot of the group
    const int i = row % SL;                   // query row (forward/backward) - and key row for the dK / dV outputs
    const uint32_t tq = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sp = smem_u32(smem + OFF_P * AU_TILE), sds = smem_u32(smem + OFF_DS * AU_TILE);
    for (long long n = 0; n < my_groups; ++n) {
      const uint32_t par = (uint32_t)n & 1u;
      const long long grp = blockIdx.x + n * gridDim.x;
      const int h = (int)(grp % a.H), r = (int)(grp / a.H) * NS + slot;
      const bool live = r < a.R && i < a.L;
      const unsigned long long rh = (unsigned long long)r * a.H + h;
      const bool dropping = a.drop.thresh16 != 0;
      mbar_wait(s_full, par);
      tc_fence_after();
      float out_scale = 1.f;   // SL > 32 forward: 1 / row sum, applied when the output row is stored
      if constexpr (SL > 32) {
        // ---------------- rows of up to SL keys: walk the own diagonal block in 32-column chunks ----------------
        uint32_t km[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) km[c] = (r < a.R && c < a.kw) ? (a.keymask[(size_t)r * a.kw + c] & au_len_mask(a.L, 32 * c)) : 0u;
        const uint32_t scol = tq + (uint32_t)(SL * slot);
        const float sl = a.scale * AU_LOG2E;
        float p[32];
        float mx = -INFINITY, inv = 0.f;
        // rows L..SL-1 of a slot and sequences beyond R are padding: a warp made only of such rows skips the TMEM passes and just
        // writes its zero rows (with L = 66 in a 128-row tile that is 1.9 of the 4 softmax warps). tcgen05.ld is warp-collective,
        // so the predicate is warp-uniform.
        const bool wlive = __any_sync(0xffffffffu, live);
        if (wlive) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (c * 32 >= a.L) continue;
            tmem_ld32(scol + 32 * c, p);
            au_mask_scale(p, km[c], sl);
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, p[j]);
          }
        }
        if (!BWD) {
          // P lives where Q / K were: write this row of both key chunks completely (zeros off the diagonal block).
          // The tile holds exp2(s - max), NOT yet divided by the row sum: 1 / sum is applied to the output row instead (one pass
          // over the scores fewer; the bf16 rounding then happens on values in (0, 1] whatever the row sum is).
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) {
            if (cc / (SL / 8) == slot) continue;
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sp + (uint32_t)(cc >> 3) * AU_TILE + (uint32_t)row * 128u + (uint32_t)(((cc & 7) ^ (row & 7)) << 4)),
                         "r"(0u) : "memory");
          }
          float sum = 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (wlive && c * 32 < a.L) {
              tmem_ld32(scol + 32 * c, p);
              au_mask_scale(p, km[c], sl);
#pragma unroll
              for (int j = 0; j < 32; ++j) { p[j] = live ? ex2_ftz(p[j] - mx) : 0.f; sum += p[j]; }   // a fully masked row gives NaN, as the reference does
              if (dropping) au_row_drop(a.drop, rh, i, a.L, p, c);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) p[j] = 0.f;
            }
            au_store_row32(sp, row, NC * slot + c, p);
          }
          inv = 1.f / sum;
          out_scale = inv;
        } else {
          float dp[32];
          uint32_t kb[NC];   // dropout keep bits per key chunk, drawn once (first pass) and reused when dS is written
          const uint32_t dcol = tq + 128u + (uint32_t)(SL * slot);
          float sum = 0.f, edot = 0.f;   // sum_j e_j and sum_j e_j dP_j (through the dropout): dot = edot / sum
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            kb[c] = 0xffffffffu;
            if (!wlive || c * 32 >= a.L) continue;
            tmem_ld32(scol + 32 * c, p);
            tmem_ld32(dcol + 32 * c, dp);
            au_mask_scale(p, km[c], sl);
            if (dropping) kb[c] = au_row_keep_bits(a.drop, rh, i, a.L, c);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = ex2_ftz(p[j] - mx);
              sum += e;
              const float kj = (kb[c] & (1u << j)) ? a.drop.scale : 0.f;
              edot = fmaf(dropping ? dp[j] * kj : dp[j], e, edot);
            }
          }
          inv = 1.f / sum;
          const float dot = live ? edot * inv : 0.f;
          // V's tile doubles as key chunk 0 of P and has just been clobbered by TMA: rewrite the part of this row outside the own block
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            if (cc / (SL / 8) == slot) continue;
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sp + (uint32_t)row * 128u + (uint32_t)((cc ^ (row & 7)) << 4)), "r"(0u) : "memory");
          }
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (wlive && c * 32 < a.L) {
              tmem_ld32(scol + 32 * c, p);
              tmem_ld32(dcol + 32 * c, dp);
              au_mask_scale(p, km[c], sl);
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float pj = live ? ex2_ftz(p[j] - mx) * inv : 0.f;
                const float kj = (kb[c] & (1u << j)) ? a.drop.scale : 0.f;
                const float dpj = dropping ? dp[j] * kj : dp[j];        // gradient through the dropout
                dp[j] = pj * (dpj - dot) * a.scale;                     // dS
                p[j] = dropping ? pj * kj : pj;                         // dropped probabilities: what multiplied V in the forward
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) { p[j] = 0.f; dp[j] = 0.f; }
            }
            au_store_row32(sp, row, NC * slot + c, p);
            au_store_row32(sds, row, NC * slot + c, dp);
          }
        }
      } else {
      const uint32_t keybits = r < a.R ? (a.keymask[r] & au_len_mask(a.L, 0)) : 0u;
      float p[32];
      tmem_ld32(tq + 32 * slot, p);       // this row's scores against the 32 key slots of its own sequence (diagonal block)
      {
        const float sl = a.scale * AU_LOG2E;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = (keybits & (1u << j)) ? p[j] * sl : -INFINITY;
          mx = fmaxf(mx, p[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { p[j] = ex2_ftz(p[j] - mx); sum += p[j]; }   // a fully masked row gives NaN, as the reference does
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 32; ++j) p[j] = live ? p[j] * inv : 0.f;
      }
      if (!BWD) {
        if (dropping) au_row_drop(a.drop, rh, i, a.L, p);
        // P lives where Q / K were: write this row of both key chunks completely (zeros off the diagonal block)
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          if ((cc >> 2) == slot) continue;
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sp + (uint32_t)(cc >> 3) * AU_TILE + (uint32_t)row * 128u + (uint32_t)(((cc & 7) ^ (row & 7)) << 4)),
                       "r"(0u) : "memory");
        }
        au_store_row32(sp, row, slot, p);
      } else {
        float ks[32];
        if (dropping) {
#pragma unroll
          for (int j = 0; j < 32; ++j) ks[j] = 0.f;
          au_row_keep(a.drop, rh, i, a.L, ks);
        }
        float dp[32];
        tmem_ld32(tq + 128 + 32 * slot, dp);
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (dropping) dp[j] *= ks[j];       // gradient through the dropout
          dot = fmaf(dp[j], p[j], dot);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) dp[j] = p[j] * (dp[j] - dot) * a.scale;   // dS (p == 0 on dead rows / masked keys)
        if (dropping) {
#pragma unroll
          for (int j = 0; j < 32; ++j) p[j] *= ks[j];   // dropped probabilities: what multiplied V in the forward
        }
        // V's tile doubles as the first half (key chunk 0) of P and has just been clobbered by TMA: rewrite this row of it completely
        if (slot >= 2) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sp + (uint32_t)row * 128u + (uint32_t)(c << 4)), "r"(0u) : "memory");
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sp + (uint32_t)row * 128u + (uint32_t)(((((slot ^ 1) & 1) * 4 + c) ^ (row & 7)) << 4)),
                         "r"(0u) : "memory");
        }
        au_store_row32(sp, row, slot, p);
        au_store_row32(sds, row, slot, dp);
      }
      }  // SL == 32
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P / dS (generic proxy) -> tcgen05.mma operand reads (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      mbar_wait(o_full, par);
      tc_fence_after();
      if (!BWD) {
        au_store_out64(tq + COL_O, a.out + ((size_t)r * a.L + i) * a.D + h * AU_DH, live, SL > 32 ? out_scale : 1.f);
      } else {
        __nv_bfloat16* o = a.out + ((size_t)r * a.L + i) * 3 * a.D + h * AU_DH;
        au_store_out64(tq + 128, o, live);               // dQ
        au_store_out64(tq + 64, o + a.D, live);          // dK (row = key i)
        au_store_out64(tq, o + 2 * a.D, live);           // dV
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, AU_TMEM_COLS);
}

template <bool BWD, int SL>
static int launch_attn_umma_sl(const __nv_bfloat16* qkv, const __nv_bfloat16* dctx, const uint32_t* keymask, int R, int L, int D, int H,
                               __nv_bfloat16* out, const DropoutCfg& drop, cudaStream_t st) {
  constexpr int NS = 128 / SL;
  CUtensorMap tm_qkv, tm_do;
  int rc;
  {
    const unsigned long long dims[3] = {3ull * D, (unsigned long long)L, (unsigned long long)R};
    const unsigned long long str[2] = {3ull * D * 2, 3ull * D * 2 * L};
    const uint32_t box[3] = {AU_DH, SL, NS};
    if ((rc = make_tmap_3d_bf16(&tm_qkv, qkv, dims, str, box))) return rc;
  }
  tm_do = tm_qkv;
  if (BWD) {
    const unsigned long long dims[3] = {(unsigned long long)D, (unsigned long long)L, (unsigned long long)R};
    const unsigned long long str[2] = {(unsigned long long)D * 2, (unsigned long long)D * 2 * L};
    const uint32_t box[3] = {AU_DH, SL, NS};
    if ((rc = make_tmap_3d_bf16(&tm_do, dctx, dims, str, box))) return rc;
  }
  AttUArgs a;
  a.keymask = keymask; a.out = out; a.R = R; a.L = L; a.D = D; a.H = H; a.drop = drop; a.scale = 0.125f;  // 1 / sqrt(64)
  a.kw = (L + 31) / 32;
  a.groups_per_head = (R + NS - 1) / NS;
  a.groups = (long long)a.groups_per_head * H;
  const size_t smem = (size_t)(BWD ? 7 : 3) * AU_TILE + 64;
  static bool attr_set = false;
  if (!attr_set) {
    CLIPDLM_CUDA_OK(cudaFuncSetAttribute(attn_umma_kernel<BWD, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const long long want = (BWD ? 2LL : 4LL) * num_sms();
  const int grid = (int)(a.groups < want ? a.groups : want);
  CLIPDLM_CUDA_OK(launch_pdl(attn_umma_kernel<BWD, SL>, dim3(grid), dim3(AU_THREADS), smem, st, tm_qkv, tm_do, a));
  return 0;
}

// L <= 32: four sequences per tile; L <= 64: two; L <= 128: one.
template <bool BWD>
int launch_attn_umma(const __nv_bfloat16* qkv, const __nv_bfloat16* dctx, const uint32_t* keymask, int R, int L, int D, int H,
                     __nv_bfloat16* out, const DropoutCfg& drop, cudaStream_t st) {
  CLIPDLM_CHECK(L >= 1 && L <= 128, "tcgen05 attention: L %d out of range (1..128)", L);
  if (L <= 32) return launch_attn_umma_sl<BWD, 32>(qkv, dctx, keymask, R, L, D, H, out, drop, st);
  if (L <= 64) return launch_attn_umma_sl<BWD, 64>(qkv, dctx, keymask, R, L, D, H, out, drop, st);
  return launch_attn_umma_sl<BWD, 128>(qkv, dctx, keymask, R, L, D, H, out, drop, st);
}
template int launch_attn_umma<false>(const __nv_bfloat16*, const __nv_bfloat16*, const uint32_t*, int, int, int, int, __nv_bfloat16*, const DropoutCfg&,
                                     cudaStream_t);
template int launch_attn_umma<true>(const __nv_bfloat16*, const __nv_bfloat16*, const uint32_t*, int, int, int, int, __nv_bfloat16*, const DropoutCfg&,
                                    cudaStream_t);

}  // namespace clipdlm
