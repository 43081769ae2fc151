// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues.  D[M,N] = A[M,K] * B[N,K]^T.
//
// One persistent, warp-specialised kernel (384 threads, 1 CTA / SM), run as CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles)
// or single CTAs (cta_group::1, 128 x 256 tiles; see Geo<CG>):
//   warp 0   : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, 4 stages x (A 16 KB + B 32 KB))
//   warp 1   : MMA issuer    (one lane issues tcgen05.mma kind::f16, 128x256x16, accumulators in TMEM)
//   warp 2   : TMEM allocator (512 columns = 2 accumulator stages of 256 fp32 columns)
//   warps 4-7: epilogue      (tcgen05.ld -> registers -> fused math -> smem transpose -> coalesced global I/O)
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue), static persistent tile loop.
// Operand majors: K-major (fwd), K-major x MN-major (dgrad: B = W[N_red][K_out]), MN x MN (wgrad: reduction over tokens).
// Split precision (bf16x3): the k-loop is run over up to three (A part, B part) pairs hi*hi, lo*hi, hi*lo into the
// same TMEM accumulator, giving fp32-class products from the bf16 tensor pipe.
//
// Replaces: every nn.Linear forward/backward GEMM on the reference hot path (see include/clipdlm.h).
#include "common.cuh"
#include "../../include/clipdlm.h"
#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>

namespace clipdlm {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;          // 16384
constexpr int EPI_WARPS = 8;                        // two epilogue warps per TMEM lane quadrant, each owning 128 of the 256 tile columns
// (Round 2 measured a 16-warp variant of the math-heavy K = 768 epilogues - four warps per quadrant, 64 columns each, 16-column TMEM pieces,
// 64B-swizzled 32-column TMA-store boxes, 92 registers: parity-green but SLOWER, STORE_GELU_DERIV 0.713 vs 0.691 ms, STORE_MULAUX 0.694 vs
// 0.599 ms per 8192-row chunk. Both layouts issue ~0.5 instructions per cycle and scheduler: FFMA2 / FMNMX / F2FP / LOP3 take two dispatch
// cycles each and MUFU.EX2 eight, so ~15 instructions per element cost about what the 6144 MMA cycles of a tile offer whatever the warp
// count, and the narrower stores and auxiliary loads cost more than the extra latency hiding bought. Removed)
constexpr int STG_PITCH = 80;                       // fp32 staging (WGRAD): bytes per staged row (64 B payload = 16 fp32, + 16 B pad)
constexpr int STG_WARP_BYTES = 8192;                // per epilogue warp: two [32 rows][128 B] swizzled tiles feeding TMA stores (or the fp32 staging)
constexpr int TMEM_COLS = 512;
// Geometry per CTA-group size.  CG = 1: one CTA computes a 128 x 256 tile (tcgen05.mma.cta_group::1, M = 128).
// CG = 2: a CTA pair (cluster of two SMs of one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2 (M = 256): each CTA
// stages its own 128 rows of A and only HALF of the B tile (128 of the 256 N rows), so the shared-memory traffic per MMA
// (TMA writes + tensor-core operand reads) drops by a third - the CG = 1 kernel is shared-memory-bandwidth bound at ~65 % of
// the tensor pipe - and the freed shared memory deepens the ring (5 stages of 32 KB per CTA next to the 64 KB of TMA-store staging).
template <int CG>
struct Geo {
  static constexpr int BN_CTA = BN / CG;
  static constexpr int B_STAGE_BYTES = BN_CTA * BK * 2;   // 32768 / 16384
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = CG == 2 ? 5 : 3;   // + 64 KB of epilogue staging = 227 KB
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_WARPS * STG_WARP_BYTES + BN * 4 /*bias*/ + 256 /*barriers*/;
};
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. epilogue

struct GemmArgs {
  int M, N, K;
  int num_m_tiles, num_n_tiles, k_splits, kb_per_split, kb_total;
  int nparts;
  int part_a[3], part_b[3];  // 0 = hi map, 1 = lo map
  int gather_len;            // A gather (3-D tensor map) if > 0
  uint32_t mn_lbo, mn_sbo;   // MN-major descriptor strides (bytes)
  // epilogue
  __nv_bfloat16 *out_hi, *out_lo, *out2_hi, *out2_lo;
  float* out_f32;
  long long ldo;
  const float* bias;
  const __nv_bfloat16 *res_hi, *res_lo;
  long long ldr;
  const __nv_bfloat16 *u_hi, *u_lo;
  long long ldu;
  int scatter_len, scatter_stride;
  int al32;  // every bf16 epilogue operand is 32-byte aligned with a pitch that is a multiple of 16 elements
  uint32_t dbg;  // clipdlm_gemm_debug_flags
  int tma_out;   // the bf16 outputs of the specialised STORE epilogue / the LSE logits go out through TMA stores (tmO / tmO2)
  int fast_mode; // >= 0: specialised STORE epilogue (bit 0 bias, bits 1-2 aux 0 none / 1 residual / 2 gelu'(u), bit 3 gelu dual store, bit 4 dropout)
  DropoutCfg drop;
  float* part_max;
  float* part_sum;
  int* part_arg;
  float* tgt_logit;
  const int* targets;
  int tgt_period;
  const float* lse;          // SMGRAD: lse[M];  LSE_EXP: the exponent shift (device scalar, may be NULL);  STORE_ROWSCALE: row_scale[M]
  float grad_scale;          // (one slot for the three: GemmArgs keeps its layout, so the other instantiations compile to the same SASS)
  uint64_t pol_a, pol_b;     // L2 cache-policy operands of the plain 2-D TMA loads of A / B (0 = default policy)
  int band;                  // > 0: tile order in bands of `band` row tiles, column tile outer / row tile inner within a band (tile_mn)
};

// Work item -> (row tile, column tile). Default: column tile fastest - the CTA pairs that run concurrently share a few A row tiles and
// sweep all of B, right when B (a weight matrix) is small. Band order (lm_head passes: 120 column tiles of a 47 MB table, gigabytes of
// output streaming through L2): the `band` pairs that run concurrently take `band` DIFFERENT row tiles of the SAME column tile, step by step -
// a B tile is fetched once per band (7 x 47 MB per pass) and the band's A tiles (29 MB) stay L2-resident over the sweep, instead of B being
// re-fetched whenever the output stream has pushed it out.
__device__ __forceinline__ void tile_mn(const GemmArgs& g, int tile, int& m_blk, int& n_blk) {
  if (g.band <= 0) { n_blk = tile % g.num_n_tiles; m_blk = tile / g.num_n_tiles; return; }
  const int per_band = g.band * g.num_n_tiles;
  const int b = tile / per_band, r = tile - b * per_band;
  const int rows = min(g.band, g.num_m_tiles - b * g.band);
  n_blk = r / rows;
  m_blk = b * g.band + (r - n_blk * rows);
}

__device__ __forceinline__ long long map_row(const GemmArgs& g, int m) {
  return g.scatter_len > 0 ? (long long)(m / g.scatter_len) * g.scatter_stride + (m % g.scatter_len) : (long long)m;
}

// ---- row-wise epilogue I/O --------------------------------------------------------------------------------------
// After tcgen05.ld (32x32b) thread i of a TMEM lane quadrant owns row i of the tile and 32 consecutive columns, i.e. 64
// contiguous bytes of a bf16 row. Those are read / written directly with two 256-bit accesses (LDG/STG.256: one full 32-byte
// sector per access) - no shared-memory transpose, no extra synchronisation. Falls back to 128-bit accesses when the
// operand is only 16-byte aligned.
__device__ __forceinline__ void ld_row64(const __nv_bfloat16* p, bool al32, uint32_t (&r)[16]) {
  if (al32) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[h * 8 + 0]), "=r"(r[h * 8 + 1]), "=r"(r[h * 8 + 2]), "=r"(r[h * 8 + 3]), "=r"(r[h * 8 + 4]), "=r"(r[h * 8 + 5]),
                     "=r"(r[h * 8 + 6]), "=r"(r[h * 8 + 7])
                   : "l"(p + h * 16));
  } else {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const uint4 x = *reinterpret_cast<const uint4*>(p + h * 8);
      r[h * 4 + 0] = x.x; r[h * 4 + 1] = x.y; r[h * 4 + 2] = x.z; r[h * 4 + 3] = x.w;
    }
  }
}
__device__ __forceinline__ void st_row64(__nv_bfloat16* p, bool al32, const uint32_t (&r)[16]) {
  if (al32) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + h * 16), "r"(r[h * 8 + 0]), "r"(r[h * 8 + 1]),
                   "r"(r[h * 8 + 2]), "r"(r[h * 8 + 3]), "r"(r[h * 8 + 4]), "r"(r[h * 8 + 5]), "r"(r[h * 8 + 6]), "r"(r[h * 8 + 7])
                   : "memory");
  } else {
#pragma unroll
    for (int h = 0; h < 4; ++h) *reinterpret_cast<uint4*>(p + h * 8) = make_uint4(r[h * 4 + 0], r[h * 4 + 1], r[h * 4 + 2], r[h * 4 + 3]);
  }
}
// v (+)= unpack(raw)
template <bool ACC>
__device__ __forceinline__ void row_unpack(const uint32_t (&raw)[16], float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 f = unpack_bf16x2(raw[i]);
    if (ACC) { v[2 * i] += f.x; v[2 * i + 1] += f.y; } else { v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
}
// 32 values of a bf16 (pair) row segment -> floats
__device__ __forceinline__ void row_load_pair(const __nv_bfloat16* hi, const __nv_bfloat16* lo, bool al32, float (&r)[32]) {
  uint32_t raw[16];
  ld_row64(hi, al32, raw);
  row_unpack<false>(raw, r);
  if (lo != nullptr) {
    ld_row64(lo, al32, raw);
    row_unpack<true>(raw, r);
  }
}
__device__ __forceinline__ void row_store_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, bool al32, const float (&v)[32]) {
  uint32_t raw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) raw[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  st_row64(hi, al32, raw);
  if (lo != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i) raw[i] = pack_bf16x2(v[2 * i] - bf16_round(v[2 * i]), v[2 * i + 1] - bf16_round(v[2 * i + 1]));
    st_row64(lo, al32, raw);
  }
}
// fp32 tile 32 rows x 32 cols, staged as two 16-column halves (64 B per row). Coalesced side: lane l handles row (l>>2)+8j, piece (l&3).
template <bool RED>
__device__ __forceinline__ void tile_store_f32(const GemmArgs& g, float* base, long long ld, int row_base, int n0, uint8_t* stg,
                                               const float (&v)[32]) {
  const int lane = lane_id();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(stg + lane * STG_PITCH + i * 16) =
          make_float4(v[h * 16 + i * 4], v[h * 16 + i * 4 + 1], v[h * 16 + i * 4 + 2], v[h * 16 + i * 4 + 3]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = (lane >> 2) + 8 * j;
      const int m = row_base + rr;
      if (m < g.M) {
        const float4 x = *reinterpret_cast<const float4*>(stg + rr * STG_PITCH + (lane & 3) * 16);
        float* dst = base + map_row(g, m) * ld + n0 + h * 16 + (lane & 3) * 4;
        if (RED) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
        } else {
          *reinterpret_cast<float4*>(dst) = x;
        }
      }
    }
    __syncwarp();
  }
}

// ---- specialised STORE epilogue ------------------------------------------------------------------------------------
// The generic epilogue below decides everything per 32-column chunk at run time (which outputs exist, pair or plain storage,
// alignment, tails): ~11 bookkeeping instructions per element next to ~12 useful ones in the GELU epilogues, and the K = 768
// GEMMs are epilogue-issue bound.  The engine's hot launches all fall into a handful of shapes (plain bf16, 32-byte aligned
// rows, N a multiple of 256), so those get straight-line code: the four chunks of a warp are fully unrolled, the auxiliary
// operand (residual or gelu' input) is double-buffered in registers one chunk ahead, no per-chunk branches remain.
// 32 bf16 (one chunk of this thread's row) into the warp's [32 rows][128 B] swizzled staging tile, half `half` of the 64-column pair
__device__ __forceinline__ void stage_row32(uint32_t buf, int row, int half, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t addr = buf + (uint32_t)row * 128u + ((uint32_t)((half * 4 + q) ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[8 * q], v[8 * q + 1])),
                 "r"(pack_bf16x2(v[8 * q + 2], v[8 * q + 3])), "r"(pack_bf16x2(v[8 * q + 4], v[8 * q + 5])),
                 "r"(pack_bf16x2(v[8 * q + 6], v[8 * q + 7]))
                 : "memory");
  }
}

// GD: 0 = none; 1 = forward of a GELU layer whose backward is to be a plain multiply: out = gelu'(acc + bias), out2 = gelu(acc + bias)
// (STORE_GELU_DERIV); 2 = that backward: out = acc * aux, aux = the stored gelu'(u) (STORE_MULAUX).
template <bool BIAS, int AUX, bool DUAL, bool DROP, bool RSCALE = false, int GD = 0>
__device__ __forceinline__ void epi_store_fast(const GemmArgs& g, uint32_t taddr, int col0, int m, bool valid, long long mr, uint32_t sbias_u32,
                                               uint64_t* tfull, uint32_t tphase, const CUtensorMap* tmO, const CUtensorMap* tmO2, uint32_t stg,
                                               int row_base) {
  const __nv_bfloat16* aux = AUX == 2 ? g.u_hi : g.res_hi;
  const long long ld_aux = AUX == 2 ? g.ldu : g.ldr;
  const __nv_bfloat16* aux_row = AUX != 0 ? aux + mr * ld_aux + col0 : nullptr;
  __nv_bfloat16* out_row = g.out_hi != nullptr ? g.out_hi + mr * g.ldo + col0 : nullptr;
  __nv_bfloat16* out2_row = DUAL ? g.out2_hi + mr * g.ldo + col0 : nullptr;
  uint32_t aux0[16], aux1[16];
  if (AUX != 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { aux0[i] = 0u; aux1[i] = 0u; }
    if (valid) ld_row64(aux_row, true, aux0);      // first chunk: in flight while the MMAs of this tile still run
  }
  float rs = 1.f;
  if (RSCALE) rs = valid ? g.lse[m] : 0.f;   // STORE_ROWSCALE: this thread's accumulator row is scaled before the residual is added
  mbar_wait(tfull, tphase);
  tc_fence_after();
  // one 32-column chunk; cur holds its auxiliary operand, nxt receives the next chunk's
  auto chunk = [&](int cc, uint32_t (&cur)[16], uint32_t (&nxt)[16]) {
    float v[32];
    tmem_ld32(taddr + cc * 32, v);
    if (AUX != 0 && cc + 1 < 4 && valid) ld_row64(aux_row + (cc + 1) * 32, true, nxt);
    if (RSCALE) {
      const f32x2 rs2 = pk2(rs, rs);
#pragma unroll
      for (int j = 0; j < 16; ++j) upk2(mul2(pk2(v[2 * j], v[2 * j + 1]), rs2), v[2 * j], v[2 * j + 1]);
    }
    if (BIAS) {
      // bias of this chunk's 32 columns: from the per-tile shared array (default), or - experiment, host debug bit 13 - straight from global
      // memory: every lane reads the same 16 bytes (one L1 wavefront per load), so the per-tile fill and its two epilogue-wide bar.syncs go
      // away (ncu source view: 8 % of the epilogue warps' samples on the lin1 forward GEMM sat at those barriers; -1.5 to -3.5 % per launch).
      // Never ld.global.nc: the bias is a trainable parameter, and the non-coherent path served stale values (gemm_dispatch).
      const uint32_t sb = sbias_u32 + cc * 128;
      const float* gb = g.bias + col0 + cc * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b4;
        if (g.dbg & 8192u) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(sb + j * 16));
        else asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "l"(gb + j * 4) : "memory");
        upk2(add2(pk2(v[4 * j], v[4 * j + 1]), pk2(b4.x, b4.y)), v[4 * j], v[4 * j + 1]);
        upk2(add2(pk2(v[4 * j + 2], v[4 * j + 3]), pk2(b4.z, b4.w)), v[4 * j + 2], v[4 * j + 3]);
      }
    }
    if (DROP) {
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        const unsigned long long idx = (unsigned long long)m * (unsigned long long)g.N + (unsigned long long)(col0 + cc * 32 + j8 * 8);
        const uint32_t keep = dropout_keep8(g.drop, idx >> 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j8 * 8 + i] = ((keep >> i) & 1u) ? v[j8 * 8 + i] * g.drop.scale : 0.f;
      }
    }
    if (AUX == 2 && GD == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) upk2(mul2(pk2(v[2 * j], v[2 * j + 1]), bf2_to_f2(cur[j])), v[2 * j], v[2 * j + 1]);
    } else if (AUX == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const f32x2 r = mul2(pk2(v[2 * j], v[2 * j + 1]), dgelu2(__uint_as_float(cur[j] << 16), __uint_as_float(cur[j] & 0xffff0000u)));
        upk2(r, v[2 * j], v[2 * j + 1]);
      }
    } else if (AUX == 1) {
      row_unpack<true>(cur, v);
    }
    float dv[32];   // GD == 1: gelu'(pre-activation) for the first output; v becomes gelu(pre-activation) for the second
    if (GD == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f32x2 gg, dd;
        gelu_dgelu2(v[2 * j], v[2 * j + 1], gg, dd);
        upk2(dd, dv[2 * j], dv[2 * j + 1]);
        upk2(gg, v[2 * j], v[2 * j + 1]);
      }
    }
    if (g.tma_out) {
      // Outputs leave through shared memory + TMA: the row-wise 32-byte stores cost one L1 transaction per sector (2 K per
      // 128 x 256 output tile, twice that for the dual store) and made the GELU GEMMs store-bound. Buffers alternate (single
      // output: pair 0 -> A, pair 1 -> B; dual: u -> A, gelu(u) -> B), a buffer is rewritten only after wait_group.read.
      const int lane = m - row_base;
      const bool single = !DUAL || out_row == nullptr;   // one output stream (also: gelu-only inference variant of the dual mode)
      const uint32_t bufU = stg + (single ? (uint32_t)((cc >> 1) * 4096) : 0u), bufG = stg + (single ? (uint32_t)((cc >> 1) * 4096) : 4096u);
      // (the TMA stores stay with lane 0: bulk async-groups are per thread, and the variant that lets elect.sync pick the issuing lane needs EVERY
      //  lane to commit / wait - measured +7 % on the LSE_EXP pass and -0.8 % on the denoise loop, same box)
      if ((cc & 1) == 0) {
        if (lane == 0) { if (single) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
        __syncwarp();
      }
      if (!DUAL || out_row != nullptr) stage_row32(bufU, lane, cc & 1, GD == 1 ? dv : v);
      if (DUAL) {
        if (GD != 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) upk2(gelu2(v[2 * j], v[2 * j + 1]), v[2 * j], v[2 * j + 1]);
        }
        stage_row32(bufG, lane, cc & 1, v);
      }
      if (cc & 1) {
        fence_async_smem();
        __syncwarp();
        if (GD == 2 && g.part_max != nullptr) {
          // STORE_MULAUX with a fused bias gradient: column sums of this warp's staged [32 rows][64 columns] bf16 tile (the values the
          // TMA store below writes; rows past M are exact zeros), one fp32 red.add pair per lane. Lane l owns columns 2l, 2l + 1: the
          // 16-byte chunk (l >> 2) of every row, swizzled like stage_row32 - the 32 lanes of a load hit 32 different banks.
          float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w)
                         : "r"(bufU + (uint32_t)r * 128u + ((uint32_t)((lane >> 2) ^ (r & 7)) << 4) + (uint32_t)(lane & 3) * 4u));
            s0 += __uint_as_float(w << 16);
            s1 += __uint_as_float(w & 0xffff0000u);
          }
          asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(g.part_max + col0 + (cc >> 1) * 64 + 2 * lane), "f"(s0), "f"(s1) : "memory");
        }
        if (lane == 0) {
          const int c0 = col0 + (cc >> 1) * 64;
          if (!DUAL || out_row != nullptr) { tma_store_2d(tmO, bufU, c0, row_base); bulk_commit(); }
          if (DUAL) { tma_store_2d(tmO2, bufG, c0, row_base); bulk_commit(); }
        }
      }
    } else if (valid) {
      if ((!DUAL || out_row != nullptr) && !(g.dbg & 256u)) row_store_pair(out_row + cc * 32, nullptr, true, GD == 1 ? dv : v);
      if (DUAL) {
        if (GD != 1 && !(g.dbg & 128u)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) upk2(gelu2(v[2 * j], v[2 * j + 1]), v[2 * j], v[2 * j + 1]);
        }
        row_store_pair(out2_row + cc * 32, nullptr, true, v);
      }
    }
  };
  // Math-heavy bodies (GELU, gelu', Philox) stay rolled: four unrolled copies (~25 KB of SASS per mode) thrash the instruction
  // cache and measured slower than the generic loop; the light ones are fully unrolled.
  constexpr bool HEAVY = DUAL || DROP || (AUX == 2 && GD != 2);
  if (HEAVY) {
#pragma unroll 1
    for (int c2 = 0; c2 < 4; c2 += 2) {
      chunk(c2, aux0, aux1);
      chunk(c2 + 1, aux1, aux0);
    }
  } else {
    chunk(0, aux0, aux1);
    chunk(1, aux1, aux0);
    chunk(2, aux0, aux1);
    chunk(3, aux1, aux0);
  }
}

template <int AMAJ, int BMAJ, int EPI, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
            const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2, const GemmArgs g) {
  using G = Geo<CG>;
  constexpr int STAGES = G::STAGES, STAGE_BYTES = G::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  float* sbias = reinterpret_cast<float*>(staging + EPI_WARPS * STG_WARP_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sbias) + BN * 4);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = g.num_m_tiles * g.num_n_tiles * g.k_splits;
  // work units: CTAs (CG = 1) or CTA pairs (CG = 2); both CTAs of a pair walk the same item list in lockstep
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit0 = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_stride = CG == 2 ? (int)num_clusters_x() : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    pdl_launch_dependents();
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB0);
    if (g.nparts > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
  }
  if (warp == 1 && lane == 0) {
    // full: the leader's barrier collects the leader's arrive.expect_tx (bytes of BOTH CTAs) + the peer's plain arrive
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], EPI_WARPS * CG); }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_cg2(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // barrier inits + TMEM allocation visible (pair-wide for CG = 2)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; from here on we touch global memory

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // (whole warp in the loop, one elected lane issues: see the MMA issuer below)
    {
      int stage = 0; uint32_t phase = 0;
      const uint32_t full0_cluster = CG == 2 ? mapa_u32(&full_bar[0], 0) : 0u;  // leader's full barriers (shared::cluster)
      for (int w = unit0; w < total_items; w += unit_stride) {
        const int split = w % g.k_splits;
        const int tile = w / g.k_splits;
        int n_blk, m_blk;
        tile_mn(g, tile, m_blk, n_blk);
        const int m0 = (m_blk * CG + cta_rank) * BM;           // first A row of this CTA
        const int n0 = n_blk * BN + cta_rank * G::BN_CTA;      // first B row this CTA stages
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(kb0 + g.kb_per_split, g.kb_total);
        for (int part = 0; part < g.nparts; ++part) {
          const CUtensorMap* ta = g.part_a[part] ? &tmA1 : &tmA0;
          const CUtensorMap* tb = g.part_b[part] ? &tmB1 : &tmB0;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stages + stage * STAGE_BYTES;
            uint8_t* sb = sa + A_STAGE_BYTES;
            if (!elect_one()) {
            } else if (g.dbg & 32u) {   // triage: no loads at all, the MMAs run on whatever is in shared memory
              if (CG == 1 || cta_rank == 0) mbar_arrive(&full_bar[stage]);
              else mbar_arrive_cluster(full0_cluster + stage * 8);
            } else if (CG == 1) {
              mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
              if (AMAJ == 0) {
                if (g.gather_len > 0) tma_load_3d(sa, ta, &full_bar[stage], kb * BK, 0, m0 / g.gather_len);
                else if (g.pol_a != 0) tma_load_2d_hint(sa, ta, &full_bar[stage], kb * BK, m0, g.pol_a);
                else tma_load_2d(sa, ta, &full_bar[stage], kb * BK, m0);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, ta, &full_bar[stage], m0 + j * 64, kb * BK);
              }
              if (BMAJ == 0) {
                if (g.pol_b != 0) tma_load_2d_hint(sb, tb, &full_bar[stage], kb * BK, n0, g.pol_b);
                else tma_load_2d(sb, tb, &full_bar[stage], kb * BK, n0);
              } else if (g.pol_b != 0) {
#pragma unroll
                for (int j = 0; j < G::BN_CTA / 64; ++j) tma_load_2d_hint(sb + j * 8192, tb, &full_bar[stage], n0 + j * 64, kb * BK, g.pol_b);
              } else {
#pragma unroll
                for (int j = 0; j < G::BN_CTA / 64; ++j) tma_load_2d(sb + j * 8192, tb, &full_bar[stage], n0 + j * 64, kb * BK);
              }
            } else {
              const uint32_t fb = full0_cluster + stage * 8;
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
              else mbar_arrive_cluster(fb);
              if (AMAJ == 0) {
                if (g.gather_len > 0) tma_load_3d_cg2(sa, ta, fb, kb * BK, 0, m0 / g.gather_len);
                else if (g.pol_a != 0) tma_load_2d_cg2_hint(sa, ta, fb, kb * BK, m0, g.pol_a);
                else tma_load_2d_cg2(sa, ta, fb, kb * BK, m0);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(sa + j * 8192, ta, fb, m0 + j * 64, kb * BK);
              }
              if (BMAJ == 0) {
                if (g.pol_b != 0) tma_load_2d_cg2_hint(sb, tb, fb, kb * BK, n0, g.pol_b);
                else tma_load_2d_cg2(sb, tb, fb, kb * BK, n0);
              } else if (g.pol_b != 0) {
#pragma unroll
                for (int j = 0; j < G::BN_CTA / 64; ++j) tma_load_2d_cg2_hint(sb + j * 8192, tb, fb, n0 + j * 64, kb * BK, g.pol_b);
              } else {
#pragma unroll
                for (int j = 0; j < G::BN_CTA / 64; ++j) tma_load_2d_cg2(sb + j * 8192, tb, fb, n0 + j * 64, kb * BK);
              }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA of the pair only) =====================================
    // The WHOLE warp walks the loop and lane 0 issues: under an enclosing `if (lane == 0)` every descriptor lives in vector registers and each
    // tcgen05 instruction is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall - ~130 dependent instructions per 64-wide k-block
    // against the 512 cycles its four MMAs take (ncu source view, round 2: the issuer's samples are flat over that loop, and with four epilogue
    // warps on its scheduler it, not the epilogue, bounded the K = 768 GEMMs). Uniform control flow keeps the loop state on the uniform datapath.
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, AMAJ, BMAJ);
      constexpr uint32_t a_step = (AMAJ == 0 ? UMMA_K * 2 : UMMA_K * 128) >> 4;  // K advance per MMA, 16-byte units
      constexpr uint32_t b_step = (BMAJ == 0 ? UMMA_K * 2 : UMMA_K * 128) >> 4;
      const uint32_t stages_u32 = smem_u32(stages);
      const uint32_t a_lbo = AMAJ == 0 ? 16u : g.mn_lbo, a_sbo = AMAJ == 0 ? 1024u : g.mn_sbo;
      const uint32_t b_lbo = BMAJ == 0 ? 16u : g.mn_lbo, b_sbo = BMAJ == 0 ? 1024u : g.mn_sbo;
      const uint64_t adesc0 = make_smem_desc_sw128(0u, a_lbo, a_sbo), bdesc0 = make_smem_desc_sw128(0u, b_lbo, b_sbo);   // start address added per stage
      const bool no_mma = (g.dbg & 16u) != 0;   // triage: no MMAs, the commits below fire immediately
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int w = unit0; w < total_items; w += unit_stride) {
        const int split = w % g.k_splits;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(kb0 + g.kb_per_split, g.kb_total);
        const int iters = (kb1 - kb0) * g.nparts;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = stages_u32 + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          const uint64_t adesc = adesc0 | static_cast<uint64_t>((sa >> 4) & 0x3fffu);
          const uint64_t bdesc = bdesc0 | static_cast<uint64_t>((sb >> 4) & 0x3fffu);
          if (elect_one()) {
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                if (CG == 2) umma_bf16_cg2(d_tmem, adesc + (uint64_t)(k * a_step), bdesc + (uint64_t)(k * b_step), idesc, (it > 0 || k > 0) ? 1u : 0u);
                else umma_bf16(d_tmem, adesc + (uint64_t)(k * a_step), bdesc + (uint64_t)(k * b_step), idesc, (it > 0 || k > 0) ? 1u : 0u);
              }
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if (CG == 2) umma_commit_cg2(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            // last k-block: accumulator ready for the epilogue warps (of both CTAs). Same thread as the MMAs by construction - tcgen05.commit
            // tracks the issuing thread's operations, so it must not depend on a second election naming the same lane
            if (it == iters - 1) { if (CG == 2) umma_commit_cg2(&tfull_bar[as]); else umma_commit(&tfull_bar[as]); }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    const int ew = warp - 4;
    const int q = ew & 3;       // TMEM lane quadrant == warp % 4
    const int hsel = ew >> 2;   // which 128-column half of the tile this warp drains
    const int c_lo = hsel * (BN / 64), c_hi = c_lo + BN / 64;
    uint8_t* stg = staging + ew * STG_WARP_BYTES;
    int as = 0; uint32_t aphase = 0;
    const uint32_t tempty0_leader = CG == 2 ? mapa_u32(&tempty_bar[0], 0) : 0u;
    for (int w = unit0; w < total_items; w += unit_stride) {
      const int tile = w / g.k_splits;
      int n_blk, m_blk;
      tile_mn(g, tile, m_blk, n_blk);
      const int row_base = (m_blk * CG + cta_rank) * BM + q * 32;
      const int m = row_base + lane;

      // shared bias array: generic STORE epilogue only (the specialised ones read the bias from global memory, N % 256 == 0 there)
      if ((EPI == CLIPDLM_EPI_STORE || EPI == CLIPDLM_EPI_STORE_GELU_DERIV) && g.bias != nullptr && (g.fast_mode < 0 || (g.dbg & 8192u))) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // previous tile's readers done
        for (int i = threadIdx.x - 128; i < BN; i += 32 * EPI_WARPS) {
          const int n = n_blk * BN + i;
          sbias[i] = n < g.N ? __ldcg(g.bias + n) : 0.f;   // L2 (coherence point of the optimizer's peer / multicast stores), not L1
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      }

      const bool valid = m < g.M;
      const long long mr = valid ? map_row(g, m) : 0;
      if (EPI == CLIPDLM_EPI_STORE_ROWSCALE || EPI == CLIPDLM_EPI_STORE_GELU_DERIV || EPI == CLIPDLM_EPI_STORE_MULAUX) {
        // single-mode instantiations of the specialised epilogue (the host checked its preconditions): their own kernels, so the
        // STORE kernels of the default path do not change
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c_lo * 32;
        const int col0 = n_blk * BN + c_lo * 32;
        const uint32_t sb = smem_u32(sbias) + c_lo * 128;
        if (EPI == CLIPDLM_EPI_STORE_ROWSCALE)      // out = acc * row_scale[m] + residual
          epi_store_fast<false, 1, false, false, true>(g, ta, col0, m, valid, mr, sb, &tfull_bar[as], aphase, &tmO, &tmO2, smem_u32(stg), row_base);
        else if (EPI == CLIPDLM_EPI_STORE_GELU_DERIV)   // out = gelu'(acc + bias), out2 = gelu(acc + bias)
          epi_store_fast<true, 0, true, false, false, 1>(g, ta, col0, m, valid, mr, sb, &tfull_bar[as], aphase, &tmO, &tmO2, smem_u32(stg), row_base);
        else                                            // out = acc * u  (u = the stored gelu')
          epi_store_fast<false, 2, false, false, false, 2>(g, ta, col0, m, valid, mr, sb, &tfull_bar[as], aphase, &tmO, &tmO2, smem_u32(stg), row_base);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty0_leader + as * 8);
          else mbar_arrive(&tempty_bar[as]);
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
        continue;
      }
      if (EPI == CLIPDLM_EPI_STORE && g.fast_mode >= 0) {
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c_lo * 32;
        const int col0 = n_blk * BN + c_lo * 32;
        const uint32_t sb = smem_u32(sbias) + c_lo * 128;
#define FAST_CASE(MODE, BIAS, AUX, DUAL, DROP) \
  case MODE: epi_store_fast<BIAS, AUX, DUAL, DROP>(g, ta, col0, m, valid, mr, sb, &tfull_bar[as], aphase, &tmO, &tmO2, smem_u32(stg), row_base); break;
        switch (g.fast_mode) {
          FAST_CASE(0, false, 0, false, false)    // dgrad
          FAST_CASE(1, true, 0, false, false)     // q/k/v projection
          FAST_CASE(2, false, 1, false, false)    // dgrad + residual branch (also the scattered lm_head dgrad)
          FAST_CASE(3, true, 1, false, false)     // out projection / lin2 (eval) + residual
          FAST_CASE(4, false, 2, false, false)    // lin2 dgrad * gelu'(u)
          FAST_CASE(9, true, 0, true, false)      // lin1 / vocab_transform: pre-activation + gelu
          FAST_CASE(19, true, 1, false, true)     // lin2 (train): bias, dropout, residual
          default: break;
        }
#undef FAST_CASE
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty0_leader + as * 8);
          else mbar_arrive(&tempty_bar[as]);
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
        continue;
      }
      // generic path. plain-bf16 residual / gelu'(u) operand: prefetch its first 32-column chunk while the MMAs of this tile are still running
      const bool al32 = g.al32 != 0;
      const bool aux_is_u = g.u_hi != nullptr;
      const __nv_bfloat16* aux = aux_is_u ? g.u_hi : g.res_hi;
      const long long ld_aux = aux_is_u ? g.ldu : g.ldr;
      const bool fast_aux = EPI == CLIPDLM_EPI_STORE && aux != nullptr && g.u_lo == nullptr && g.res_lo == nullptr &&
                            !(g.u_hi != nullptr && g.res_hi != nullptr);
      const bool dbg_nostore = (g.dbg & 1u) != 0, dbg_noaux = (g.dbg & 2u) != 0, dbg_drain = (g.dbg & 4u) != 0;
      uint32_t aux_nxt[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) aux_nxt[i] = 0u;
      if (fast_aux && valid && !dbg_noaux && n_blk * BN + c_lo * 32 < g.N) ld_row64(aux + mr * ld_aux + n_blk * BN + c_lo * 32, al32, aux_nxt);

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

      if (EPI == CLIPDLM_EPI_LSE) {
        // Online log-sum-exp over this warp's 128 columns, in the base-2 domain: one FMNMX + FFMA + MUFU.EX2 + FADD per logit.
        // The running arg-max (3 more instructions per logit) is only tracked when the caller asked for it (denoise loop).
        constexpr float LOG2E = 1.4426950408889634f;
        float mx = -INFINITY, sum = 0.f; int arg = 0;
        const int tgt = (m < g.M && g.targets != nullptr) ? g.targets[m % g.tgt_period] : -1;
        const bool want_arg = g.part_arg != nullptr;
        float tl = 0.f; bool has_t = false;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= g.N) break;
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          // training, plain bf16: keep the logits (bf16) so the backward turns them into d(logits) in place instead of recomputing this GEMM
          if (g.out_hi != nullptr && m < g.M) row_store_pair(g.out_hi + (long long)m * g.ldo + n0, nullptr, g.al32 != 0, v);
          if (n0 + 32 > g.N) {   // vocabulary tail (last tile only): padding columns never win and add exp(-inf) = 0
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (n0 + j) < g.N ? v[j] : -INFINITY;
          }
          const float prev_max = mx;
          if (want_arg) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (v[j] > mx) { mx = v[j]; arg = n0 + j; }  // strict '>' keeps the first maximum (torch.argmax tie rule)
          } else {
            float m0 = fmaxf(v[0], v[1]), m1 = fmaxf(v[2], v[3]), m2 = fmaxf(v[4], v[5]), m3 = fmaxf(v[6], v[7]);
#pragma unroll
            for (int j = 8; j < 32; j += 4) { m0 = fmaxf(m0, v[j]); m1 = fmaxf(m1, v[j + 1]); m2 = fmaxf(m2, v[j + 2]); m3 = fmaxf(m3, v[j + 3]); }
            mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          }
          if ((unsigned)(tgt - n0) < 32u) {   // the target logit lives in this chunk (1 chunk in ~950 per row)
#pragma unroll
            for (int j = 0; j < 32; ++j) if (n0 + j == tgt) tl = v[j];
            has_t = true;
          }
          // rescale the running sum from prev_max to the new max (every chunk has >= 1 valid column, so mx is finite)
          const float nb = -mx * LOG2E;
          float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            c0 += ex2_ftz(fmaf(v[j], LOG2E, nb)); c1 += ex2_ftz(fmaf(v[j + 1], LOG2E, nb));
            c2 += ex2_ftz(fmaf(v[j + 2], LOG2E, nb)); c3 += ex2_ftz(fmaf(v[j + 3], LOG2E, nb));
          }
          sum = fmaf(sum, ex2_ftz(fmaf(prev_max, LOG2E, nb)), (c0 + c1) + (c2 + c3));
        }
        if (m < g.M) {  // one partial per (row, 128-column half tile); an all-padding half leaves the neutral (-inf, 0)
          const size_t slot = (size_t)(n_blk * 2 + hsel) * g.M + m;
          g.part_max[slot] = mx;
          g.part_sum[slot] = sum;
          if (want_arg) g.part_arg[slot] = arg;
          if (has_t && g.tgt_logit != nullptr) g.tgt_logit[m] = tl;
        }
      } else if (EPI == CLIPDLM_EPI_LSE_EXP) {
        // Factored softmax gradient, pass 1: e = exp(logit - shift) with a per-launch constant shift - no running maximum, one FFMA +
        // FMNMX + MUFU.EX2 + FADD per logit. bf16(e) is what the gradient GEMM reads; the fp32 sum over this warp's 128 columns is the
        // partial of log-sum-exp = shift + log(sum e), reported with part_max = shift so that lse_combine_kernel serves both variants.
        constexpr float LOG2E = 1.4426950408889634f;
        const float shift = g.lse != nullptr ? *g.lse : 0.f;
        const float nb = -shift * LOG2E;
        const int tgt = (m < g.M && g.targets != nullptr) ? g.targets[m % g.tgt_period] : -1;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, tl = 0.f; bool has_t = false;
        // The 8 GB of bf16 exponentials per chunk leave through shared memory + TMA stores (two alternating [32 rows][64 columns] swizzled
        // staging tiles per warp, as in epi_store_fast): the row-wise 64-byte stores of row_store_pair cost one L1 transaction per sector and
        // kept this pass store-bound at ~1065 TFLOP/s whatever the tile order.
        const bool tma = g.tma_out != 0;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          const int n0 = n_blk * BN + c * 32;
          const int cc = c - c_lo;
          if (n0 >= g.N && (!tma || (cc & 1) == 0)) break;   // (a staged pair is always completed: its second half is written as zeros)
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          if ((unsigned)(tgt - n0) < 32u) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (n0 + j == tgt) tl = v[j];
            has_t = true;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ex2_ftz(fminf(fmaf(v[j], LOG2E, nb), 100.f));
          if (n0 + 32 > g.N) {   // vocabulary tail: the padding columns add nothing (the gradient GEMM never reads them: TMA clips at K = N)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (n0 + j) < g.N ? v[j] : 0.f;
          }
          if (tma) {
            const uint32_t buf = smem_u32(stg) + (uint32_t)((cc >> 1) * 4096);
            if ((cc & 1) == 0) {
              if (lane == 0) bulk_wait_read<1>();   // the store that last read this buffer (two commits ago) has drained it
              __syncwarp();
            }
            stage_row32(buf, lane, cc & 1, v);
            if (cc & 1) {
              fence_async_smem();
              __syncwarp();
              if (lane == 0) { tma_store_2d(&tmO, buf, n_blk * BN + (c - 1) * 32, row_base); bulk_commit(); }   // rows >= M are clipped by the map
            }
          } else if (m < g.M) {
            row_store_pair(g.out_hi + (long long)m * g.ldo + n0, nullptr, g.al32 != 0, v);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) { s0 += v[j]; s1 += v[j + 1]; s2 += v[j + 2]; s3 += v[j + 3]; }
        }
        if (m < g.M) {
          const size_t slot = (size_t)(n_blk * 2 + hsel) * g.M + m;
          g.part_max[slot] = shift;
          g.part_sum[slot] = (s0 + s1) + (s2 + s3);
          if (has_t && g.tgt_logit != nullptr) g.tgt_logit[m] = tl;
        }
      } else {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          const int n0 = n_blk * BN + c * 32;
          if (EPI != CLIPDLM_EPI_SMGRAD && n0 >= g.N) break;
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          if (EPI == CLIPDLM_EPI_WGRAD) {
            tile_store_f32<true>(g, g.out_f32, g.ldo, row_base, n0, stg, v);
          } else if (EPI == CLIPDLM_EPI_SMGRAD) {
            // (softmax - onehot) * scale with the scale folded into the exponent: one FFMA + MUFU.EX2 per logit
            constexpr float LOG2E = 1.4426950408889634f;
            const int tgt = (m < g.M) ? g.targets[m % g.tgt_period] : -1;
            const float l = (m < g.M) ? g.lse[m] : 0.f;
            const float gsc = g.part_max != nullptr ? g.grad_scale * *g.part_max : g.grad_scale;   // optional device-resident multiplier
            const float nb = gsc > 0.f ? fmaf(-l, LOG2E, __log2f(gsc)) : -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ex2_ftz(fmaf(v[j], LOG2E, nb));
            if (n0 + 32 > g.N) {   // zero the padding columns of the vocabulary tail
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (n0 + j) < g.N ? v[j] : 0.f;
            }
            if ((unsigned)(tgt - n0) < 32u) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (n0 + j == tgt) v[j] -= gsc;
            }
            if (valid) row_store_pair(g.out_hi + mr * g.ldo + n0, g.out_lo ? g.out_lo + mr * g.ldo + n0 : nullptr, al32, v);
          } else {  // STORE
            if (dbg_drain) {
              if (v[0] == 1.2345e-30f && valid) g.out_hi[0] = __float2bfloat16(v[1]);  // keep the TMEM load live
              continue;
            }
            uint32_t aux_cur[16];
            if (fast_aux) {
#pragma unroll
              for (int j = 0; j < 16; ++j) aux_cur[j] = aux_nxt[j];
              if (valid && !dbg_noaux && c + 1 < c_hi && n0 + 32 < g.N) ld_row64(aux + mr * ld_aux + n0 + 32, al32, aux_nxt);
            }
            if (g.bias != nullptr) {   // explicit ld.shared (warp-wide broadcast): through the generic pointer these compile to LD
              const uint32_t sb = smem_u32(sbias) + c * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 b4;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(sb + j * 16));
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
            }
            if (g.drop.thresh16 != 0) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                const unsigned long long idx = (unsigned long long)m * (unsigned long long)g.N + (unsigned long long)(n0 + j8 * 8);
                const uint32_t keep = dropout_keep8(g.drop, idx >> 3);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[j8 * 8 + i] = ((keep >> i) & 1u) ? v[j8 * 8 + i] * g.drop.scale : 0.f;
              }
            }
            if (fast_aux) {
              if (aux_is_u) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float2 f = unpack_bf16x2(aux_cur[j]);
                  v[2 * j] *= dgelu_f(f.x);
                  v[2 * j + 1] *= dgelu_f(f.y);
                }
              } else {
                row_unpack<true>(aux_cur, v);
              }
            } else if (valid) {
              if (g.u_hi != nullptr) {
                float u[32];
                row_load_pair(g.u_hi + mr * g.ldu + n0, g.u_lo ? g.u_lo + mr * g.ldu + n0 : nullptr, al32, u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= dgelu_f(u[j]);
              }
              if (g.res_hi != nullptr) {
                float r[32];
                row_load_pair(g.res_hi + mr * g.ldr + n0, g.res_lo ? g.res_lo + mr * g.ldr + n0 : nullptr, al32, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += r[j];
              }
            }
            if (valid && dbg_nostore) {
              float sacc = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) sacc += (g.out2_hi != nullptr) ? gelu_f(v[j]) + v[j] : v[j];
              if (sacc == 1.2345e-30f) g.out_hi[0] = __float2bfloat16(sacc);
            } else if (valid) {
              if (g.out_hi != nullptr) row_store_pair(g.out_hi + mr * g.ldo + n0, g.out_lo ? g.out_lo + mr * g.ldo + n0 : nullptr, al32, v);
              if (g.out_f32 != nullptr) {
                float* dst = g.out_f32 + mr * g.ldo + n0;
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
              if (g.out2_hi != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
                row_store_pair(g.out2_hi + mr * g.ldo + n0, g.out2_lo ? g.out2_lo + mr * g.ldo + n0 : nullptr, al32, v);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty0_leader + as * 8);   // the MMA issuer lives in the leader CTA
        else mbar_arrive(&tempty_bar[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  if (warp >= 4 && lane == 0) bulk_wait_read<0>();   // outstanding TMA stores still read this CTA's staging tiles
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: nobody leaves while the peer may still read its smem / signal its barriers
  if (warp == 2) {
    if (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// LSE partial combine: one thread per row, tiles scanned in order (lowest tile wins argmax ties).
// ------------------------------------------------------------------------------------------------
__global__ void lse_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, const int* __restrict__ parg,
                                   int n_tiles, int M, const float* __restrict__ tgt_logit, float* __restrict__ lse,
                                   int* __restrict__ argmax, double* loss_acc, double scale) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (m < M) {
    float mx = -INFINITY; int arg = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const float v = pmax[(size_t)t * M + m];
      if (v > mx) { mx = v; if (parg != nullptr) arg = parg[(size_t)t * M + m]; }
    }
    float s = 0.f;
    for (int t = 0; t < n_tiles; ++t) s += psum[(size_t)t * M + m] * __expf(pmax[(size_t)t * M + m] - mx);
    const float l = mx + logf(s);
    if (lse != nullptr) lse[m] = l;
    if (argmax != nullptr) argmax[m] = arg;
    if (tgt_logit != nullptr) loss = l - tgt_logit[m];
  }
  if (loss_acc != nullptr) {
    loss = warp_sum(loss);
    __shared__ float wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += (double)wsum[i];
      atomicAdd(loss_acc, t * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// d(logits) in place from stored bf16 logits: x[m, n] <- (exp(x[m, n] - lse[m]) - [n == tgt[m]]) * scale  (n < N), 0 for the padding
// columns n >= N.  One 16-byte vector (8 logits) per thread, grid-stride; HBM-bound (read + write of the [M, ld] buffer).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_grad_inplace_kernel(__nv_bfloat16* __restrict__ x, long long ld, int M, int N,
                                                                   const float* __restrict__ lse, const int* __restrict__ targets, int tgt_period,
                                                                   float scale, float log2_scale, const float* __restrict__ scale_mul) {
  if (scale_mul != nullptr) {   // device-resident multiplier of the gradient scale (dynamic rounding weight)
    const float mul = *scale_mul;
    scale *= mul;
    log2_scale = mul > 0.f ? log2_scale + __log2f(mul) : -INFINITY;
  }
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr int U = 5;   // 16-byte vectors in flight per thread (5 x 2048 threads x 16 B = 160 KB per SM outstanding)
  const int vec_per_row = (int)(ld >> 3);
  for (int m = blockIdx.x; m < M; m += gridDim.x) {   // one row per block iteration: no index divisions, lse / target read once
    uint4* row = reinterpret_cast<uint4*>(x + (long long)m * ld);
    const float nb = fmaf(-lse[m], LOG2E, log2_scale);
    const int tgt = targets[m % tgt_period];
    for (int v0 = threadIdx.x; v0 < vec_per_row; v0 += 256 * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int vi = v0 + u * 256;
        raw[u] = (vi < vec_per_row && vi * 8 < N) ? row[vi] : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int vi = v0 + u * 256;
        if (vi >= vec_per_row) continue;
        const int n0 = vi * 8;
        if (n0 >= N) { row[vi] = make_uint4(0u, 0u, 0u, 0u); continue; }
        float v[8];
        { float2 a = unpack_bf16x2(raw[u].x), b = unpack_bf16x2(raw[u].y), c = unpack_bf16x2(raw[u].z), d = unpack_bf16x2(raw[u].w);
          v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y; }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ex2_ftz(fmaf(v[j], LOG2E, nb));
        if ((unsigned)(tgt - n0) < 8u) {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (n0 + j == tgt) v[j] -= scale;
        }
        if (n0 + 8 > N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (n0 + j) < N ? v[j] : 0.f;
        }
        row[vi] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
    }
  }
}

int softmax_grad_inplace_dispatch(void* logits, long long ld, int M, int N, const float* lse, const int* targets, int tgt_period, float scale,
                                  cudaStream_t st, const float* scale_mul) {
  CLIPDLM_CHECK(logits && lse && targets && M > 0 && N > 0 && ld >= N && ld % 8 == 0 && tgt_period > 0, "softmax_grad_inplace: bad arguments");
  CLIPDLM_CHECK((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "softmax_grad_inplace: buffer must be 16-byte aligned");
  int sms = 148;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  long long blocks = M;
  if (blocks > 8LL * sms) blocks = 8LL * sms;
  softmax_grad_inplace_kernel<<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)logits, ld, M, N, lse, targets, tgt_period, scale,
                                                                scale > 0.f ? log2f(scale) : -INFINITY, scale_mul);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Row terms of the factored softmax gradient (CLIPDLM_EPI_LSE_EXP / _STORE_ROWSCALE): one warp per lm_head row.
//   row_scale[m] = scale * exp(shift - lse[m]);   dx[row(m), :] -= scale * W[tgt(m), :]   (bf16 read-modify-write, fp32 math)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ce_row_terms_kernel(const float* __restrict__ lse, const float* __restrict__ exp_shift,
                                                           const int* __restrict__ targets, int tgt_period, float scale, int M,
                                                           const __nv_bfloat16* __restrict__ W, long long ldw, __nv_bfloat16* __restrict__ dx,
                                                           long long ldx, int scatter_len, int scatter_stride, int D, float* __restrict__ row_scale,
                                                           const float* __restrict__ scale_mul) {
  if (scale_mul != nullptr) scale *= *scale_mul;   // device-resident multiplier of the gradient scale (dynamic rounding weight)
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (m >= M) return;
  if (lane == 0) {
    const float shift = exp_shift != nullptr ? *exp_shift : 0.f;
    const float r = scale * __expf(shift - lse[m]);
    row_scale[m] = (r == r && fabsf(r) <= 3.0e38f) ? r : 0.f;   // sum exp(s - shift) flushed to zero or overflowed: no softmax term rather than NaN
  }
  const int tgt = targets[m % tgt_period];
  const long long row = scatter_len > 0 ? (long long)(m / scatter_len) * scatter_stride + (m % scatter_len) : (long long)m;
  const uint4* w = reinterpret_cast<const uint4*>(W + (long long)tgt * ldw);
  uint4* d = reinterpret_cast<uint4*>(dx + row * ldx);
  for (int v = lane; v < D / 8; v += 32) {
    const uint4 a = d[v], b = w[v];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16x2(aw[i]), y = unpack_bf16x2(bw[i]);
      o[i] = pack_bf16x2(fmaf(-scale, y.x, x.x), fmaf(-scale, y.y, x.y));
    }
    d[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int ce_row_terms_dispatch(const float* lse, const float* exp_shift, const int* targets, int tgt_period, float scale, int M, const void* w,
                          long long ldw, void* dx, long long ldx, int scatter_len, int scatter_stride, int D, float* row_scale, cudaStream_t st,
                          const float* scale_mul) {
  CLIPDLM_CHECK(lse && targets && w && dx && row_scale && M > 0 && tgt_period > 0 && D > 0, "ce_row_terms: bad arguments");
  CLIPDLM_CHECK(D % 8 == 0 && ldw % 8 == 0 && ldx % 8 == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
                "ce_row_terms: rows must be 16-byte aligned (D, pitches multiples of 8 elements)");
  CLIPDLM_CHECK(scatter_len >= 0 && (scatter_len == 0 || scatter_stride >= scatter_len), "ce_row_terms: bad scatter %d / %d", scatter_len, scatter_stride);
  ce_row_terms_kernel<<<(unsigned)((M + 3) / 4), 128, 0, st>>>(lse, exp_shift, targets, tgt_period, scale, M, (const __nv_bfloat16*)w, ldw,
                                                              (__nv_bfloat16*)dx, ldx, scatter_len, scatter_stride, D, row_scale, scale_mul);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// 2-D / 3-D bf16 tensor map with 128-byte swizzle. dims/strides innermost first; strides in bytes for dims 1..rank-1.
static int encode_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  auto fn = get_encode_fn();
  CLIPDLM_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CLIPDLM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu box %u %u stride %llu base %p)",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1],
                (unsigned long long)strides_bytes[0], base);
  return 0;
}

// 2-D bf16 tensor map (128-byte swizzle) for other kernels of the library (attention): dims {inner, outer}, box {box_inner, box_outer}.
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, unsigned long long inner, unsigned long long outer, unsigned long long pitch_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  CLIPDLM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && pitch_bytes % 16 == 0, "tensor map: base / pitch must be 16-byte aligned");
  uint64_t dims[2] = {inner, outer};
  uint64_t str[1] = {pitch_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return encode_map(tm, base, 2, dims, str, box);
}

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, const unsigned long long* dims, const unsigned long long* strides_bytes,
                      const uint32_t* box) {
  CLIPDLM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && strides_bytes[0] % 16 == 0 && strides_bytes[1] % 16 == 0,
                "tensor map: base / strides must be 16-byte aligned");
  uint64_t d[3] = {dims[0], dims[1], dims[2]};
  uint64_t s[2] = {strides_bytes[0], strides_bytes[1]};
  return encode_map(tm, base, 3, d, s, box);
}

// Operand map. major 0: stored [rows][K] pitch ld -> dims {K, rows}, box {64, tile_rows}.
//              major 1: stored [K][rows] pitch ld -> dims {rows, K}, box {64, 64}.
static int operand_map(CUtensorMap* tm, const void* base, int major, int rows, int K, long long ld, int tile_rows, int gather_len,
                       int gather_stride) {
  CLIPDLM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "GEMM operand base %p not 16-byte aligned", base);
  CLIPDLM_CHECK((ld * 2) % 16 == 0, "GEMM operand pitch %lld elements not a multiple of 8", ld);
  if (major == 0) {
    if (gather_len > 0) {
      CLIPDLM_CHECK(tile_rows % gather_len == 0 && rows % gather_len == 0, "gather_len %d must divide %d and M %d", gather_len,
                    tile_rows, rows);
      uint64_t dims[3] = {(uint64_t)K, (uint64_t)gather_len, (uint64_t)(rows / gather_len)};
      uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)gather_stride};
      uint32_t box[3] = {64, (uint32_t)gather_len, (uint32_t)(tile_rows / gather_len)};
      return encode_map(tm, base, 3, dims, str, box);
    }
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)ld * 2};
    uint32_t box[2] = {64, (uint32_t)tile_rows};
    return encode_map(tm, base, 2, dims, str, box);
  }
  CLIPDLM_CHECK(gather_len == 0, "gather is only supported for K-major A");
  uint64_t dims[2] = {(uint64_t)rows, (uint64_t)K};
  uint64_t str[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {64, 64};
  return encode_map(tm, base, 2, dims, str, box);
}

static uint32_t g_dbg_mn_lbo = 0, g_dbg_mn_sbo = 0, g_dbg_flags = 0;
static int g_num_sms = 0;

template <int AMAJ, int BMAJ, int EPI, int CG>
static int launch_gemm_cg(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1, const CUtensorMap& o0,
                          const CUtensorMap& o1, const GemmArgs& ga, int units, cudaStream_t st) {
  static bool attr_set = false;
  static int max_units = 0;   // co-resident work units (CTAs / CTA pairs) of this instantiation: the kernel is persistent
  auto kfn = gemm_kernel<AMAJ, BMAJ, EPI, CG>;
  if (!attr_set) {
    CLIPDLM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<CG>::SMEM_BYTES));
    max_units = g_num_sms / CG;
    if (CG == 2) {
      // SM floor-sweeping can leave TPCs with a single SM: only complete TPCs can host a CTA pair. A persistent grid larger
      // than the number of co-resident pairs would run its tail as a second wave (2x the time), so ask the driver.
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3((unsigned)g_num_sms, 1, 1);
      q.blockDim = dim3(NUM_THREADS, 1, 1);
      q.dynamicSmemBytes = Geo<CG>::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CG; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n = 0;
      CLIPDLM_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kfn, &q));
      if (n > 0 && n < max_units) max_units = n;
      if (getenv("CLIPDLM_DEBUG")) fprintf(stderr, "clipdlm: gemm<%d,%d,%d,cg2> max active CTA pairs %d (SMs %d)\n", AMAJ, BMAJ, EPI, n, g_num_sms);
    }
    attr_set = true;
  }
  if (units > max_units) units = max_units;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(units * CG), 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Geo<CG>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dbg_flags & 512u ? 1 : 2;   // debug bit 9: no programmatic dependent launch
  CLIPDLM_CUDA_OK(cudaLaunchKernelEx(&cfg, kfn, a0, a1, b0, b1, o0, o1, ga));
  return 0;
}
template <int AMAJ, int BMAJ, int EPI>
static int launch_gemm(int cg, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1, const CUtensorMap& o0,
                       const CUtensorMap& o1, const GemmArgs& ga, int units, cudaStream_t st) {
  if (cg == 2) return launch_gemm_cg<AMAJ, BMAJ, EPI, 2>(a0, a1, b0, b1, o0, o1, ga, units, st);
  return launch_gemm_cg<AMAJ, BMAJ, EPI, 1>(a0, a1, b0, b1, o0, o1, ga, units, st);
}

int gemm_dispatch(const clipdlm_gemm_t* g, cudaStream_t st) {
  CLIPDLM_CHECK(g != nullptr, "null gemm descriptor");
  CLIPDLM_CHECK(g->M > 0 && g->N > 0 && g->K > 0, "bad GEMM shape %d %d %d", g->M, g->N, g->K);
  CLIPDLM_CHECK(g->a_hi && g->b_hi, "null GEMM operand");
  if (g_num_sms == 0) {
    int dev = 0;
    CLIPDLM_CUDA_OK(cudaGetDevice(&dev));
    CLIPDLM_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // CTA pairs (256 x 256 tiles, tcgen05.mma.cta_group::2) unless the problem has a single 128-row tile or the debug hook says no
  const int cg = (g->M > BM && !(g_dbg_flags & 8u) && g_num_sms >= 2) ? 2 : 1;
  const int units_max = g_num_sms / cg;
  CUtensorMap a0, a1, b0, b1;
  int rc;
  if ((rc = operand_map(&a0, g->a_hi, g->a_major, g->M, g->K, g->lda, BM, g->gather_len, g->gather_stride))) return rc;
  if ((rc = operand_map(&b0, g->b_hi, g->b_major, g->N, g->K, g->ldb, BN / cg, 0, 0))) return rc;
  a1 = a0; b1 = b0;
  if (g->a_lo && (rc = operand_map(&a1, g->a_lo, g->a_major, g->M, g->K, g->lda, BM, g->gather_len, g->gather_stride))) return rc;
  if (g->b_lo && (rc = operand_map(&b1, g->b_lo, g->b_major, g->N, g->K, g->ldb, BN / cg, 0, 0))) return rc;

  GemmArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.M = g->M; ga.N = g->N; ga.K = g->K;
  ga.num_m_tiles = (g->M + BM * cg - 1) / (BM * cg);
  ga.num_n_tiles = (g->N + BN - 1) / BN;
  ga.kb_total = (g->K + BK - 1) / BK;
  ga.k_splits = 1;
  if (g->epilogue == CLIPDLM_EPI_WGRAD) {
    const int tiles = ga.num_m_tiles * ga.num_n_tiles;
    int ks = g->k_splits;
    if (ks <= 0) {
      // split-K factor minimising the makespan: waves x (k-blocks per split + the cost of draining one accumulator tile with
      // fp32 red.add, ~12 k-block times), k-blocks per split kept >= 8
      const int max_ks = ga.kb_total / 8 > 0 ? ga.kb_total / 8 : 1;
      long long best = -1;
      ks = 1;
      for (int c = 1; c <= max_ks && c <= 4096; ++c) {
        const int per = (ga.kb_total + c - 1) / c;
        const int eff = (ga.kb_total + per - 1) / per;
        const long long waves = ((long long)tiles * eff + units_max - 1) / units_max;
        const long long cost = waves * (per + 12);
        if (best < 0 || cost < best) { best = cost; ks = eff; }
      }
    }
    if (ks > ga.kb_total) ks = ga.kb_total;
    ga.kb_per_split = (ga.kb_total + ks - 1) / ks;
    ga.k_splits = (ga.kb_total + ga.kb_per_split - 1) / ga.kb_per_split;
  } else {
    ga.kb_per_split = ga.kb_total;
  }
  ga.nparts = 1; ga.part_a[0] = 0; ga.part_b[0] = 0;
  if (g->a_lo) { ga.part_a[ga.nparts] = 1; ga.part_b[ga.nparts] = 0; ga.nparts++; }
  if (g->b_lo) { ga.part_a[ga.nparts] = 0; ga.part_b[ga.nparts] = 1; ga.nparts++; }
  ga.gather_len = g->gather_len;
  // lm_head GEMMs (A gathered from x_out[:, :16] / output scattered back): B is the embedding table, re-read by every row tile while the
  // logits (written, or read exactly once as the A operand of the gradient GEMM) stream through L2 - keep the table, let the logits go
  // (measured, profiles/r02_ncu_launches_train_step.csv: evict_last on the table + evict_first on the streamed operand made the DRAM reads of the
  // lm_head passes WORSE - 4.3 GB instead of 2.4 GB (LSE), 21 GB instead of 8 GB (gradient GEMM: its three column tiles re-read every A tile from
  // L2, which evict_first defeats). The hints stay plumbed but off; the tile order below is what keeps the table resident.)
  ga.pol_a = 0ull;
  ga.pol_b = 0ull;
  ga.mn_lbo = g_dbg_mn_lbo ? g_dbg_mn_lbo : 8192;
  ga.mn_sbo = g_dbg_mn_sbo ? g_dbg_mn_sbo : 1024;
  ga.out_hi = (__nv_bfloat16*)g->out_hi; ga.out_lo = (__nv_bfloat16*)g->out_lo;
  ga.out2_hi = (__nv_bfloat16*)g->out2_hi; ga.out2_lo = (__nv_bfloat16*)g->out2_lo;
  ga.out_f32 = g->epilogue == CLIPDLM_EPI_WGRAD ? g->acc_f32 : g->out_f32;
  ga.ldo = g->ldo;
  ga.bias = g->bias;
  ga.res_hi = (const __nv_bfloat16*)g->res_hi; ga.res_lo = (const __nv_bfloat16*)g->res_lo; ga.ldr = g->ldr;
  ga.u_hi = (const __nv_bfloat16*)g->u_hi; ga.u_lo = (const __nv_bfloat16*)g->u_lo; ga.ldu = g->ldu;
  ga.scatter_len = g->scatter_len; ga.scatter_stride = g->scatter_stride;
  {
    const void* ptrs[8] = {g->out_hi, g->out_lo, g->out2_hi, g->out2_lo, g->res_hi, g->res_lo, g->u_hi, g->u_lo};
    bool ok = (g->ldo % 16 == 0) && (!g->res_hi || g->ldr % 16 == 0) && (!g->u_hi || g->ldu % 16 == 0);
    for (const void* q : ptrs) ok = ok && ((reinterpret_cast<uintptr_t>(q) & 31) == 0);
    ga.al32 = ok ? 1 : 0;
    ok = (g->ldo % 8 == 0) && (!g->res_hi || g->ldr % 8 == 0) && (!g->u_hi || g->ldu % 8 == 0);
    for (const void* q : ptrs) ok = ok && ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
    if (g->epilogue == CLIPDLM_EPI_STORE || g->epilogue == CLIPDLM_EPI_SMGRAD)
      CLIPDLM_CHECK(ok && (!g->out_f32 || ((reinterpret_cast<uintptr_t>(g->out_f32) & 15) == 0 && g->ldo % 4 == 0)),
                    "GEMM epilogue operands must be 16-byte aligned with pitches that are multiples of 8 elements");
  }
  ga.drop.seed = g->drop_seed; ga.drop.site = g->drop_site;
  ga.drop.thresh16 = g->drop_p > 0.f ? (uint32_t)(g->drop_p * 65536.f + 0.5f) : 0u;
  ga.drop.scale = g->drop_p > 0.f ? 1.f / (1.f - g->drop_p) : 1.f;
  ga.part_max = g->part_max; ga.part_sum = g->part_sum; ga.part_arg = g->part_arg; ga.tgt_logit = g->tgt_logit;
  ga.targets = g->targets; ga.tgt_period = g->tgt_period > 0 ? g->tgt_period : 1;
  ga.lse = g->lse; ga.grad_scale = g->grad_scale;
  if (g->epilogue == CLIPDLM_EPI_LSE_EXP) ga.lse = g->exp_shift;
  if (g->epilogue == CLIPDLM_EPI_STORE_ROWSCALE) ga.lse = g->row_scale;
  ga.dbg = g_dbg_flags;
  // Bias of the specialised epilogues: the per-tile shared-memory array (kernel bit 13 set) unless debug bit 13 asks for direct global loads
  // (and the bias vector is 16-byte aligned). The direct loads measured 1.5-3.5 % faster on the bias GEMMs, but tests/test_dp_fused_gpu.py
  // (two GPUs, fused optimizer step against the NCCL step) diverged with them - with ld.global.nc on both exchange paths, with a plain
  // ld.global still on the peer-store path - so they stay an opt-in experiment until that is understood.
  if (g_dbg_flags & 8192u) ga.dbg &= ~8192u; else ga.dbg |= 8192u;
  if (g->bias != nullptr && (reinterpret_cast<uintptr_t>(g->bias) & 15) != 0) ga.dbg |= 8192u;
  ga.fast_mode = -1;
  if (g->epilogue == CLIPDLM_EPI_STORE_ROWSCALE) {
    CLIPDLM_CHECK(g->row_scale && g->out_hi && g->res_hi && !g->bias && !g->u_hi && !g->out2_hi && !g->out_f32 && !g->out_lo && !g->res_lo &&
                      !g->a_lo && !g->b_lo && g->drop_p == 0.f,
                  "STORE_ROWSCALE: plain-bf16 out = acc * row_scale + residual only");
    CLIPDLM_CHECK(ga.al32 && g->N % BN == 0, "STORE_ROWSCALE needs 32-byte aligned rows (pitches %% 16 == 0) and N %% 256 == 0 (N = %d)", g->N);
    ga.fast_mode = 2;   // epi_store_fast<no bias, residual> + row scale; also switches the TMA-store output maps on below
  }
  if (g->epilogue == CLIPDLM_EPI_STORE_GELU_DERIV) {
    CLIPDLM_CHECK(g->bias && g->out_hi && g->out2_hi && !g->res_hi && !g->u_hi && !g->out_f32 && !g->out_lo && !g->out2_lo && !g->a_lo && !g->b_lo &&
                      g->drop_p == 0.f && g->scatter_len == 0 && g->b_major == 0,
                  "STORE_GELU_DERIV: plain-bf16 out = gelu'(acc + bias), out2 = gelu(acc + bias), K-major operands only");
    CLIPDLM_CHECK(ga.al32 && g->N % BN == 0, "STORE_GELU_DERIV needs 32-byte aligned rows (pitch %% 16 == 0) and N %% 256 == 0 (N = %d)", g->N);
    ga.fast_mode = 9;
  }
  if (g->epilogue == CLIPDLM_EPI_STORE_MULAUX) {
    CLIPDLM_CHECK(g->u_hi && g->out_hi && !g->bias && !g->res_hi && !g->out2_hi && !g->out_f32 && !g->out_lo && !g->u_lo && !g->a_lo && !g->b_lo &&
                      g->drop_p == 0.f && g->b_major == 1,
                  "STORE_MULAUX: plain-bf16 out = acc * u with an MN-major B operand only");
    CLIPDLM_CHECK(ga.al32 && g->N % BN == 0, "STORE_MULAUX needs 32-byte aligned rows (pitches %% 16 == 0) and N %% 256 == 0 (N = %d)", g->N);
    ga.fast_mode = 4;
    if (g->acc_f32 != nullptr) {   // fused bias gradient: acc_f32[n] += sum_m out[m, n]; needs the TMA-store side of the epilogue (staged tiles)
      CLIPDLM_CHECK(g->scatter_len == 0 && !(g_dbg_flags & 1024u) && (reinterpret_cast<uintptr_t>(g->acc_f32) & 7) == 0,
                    "STORE_MULAUX column sums need an unscattered output and an 8-byte aligned fp32 accumulator");
      ga.part_max = g->acc_f32;
    }
  }
  if (g->epilogue == CLIPDLM_EPI_STORE && ga.al32 && g->N % BN == 0 && !g->out_lo && !g->out2_lo && !g->res_lo && !g->u_lo && !g->out_f32 &&
      !(g->res_hi && g->u_hi) && (g->out_hi || g->out2_hi) && !(g_dbg_flags & 7u)) {  // (bits 7, 8: triage of the dual-store mode)
    const int aux = g->u_hi ? 2 : (g->res_hi ? 1 : 0);
    const int mode = (g->bias ? 1 : 0) | (aux << 1) | (g->out2_hi ? 8 : 0) | (ga.drop.thresh16 ? 16 : 0);
    switch (mode) {
      case 0: case 1: case 2: case 3: case 4: case 9: case 19: ga.fast_mode = mode; break;
      default: break;
    }
    if (!(g_dbg_flags & 64u) && ga.fast_mode >= 0 && mode != 9 && !g->out_hi) ga.fast_mode = -1;   // only the dual-store mode may omit out
    if (g_dbg_flags & 64u) ga.fast_mode = -1;   // triage: force the generic epilogue
  }
  // bf16 outputs through TMA stores (box 64 columns x 32 rows out of the epilogue warps' swizzled staging tiles)
  CUtensorMap o0 = a0, o1 = a0;
  ga.tma_out = 0;
  if (g->epilogue == CLIPDLM_EPI_LSE_EXP && g->out_hi && !(g_dbg_flags & (1024u | 2048u)) && g->ldo % 8 == 0 && g->ldo >= (long long)ga.num_n_tiles * BN) {
    // the exponentials of the factored softmax gradient: [M][ldo] with the padding columns of the last vocabulary tile written as zeros
    uint64_t dims[2] = {(uint64_t)ga.num_n_tiles * BN, (uint64_t)g->M};
    uint64_t str[1] = {(uint64_t)g->ldo * 2};
    uint32_t box[2] = {64, 32};
    if ((rc = encode_map(&o0, g->out_hi, 2, dims, str, box))) return rc;
    ga.tma_out = 1;
  }
  if (ga.fast_mode >= 0 && g->scatter_len == 0 && !(g_dbg_flags & 1024u)) {
    uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)g->M};
    uint64_t str[1] = {(uint64_t)g->ldo * 2};
    uint32_t box[2] = {64, 32};
    if (g->out_hi && (rc = encode_map(&o0, g->out_hi, 2, dims, str, box))) return rc;
    if (g->out2_hi && (rc = encode_map(&o1, g->out2_hi, 2, dims, str, box))) return rc;
    ga.tma_out = 1;
  }

  const int total = ga.num_m_tiles * ga.num_n_tiles * ga.k_splits;
  const int grid = total < units_max ? total : units_max;   // work units: CTAs (cg = 1) or CTA pairs (cg = 2)
  // band order for the vocabulary-wide passes (LSE / LSE_EXP: N = 120 column tiles): one band = the row tiles one wave of units takes
  ga.band = 0;
  // (measured on one box, tools/gemm_perf.py --flags 0,64 at 8192 rows: LSE 4.51 vs 4.95 ms, LSE_EXP with its 8 GB output 5.63 vs 5.71 ms; the
  // split-precision SMGRAD recompute pass was 10 % SLOWER in band order - three operand passes per tile - and keeps the default order)
  if ((g->epilogue == CLIPDLM_EPI_LSE || g->epilogue == CLIPDLM_EPI_LSE_EXP) && ga.k_splits == 1 &&
      ga.num_n_tiles >= 16 && ga.num_m_tiles > 1 && !(g_dbg_flags & 4096u))
    ga.band = grid < ga.num_m_tiles ? grid : ga.num_m_tiles;

  switch (g->epilogue) {
    case CLIPDLM_EPI_STORE:
      CLIPDLM_CHECK(g->N % 32 == 0, "STORE epilogue needs N %% 32 == 0 (N = %d)", g->N);
      CLIPDLM_CHECK(g->out_hi || g->out_f32 || g->out2_hi, "STORE epilogue without output");
      CLIPDLM_CHECK(g->a_major == 0, "STORE epilogue expects K-major A");
      if (g->b_major == 0) return launch_gemm<0, 0, CLIPDLM_EPI_STORE>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
      return launch_gemm<0, 1, CLIPDLM_EPI_STORE>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    case CLIPDLM_EPI_WGRAD:
      CLIPDLM_CHECK(g->a_major == 1 && g->b_major == 1, "WGRAD epilogue expects MN-major A and B");
      CLIPDLM_CHECK(g->N % 32 == 0 && g->acc_f32, "WGRAD needs N %% 32 == 0 and an fp32 accumulator");
      return launch_gemm<1, 1, CLIPDLM_EPI_WGRAD>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    case CLIPDLM_EPI_LSE:
      CLIPDLM_CHECK(g->a_major == 0 && g->b_major == 0, "LSE epilogue expects K-major operands");
      CLIPDLM_CHECK(g->part_max && g->part_sum && (!g->targets || g->tgt_logit), "LSE epilogue buffers missing");
      CLIPDLM_CHECK(!g->out_hi || (!g->out_lo && g->ldo >= (long long)ga.num_n_tiles * BN && g->ldo % 8 == 0 &&
                                   (reinterpret_cast<uintptr_t>(g->out_hi) & 15) == 0),
                    "LSE epilogue: the optional bf16 logits output needs a 16-byte aligned plain-bf16 buffer with pitch >= %d", ga.num_n_tiles * BN);
      return launch_gemm<0, 0, CLIPDLM_EPI_LSE>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    case CLIPDLM_EPI_LSE_EXP:
      CLIPDLM_CHECK(g->a_major == 0 && g->b_major == 0, "LSE_EXP epilogue expects K-major operands");
      CLIPDLM_CHECK(g->part_max && g->part_sum && (!g->targets || g->tgt_logit) && !g->part_arg, "LSE_EXP epilogue buffers missing (no arg-max tracking in this mode)");
      CLIPDLM_CHECK(g->out_hi && !g->out_lo && !g->a_lo && !g->b_lo && g->ldo >= (long long)ga.num_n_tiles * BN && g->ldo % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(g->out_hi) & 15) == 0,
                    "LSE_EXP epilogue: needs a 16-byte aligned plain-bf16 output with pitch >= %d", ga.num_n_tiles * BN);
      return launch_gemm<0, 0, CLIPDLM_EPI_LSE_EXP>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    case CLIPDLM_EPI_STORE_ROWSCALE:
      CLIPDLM_CHECK(g->a_major == 0, "STORE_ROWSCALE epilogue expects K-major A");
      if (g->b_major == 0) return launch_gemm<0, 0, CLIPDLM_EPI_STORE_ROWSCALE>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
      return launch_gemm<0, 1, CLIPDLM_EPI_STORE_ROWSCALE>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    case CLIPDLM_EPI_STORE_GELU_DERIV:
      CLIPDLM_CHECK(g->a_major == 0, "STORE_GELU_DERIV epilogue expects K-major A");
      return launch_gemm<0, 0, CLIPDLM_EPI_STORE_GELU_DERIV>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    case CLIPDLM_EPI_STORE_MULAUX:
      CLIPDLM_CHECK(g->a_major == 0, "STORE_MULAUX epilogue expects K-major A");
      return launch_gemm<0, 1, CLIPDLM_EPI_STORE_MULAUX>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    case CLIPDLM_EPI_SMGRAD:
      CLIPDLM_CHECK(g->a_major == 0 && g->b_major == 0, "SMGRAD epilogue expects K-major operands");
      CLIPDLM_CHECK(g->out_hi && g->lse && g->targets, "SMGRAD epilogue buffers missing");
      CLIPDLM_CHECK(g->ldo >= (long long)ga.num_n_tiles * BN, "SMGRAD output pitch %lld < %d", (long long)g->ldo, ga.num_n_tiles * BN);
      return launch_gemm<0, 0, CLIPDLM_EPI_SMGRAD>(cg, a0, a1, b0, b1, o0, o1, ga, grid, st);
    default:
      CLIPDLM_CHECK(false, "unknown epilogue %d", g->epilogue);
  }
  return 0;
}

int lse_combine_dispatch(const float* pmax, const float* psum, const int* parg, int n_tiles, int M, const float* tgt_logit, float* lse,
                         int* argmax, double* loss_acc, double scale, cudaStream_t st) {
  const int threads = 256;
  lse_combine_kernel<<<(M + threads - 1) / threads, threads, 0, st>>>(pmax, psum, parg, n_tiles, M, tgt_logit, lse, argmax, loss_acc,
                                                                     scale);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

void gemm_debug_mn_desc(uint32_t lbo, uint32_t sbo) { g_dbg_mn_lbo = lbo; g_dbg_mn_sbo = sbo; }
void gemm_debug_flags(uint32_t flags) { g_dbg_flags = flags; }

}  // namespace clipdlm
