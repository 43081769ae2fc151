// Short-sequence attention on tcgen05, PACKED: the sequences of one 128-row UMMA tile sit back to back (slot = L rows) instead of in
// 32-row slots - 7 sequences of L = 18 (concat fusion) or L = 16 (add fusion) per tile where attention_umma.cu packs 4.  The tile then
// simply is 128 consecutive token rows of the [T, 3D] q|k|v matrix (one 2-D TMA box per operand, no zero fill), the output tile 126
// consecutive token rows, and every per-group cost (TMA round trip, two MMA groups, barrier hand-offs, epilogue) is spread over 1.75 x
// the work.  Replaces DistilBertSelfAttention / SDPA (HF modeling_distilbert.py:126-151,177-207) and its autograd backward for
// L in {16, 18}, plain bf16 storage, head dim 64.
//
// S = Q K^T is one 128 x 128 x 64 MMA group; sequence s owns the diagonal block rows / columns [L s, L s + L).  TMEM lane = query row, so a
// softmax thread still owns one whole score row - but the 32 rows of a warp now belong to up to three sequences, whose blocks start at
// different columns.  tcgen05.ld is warp-collective (one column window per warp): each warp loads a 64-column window that covers all of
// its blocks and every thread picks its L scores with compile-time offsets selected by its sequence index (SLOT and the warp's lane
// quadrant are template parameters: no dynamic register indexing).  The probabilities go to a block-diagonal bf16 [128 x 128] tile that
// is the K-major A operand of O = P V / dQ = dS K and the MN-major A operand of dV = P^T dO / dK = dS^T Q (attention_umma.cu).
//
// Backward, fused bias gradients of the q|k|v projection (the engine's colsum pass over d(qkv) [T, 3D] disappears):
//   d(b_v) = sum_j dV_j = sum_i (sum_j P'_ij) dO_i : the thread of query row i writes its post-dropout row sum into the dead key column
//            127 of P', so row 127 of dV = P'^T dO IS this group's d(b_v) - free on the tensor core;
//   d(b_q) = sum_i dQ_i : each thread adds its dQ rows into 64 registers over all groups of its CTA (a CTA stays on one head), one
//            shared-memory reduction + 64 atomics per CTA at the end;
//   d(b_k) = sum_j dK_j = sum_i q_i (sum_j dS_ij) = 0 exactly (softmax is shift invariant: every dS row sums to zero) - not formed; the
//            reference's autograd produces rounding noise there.
//
// PAIR = split precision ("bf16x3", the parity mode): q|k|v, dO and the outputs are (hi, lo) bf16 pairs. Every product X Y runs as the three
// tensor-core passes Xh Yh + Xl Yh + Xh Yl into one fp32 TMEM accumulator (the dropped Xl Yl term is 2^-16 relative), the probabilities
// and dS are split the same way when they are written to shared memory, so the [tile set] simply exists twice (hi set, lo set): forward
// 6 tiles (96 KB, 2 CTAs per SM), backward 14 tiles (224 KB, 1 CTA per SM). Replaces the fp32 SIMT kernels of attention.cu for L = 16 / 18
// (they took 370 of the 1306 ms of a parity-mode step). The bias-gradient fold stays off in this mode (the engine's colsum runs).
#include "common.cuh"
#include "../../include/clipdlm.h"
#include <cudaTypedefs.h>

namespace clipdlm {

int num_sms();
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, unsigned long long inner, unsigned long long outer, unsigned long long pitch_bytes,
                      uint32_t box_inner, uint32_t box_outer);

namespace {

constexpr int AP_DH = 64;
constexpr int AP_THREADS = 192;          // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 softmax / epilogue
constexpr uint32_t AP_TILE = 16384;      // [128 rows][64 bf16] = 128 rows x 128 B, 128-byte swizzle
constexpr float AP_LOG2E = 1.4426950408889634f;

// same counter layout as attention.cu / attention_umma.cu: element (i, j) of pair rh draws field ((j / 8) % 4) * 2 + j % 2 of the Philox block
// (rh, lane' = (i % 8) * 4 + (j % 8) / 2, q = (i / 8) * 4 + j / 32) - every attention kernel of the library sees the same masks.
__device__ __forceinline__ uint4 ap_rand_block(const DropoutCfg& d, unsigned long long rh, int lane_p, int q) {
  return philox4x32(make_uint4((uint32_t)rh, (uint32_t)(rh >> 32), (uint32_t)lane_p | ((uint32_t)q << 8), d.site ^ 0xa77e0000u),
                    make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
}
// keep bits (bit j = key j survives the dropout) of query row i, keys 0..SLOT-1 (SLOT <= 32)
template <int SLOT>
__device__ __forceinline__ uint32_t ap_row_keep_bits(const DropoutCfg& d, unsigned long long rh, int i) {
  uint32_t bits = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 rnd = ap_rand_block(d, rh, (i & 7) * 4 + c, (i >> 3) * 4);
    const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
      if (jb * 8 >= SLOT) continue;
      bits |= ((w[jb] & 0xffffu) >= d.thresh16 ? 1u : 0u) << (jb * 8 + 2 * c);
      bits |= ((w[jb] >> 16) >= d.thresh16 ? 1u : 0u) << (jb * 8 + 2 * c + 1);
    }
  }
  return bits;
}

struct AttPArgs {
  const uint32_t* keymask;
  __nv_bfloat16* out;       // ctx [T, D] (forward) or dqkv [T, 3D] (backward)
  __nv_bfloat16* out_lo;    // PAIR: the lo halves of the output
  float* dbias;             // backward: d(qkv bias) [3D] (fp32, accumulated with atomics) or nullptr
  int R, L, D, H;
  int tiles;                // ceil(R / NS): 128-row tiles per head
  DropoutCfg drop;
  float scale;
};

// Column window of lane quadrant Q (rows 32 Q .. 32 Q + 31 of the tile): the rows belong to sequence slots S0, S0 + 1, (S0 + 2);
// the warp loads TMEM columns [C0, C0 + 64), slot S0 + k starts D(k) columns into that window.
template <int SLOT, int Q>
struct ApWin {
  static constexpr int NS = 127 / SLOT;                 // sequences per tile (row / key 127 stays free for the d(b_v) trick)
  static constexpr int S0 = (32 * Q) / SLOT;
  static constexpr int C0 = (S0 * SLOT + 64 > 128) ? 64 : S0 * SLOT;
  static constexpr int D(int k) { return (S0 + k) * SLOT - C0; }
  static constexpr bool valid(int k) { return S0 + k < NS && (S0 + k) * SLOT <= 32 * Q + 31 && D(k) >= 0 && D(k) + SLOT <= 64; }
  static_assert(SLOT >= 16 && SLOT <= 32 && SLOT % 2 == 0, "slot size");
  static_assert(((32 * Q + 31) / SLOT) - S0 <= 2, "a warp's rows span at most three sequences");
};

// out[j] = win[D(k) + j - BASE] for the slot index k of this thread, for those j whose window column falls into [BASE, BASE + 32)
template <int SLOT, int Q, int BASE>
__device__ __forceinline__ void ap_pick(const float (&win)[32], int k, float (&out)[SLOT]) {
  using W = ApWin<SLOT, Q>;
#pragma unroll
  for (int j = 0; j < SLOT; ++j) {
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      if (!W::valid(kk)) continue;
      const int idx = W::D(kk) + j - BASE;
      if (idx >= 0 && idx < 32) out[j] = (k == kk) ? win[idx] : out[j];
    }
  }
}
// this thread's SLOT values of a [128 x 128] fp32 TMEM tile (lane = its row, columns = its sequence's block)
template <int SLOT, int Q>
__device__ __forceinline__ void ap_load_row(uint32_t tlane, int k, float (&out)[SLOT]) {
  using W = ApWin<SLOT, Q>;
  float win[32];
  tmem_ld32(tlane + W::C0, win);
  ap_pick<SLOT, Q, 0>(win, k, out);
  tmem_ld32(tlane + W::C0 + 32, win);
  ap_pick<SLOT, Q, 32>(win, k, out);
}

// 64 consecutive fp32 TMEM columns of this warp's 32 lanes, WITHOUT waiting: the caller issues several and waits once (tmem_ld_wait) - every
// tcgen05.wait::ld exposes a full TMEM round trip, and the softmax threads sit on exactly that chain.
__device__ __forceinline__ void tmem_ld64_nowait(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]),
        "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
        "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// out[j] = win[D(k) + j] for the slot index k of this thread (64-column window)
template <int SLOT, int Q>
__device__ __forceinline__ void ap_pick64(const uint32_t (&win)[64], int k, float (&out)[SLOT]) {
  using W = ApWin<SLOT, Q>;
#pragma unroll
  for (int j = 0; j < SLOT; ++j) {
    uint32_t v = 0u;
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      if (!W::valid(kk)) continue;
      v = (k == kk) ? win[W::D(kk) + j] : v;
    }
    out[j] = __uint_as_float(v);
  }
}
// S row and dP row of this thread with ONE TMEM round trip: both 64-column windows in flight, one wait
template <int SLOT, int Q>
__device__ __forceinline__ void ap_load_rows2(uint32_t tlane_s, uint32_t tlane_d, int k, float (&p)[SLOT], float (&dp)[SLOT]) {
  using W = ApWin<SLOT, Q>;
  uint32_t ws[64], wd[64];
  tmem_ld64_nowait(tlane_s + W::C0, ws);
  tmem_ld64_nowait(tlane_d + W::C0, wd);
  tmem_ld_wait();
  ap_pick64<SLOT, Q>(ws, k, p);
  ap_pick64<SLOT, Q>(wd, k, dp);
}
template <int SLOT, int Q>
__device__ __forceinline__ void ap_load_row64(uint32_t tlane, int k, float (&out)[SLOT]) {
  using W = ApWin<SLOT, Q>;
  uint32_t w[64];
  tmem_ld64_nowait(tlane + W::C0, w);
  tmem_ld_wait();
  ap_pick64<SLOT, Q>(w, k, out);
}

// SLOT fp32 values -> bf16 pairs at keys [SLOT s, SLOT s + SLOT) of row `row` of a K-major [2 chunks][128 rows][128 B] tile (128-byte swizzle)
template <int SLOT>
__device__ __forceinline__ void ap_store_block(uint32_t tile, int row, int s, const float (&x)[SLOT]) {
  const uint32_t rbase = tile + (uint32_t)row * 128u;
  const uint32_t w0 = (uint32_t)(SLOT / 2) * (uint32_t)s;
#pragma unroll
  for (int jj = 0; jj < SLOT / 2; ++jj) {
    const uint32_t w = w0 + jj;                      // 32-bit word index within the 128-key row
    const uint32_t addr = rbase + (w >> 5) * AP_TILE + ((((w & 31u) >> 2) ^ (uint32_t)(row & 7)) << 4) + ((w & 3u) << 2);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack_bf16x2(x[2 * jj], x[2 * jj + 1])) : "memory");
  }
}
// PAIR: x -> (bf16(x), bf16(x - bf16(x))) into the hi tile and its twin `lo_off` bytes further
template <int SLOT>
__device__ __forceinline__ void ap_store_block_pair(uint32_t tile, uint32_t lo_off, int row, int s, const float (&x)[SLOT]) {
  float lo[SLOT];
#pragma unroll
  for (int j = 0; j < SLOT; ++j) lo[j] = x[j] - bf16_round(x[j]);
  ap_store_block<SLOT>(tile, row, s, x);
  ap_store_block<SLOT>(tile + lo_off, row, s, lo);
}
__device__ __forceinline__ void ap_zero_row_chunk(uint32_t tile_chunk, int row) {   // 128 B of one row of one 64-key chunk
  // 16-byte units visited in swizzled order: the 32 rows of a warp are 128 B apart (the same banks), the XOR spreads each store over 8 bank
  // groups (4-way instead of 32-way conflicts)
#pragma unroll
  for (int c = 0; c < 8; ++c)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_chunk + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4)), "r"(0u) : "memory");
}

// 64 fp32 accumulator columns of this thread's TMEM lane -> 64 bf16 (128 contiguous bytes) in global memory; ACC: also added into acc[]
// 64 accumulator columns of this thread's TMEM lane with ONE tcgen05.ld (x64) and one wait -> 64 bf16 in global memory; ACC: also added into acc[].
// (10 warps per CTA = 3 warps on two of the four register-file partitions: 168 registers per thread is the hard cap, so the pipelined
// backward's epilogue takes its three outputs one after the other - three TMEM round trips instead of six.)
template <bool ACC, int NACC>
__device__ __forceinline__ void ap_store_out64x(uint32_t taddr, __nv_bfloat16* dst, bool store, bool accumulate, float (&acc)[NACC]) {
  static_assert(!ACC || NACC == 64, "accumulator size");
  uint32_t q[64];
  tmem_ld64_nowait(taddr, q);
  tmem_ld_wait();
  if (store) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(dst + c * 8) =
          make_uint4(pack_bf16x2(__uint_as_float(q[8 * c]), __uint_as_float(q[8 * c + 1])), pack_bf16x2(__uint_as_float(q[8 * c + 2]), __uint_as_float(q[8 * c + 3])),
                     pack_bf16x2(__uint_as_float(q[8 * c + 4]), __uint_as_float(q[8 * c + 5])), pack_bf16x2(__uint_as_float(q[8 * c + 6]), __uint_as_float(q[8 * c + 7])));
  }
  if (ACC) {
    if (accumulate) {
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j % NACC] += __uint_as_float(q[j]);
    }
  }
}

template <bool ACC, int NACC>
__device__ __forceinline__ void ap_store_out64(uint32_t taddr, __nv_bfloat16* dst, bool store, bool accumulate, float (&acc)[NACC],
                                               __nv_bfloat16* dst_lo = nullptr) {
  static_assert(!ACC || NACC == 64, "accumulator size");
  float v[32];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(taddr + half * 32, v);
    if (store) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(dst + half * 32 + c * 8) =
            make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]), pack_bf16x2(v[8 * c + 4], v[8 * c + 5]),
                       pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
      if (dst_lo != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] -= bf16_round(v[j]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(dst_lo + half * 32 + c * 8) =
              make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]), pack_bf16x2(v[8 * c + 4], v[8 * c + 5]),
                         pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
      }
    }
    if (ACC) {
      if (accumulate) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[(half * 32 + j) % NACC] += v[j];
      }
    }
  }
}

// CTAs per SM: backward 2 (smem, TMEM), forward 4 - or 3 (EVAL3): without dropout the softmax is short enough that three groups in flight
// cover the chain, and 96 instead of 80 registers take most of the L = 18 kernel's spills away (measured, same box, R = 8192 x 12 heads:
// p = 0: 177.5 us at 3 CTAs / 184.0 at 4; p = 0.1: 200.8 / 197.9 - so training keeps 4, the denoise loop takes 3).
template <bool BWD, bool PAIR, bool EVAL3>
constexpr int ap_ctas_per_sm() { return ((BWD ? 2 : (EVAL3 ? 3 : 4))) / (PAIR ? 2 : 1); }
template <bool BWD, int SLOT, bool PAIR, bool EVAL3 = false>
__global__ void __launch_bounds__(AP_THREADS, ap_ctas_per_sm<BWD, PAIR, EVAL3>())
attn_packed_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_qkv_lo,
                   const __grid_constant__ CUtensorMap tm_do_lo, const AttPArgs a) {
  extern __shared__ __align__(1024) uint8_t ap_smem[];
  // forward : Q | K | V, P (2 tiles) overlays Q | K once S is done      = 3 tiles (48 KB), 128 TMEM columns: 4 CTAs per SM
  // backward: Q | K | dO | V + P (P chunk 0 overlays V) | dS(2)           = 7 tiles (112 KB), 256 TMEM columns: 2 CTAs per SM
  constexpr int OFF_Q = 0, OFF_K = 1, OFF_DO = 2, OFF_V = BWD ? 3 : 2, OFF_P = BWD ? 3 : 0, OFF_DS = 5;
  constexpr int NSET = BWD ? 7 : 3;                         // tiles of one precision set
  constexpr int NTILES_SMEM = NSET * (PAIR ? 2 : 1);        // PAIR: the lo set follows the hi set, same layout
  constexpr uint32_t LO = NSET * AP_TILE;                   // byte offset hi tile -> its lo twin
  constexpr uint32_t TMEM_COLS = BWD ? 256 : 128;
  constexpr uint32_t COL_O = BWD ? 128 : 0;
  constexpr int NLOADS = (BWD ? 4 : 3) * (PAIR ? 2 : 1);
  constexpr int NS = 127 / SLOT;
  constexpr int ROWS = NS * SLOT;      // live rows of a tile (<= 127)
  uint8_t* smem = ap_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NTILES_SMEM * AP_TILE);
  uint64_t* in_full = bars + 0; uint64_t* in_empty = bars + 1; uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3; uint64_t* o_full = bars + 4; uint64_t* t_empty = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0u) { if (threadIdx.x == 0) printf("clipdlm: attention smem base not 1024-byte aligned\n"); __trap(); }

  // backward: the dS tile and key chunk 1 of P are dedicated; each row only ever writes its own diagonal block there (+ column 127 of P),
  // everything else stays zero for the life of the CTA. (Chunk 0 of P is where TMA lands V: its rows are rewritten for every group.)
  if (BWD) {
    const uint32_t zbytes = 3u * AP_TILE;
    for (int set = 0; set < (PAIR ? 2 : 1); ++set) {
      uint8_t* z0 = smem + set * LO + (OFF_P + 1) * AP_TILE;
      for (uint32_t off = threadIdx.x * 16; off < zbytes; off += AP_THREADS * 16) *reinterpret_cast<uint4*>(z0 + off) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    mbar_init(in_full, 1); mbar_init(in_empty, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(t_empty, 4);
    fence_mbar_init();
    tma_prefetch_desc(&tm_qkv);
    if (BWD) tma_prefetch_desc(&tm_do);
    if (PAIR) { tma_prefetch_desc(&tm_qkv_lo); if (BWD) tma_prefetch_desc(&tm_do_lo); }
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  // a CTA stays on ONE head (the d(b_q) register accumulators are per head): head = blockIdx.x % H, tiles blockIdx.x / H, + gridDim.x / H, ...
  const int h = (int)(blockIdx.x % a.H);
  const int tile0 = (int)(blockIdx.x / a.H), tstride = (int)(gridDim.x / a.H);
  const int my_tiles = a.tiles > tile0 ? (a.tiles - tile0 + tstride - 1) / tstride : 0;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // (the whole warp walks the loop, one elect.sync-elected lane issues: under `lane == 0` ptxas wraps every TMA / tcgen05 instruction in an
    //  ELECT / PLOP3 / BRA.U.ANY waterfall - ~9 instructions per MMA on the issuer's dependent chain, see gemm_tcgen05.cu)
    for (int n = 0; n < my_tiles; ++n) {
      const int row0 = (tile0 + n * tstride) * ROWS;   // first token row of the tile
      mbar_wait(in_empty, ((uint32_t)n & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(in_full, NLOADS * AP_TILE);   // full 128-row boxes: rows past the tensor arrive as zeros
        tma_load_2d(smem + OFF_Q * AP_TILE, &tm_qkv, in_full, h * AP_DH, row0);
        tma_load_2d(smem + OFF_K * AP_TILE, &tm_qkv, in_full, a.D + h * AP_DH, row0);
        tma_load_2d(smem + OFF_V * AP_TILE, &tm_qkv, in_full, 2 * a.D + h * AP_DH, row0);
        if (BWD) tma_load_2d(smem + OFF_DO * AP_TILE, &tm_do, in_full, h * AP_DH, row0);
        if (PAIR) {
          tma_load_2d(smem + LO + OFF_Q * AP_TILE, &tm_qkv_lo, in_full, h * AP_DH, row0);
          tma_load_2d(smem + LO + OFF_K * AP_TILE, &tm_qkv_lo, in_full, a.D + h * AP_DH, row0);
          tma_load_2d(smem + LO + OFF_V * AP_TILE, &tm_qkv_lo, in_full, 2 * a.D + h * AP_DH, row0);
          if (BWD) tma_load_2d(smem + LO + OFF_DO * AP_TILE, &tm_do_lo, in_full, h * AP_DH, row0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (same schedule as attention_umma.cu) =====================================
    // Whole warp in the loop; each [MMAs + commit] group is issued by one elected lane (a commit sits with the MMAs it tracks).
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);    // S, dP: both operands K-major (contraction over the head dim)
      constexpr uint32_t idesc_kv = make_idesc_bf16(128, 64, 0, 1);    // O = P V, dQ = dS K: A K-major, B MN-major
      constexpr uint32_t idesc_tv = make_idesc_bf16(128, 64, 1, 1);    // dV = P^T dO, dK = dS^T Q: both MN-major
      const uint32_t sq = smem_u32(smem + OFF_Q * AP_TILE), sk = smem_u32(smem + OFF_K * AP_TILE), sv = smem_u32(smem + OFF_V * AP_TILE);
      const uint32_t sdo = smem_u32(smem + OFF_DO * AP_TILE), sp = smem_u32(smem + OFF_P * AP_TILE), sds = smem_u32(smem + OFF_DS * AP_TILE);
      // X Y = Xh Yh (+ Xl Yh + Xh Yl in split precision): A / B descriptors of the hi tiles, the lo twins sit LO bytes further
      constexpr uint64_t LO_DESC = (uint64_t)(LO >> 4);
      auto mm = [&](uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
        umma_bf16(d, adesc, bdesc, idesc, acc);
        if (PAIR) {
          umma_bf16(d, adesc + LO_DESC, bdesc, idesc, 1u);
          umma_bf16(d, adesc, bdesc + LO_DESC, idesc, 1u);
        }
      };
      for (int n = 0; n < my_tiles; ++n) {
        const uint32_t par = (uint32_t)n & 1u;
        mbar_wait(in_full, par);
        mbar_wait(t_empty, par ^ 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)   // S = Q K^T
          mm(tmem_base, make_smem_desc_sw128(sq, 16, 1024) + (uint64_t)(k * 2), make_smem_desc_sw128(sk, 16, 1024) + (uint64_t)(k * 2),
             idesc_s, k > 0 ? 1u : 0u);
        if (BWD) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // dP = dO V^T
            mm(tmem_base + 128, make_smem_desc_sw128(sdo, 16, 1024) + (uint64_t)(k * 2),
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
            mm(tmem_base + COL_O, make_smem_desc_sw128(sp + (k >> 2) * AP_TILE, 16, 1024) + (uint64_t)((k & 3) * 2),
               make_smem_desc_sw128(sv, 8192, 1024) + (uint64_t)(k * 128), idesc_kv, k > 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dV = P^T dO   (queries in blocks of 16: rows 16 k of P and dO)
            mm(tmem_base, make_smem_desc_sw128(sp, AP_TILE, 1024) + (uint64_t)(k * 128),
               make_smem_desc_sw128(sdo, 8192, 1024) + (uint64_t)(k * 128), idesc_tv, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dK = dS^T Q
            mm(tmem_base + 64, make_smem_desc_sw128(sds, AP_TILE, 1024) + (uint64_t)(k * 128),
               make_smem_desc_sw128(sq, 8192, 1024) + (uint64_t)(k * 128), idesc_tv, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dQ = dS K
            mm(tmem_base + 128, make_smem_desc_sw128(sds + (k >> 2) * AP_TILE, 16, 1024) + (uint64_t)((k & 3) * 2),
               make_smem_desc_sw128(sk, 8192, 1024) + (uint64_t)(k * 128), idesc_kv, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full);     // (same elected lane as the MMAs above: tcgen05.commit tracks the issuing thread's operations)
        umma_commit(in_empty);   // every operand tile of this group has been consumed
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== softmax / epilogue =====================================
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;         // TMEM lane == row of the 128-row tile == token row0 + row
    const int s = row / SLOT;                 // sequence slot within the tile
    const int i = row - s * SLOT;             // query row (and key row for the dK / dV outputs)
    const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t sp = smem_u32(smem + OFF_P * AP_TILE), sds = smem_u32(smem + OFF_DS * AP_TILE);
    const int k_slot = s - (32 * quad) / SLOT;   // 0, 1 or 2: which block of the warp's column window is this thread's
    float acc[(BWD && !PAIR) ? 64 : 1];
#pragma unroll
    for (int j = 0; j < ((BWD && !PAIR) ? 64 : 1); ++j) acc[j] = 0.f;
    const bool dropping = a.drop.thresh16 != 0;
    const float sl = a.scale * AP_LOG2E;
    for (int n = 0; n < my_tiles; ++n) {
      const uint32_t par = (uint32_t)n & 1u;
      const int tile = tile0 + n * tstride;
      const int r = tile * NS + s;               // sequence
      const bool live = row < ROWS && r < a.R;
      const unsigned long long rh = (unsigned long long)r * a.H + h;
      const uint32_t keybits = live ? (a.keymask[r] & (SLOT >= 32 ? 0xffffffffu : ((1u << SLOT) - 1u))) : 0u;
      mbar_wait(s_full, par);
      tc_fence_after();
      float p[SLOT];
#pragma unroll
      for (int j = 0; j < SLOT; ++j) p[j] = 0.f;
      if constexpr (BWD && !PAIR) {   // (2 CTAs x 192 threads: 170 registers - room for the 64-register window; forward / split precision keep the 32-register one)
        switch (quad) {
          case 0: ap_load_row64<SLOT, 0>(tq, k_slot, p); break;
          case 1: ap_load_row64<SLOT, 1>(tq, k_slot, p); break;
          case 2: ap_load_row64<SLOT, 2>(tq, k_slot, p); break;
          default: ap_load_row64<SLOT, 3>(tq, k_slot, p); break;
        }
      } else {
        switch (quad) {
          case 0: ap_load_row<SLOT, 0>(tq, k_slot, p); break;
          case 1: ap_load_row<SLOT, 1>(tq, k_slot, p); break;
          case 2: ap_load_row<SLOT, 2>(tq, k_slot, p); break;
          default: ap_load_row<SLOT, 3>(tq, k_slot, p); break;
        }
      }
      {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < SLOT; ++j) {
          p[j] = (keybits & (1u << j)) ? p[j] * sl : -INFINITY;
          mx = fmaxf(mx, p[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < SLOT; ++j) { p[j] = ex2_ftz(p[j] - mx); sum += p[j]; }   // a fully masked row gives NaN, as the reference does
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < SLOT; ++j) p[j] = live ? p[j] * inv : 0.f;
      }
      const uint32_t kb = dropping ? ap_row_keep_bits<SLOT>(a.drop, rh, i) : 0xffffffffu;
      if (!BWD) {
        if (dropping) {
#pragma unroll
          for (int j = 0; j < SLOT; ++j) p[j] = (kb & (1u << j)) ? p[j] * a.drop.scale : 0.f;
        }
        // P lives where TMA landed Q / K: rewrite this row of both key chunks completely (zeros off the diagonal block)
        ap_zero_row_chunk(sp, row);
        ap_zero_row_chunk(sp + AP_TILE, row);
        if (PAIR) {
          ap_zero_row_chunk(sp + LO, row);
          ap_zero_row_chunk(sp + LO + AP_TILE, row);
          if (live) ap_store_block_pair<SLOT>(sp, LO, row, s, p);
        } else if (live) {
          ap_store_block<SLOT>(sp, row, s, p);
        }
      } else {
        float dp[SLOT];
#pragma unroll
        for (int j = 0; j < SLOT; ++j) dp[j] = 0.f;
        switch (quad) {
          case 0: ap_load_row<SLOT, 0>(tq + 128, k_slot, dp); break;
          case 1: ap_load_row<SLOT, 1>(tq + 128, k_slot, dp); break;
          case 2: ap_load_row<SLOT, 2>(tq + 128, k_slot, dp); break;
          default: ap_load_row<SLOT, 3>(tq + 128, k_slot, dp); break;
        }
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < SLOT; ++j) {
          if (dropping) dp[j] = (kb & (1u << j)) ? dp[j] * a.drop.scale : 0.f;   // gradient through the dropout
          dot = fmaf(dp[j], p[j], dot);
        }
        float rho = 0.f;   // post-dropout row sum of P: the coefficient of dO_i in d(b_v)
#pragma unroll
        for (int j = 0; j < SLOT; ++j) {
          dp[j] = live ? p[j] * (dp[j] - dot) * a.scale : 0.f;                      // dS (p == 0 on masked keys)
          if (dropping) p[j] = (kb & (1u << j)) ? p[j] * a.drop.scale : 0.f;       // dropped probabilities: what multiplied V in the forward
          rho += p[j];
        }
        // key chunk 0 of P is where TMA landed V: rewrite this row of it completely
        ap_zero_row_chunk(sp, row);
        if (PAIR) ap_zero_row_chunk(sp + LO, row);
        if (PAIR) {
          if (row < ROWS) {
            ap_store_block_pair<SLOT>(sp, LO, row, s, p);
            ap_store_block_pair<SLOT>(sds, LO, row, s, dp);
          }
        } else if (row < ROWS) {
          ap_store_block<SLOT>(sp, row, s, p);
          ap_store_block<SLOT>(sds, row, s, dp);
          if (a.dbias != nullptr) {   // P'[i, 127] = rho_i  (word 63 of the row: keys 126 | 127; key 126 is dead as well)
            const uint32_t addr = sp + AP_TILE + (uint32_t)row * 128u + ((7u ^ (uint32_t)(row & 7)) << 4) + 12u;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack_bf16x2(0.f, rho)) : "memory");
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P / dS (generic proxy) -> tcgen05.mma operand reads (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      mbar_wait(o_full, par);
      tc_fence_after();
      const size_t tok = (size_t)tile * ROWS + row;   // == r * L + i on live rows
      if constexpr (!BWD) {
        ap_store_out64<false>(tq + COL_O, a.out + tok * a.D + h * AP_DH, live, false, acc, PAIR ? a.out_lo + tok * a.D + h * AP_DH : nullptr);
      } else if constexpr (PAIR) {
        __nv_bfloat16* o = a.out + tok * 3 * a.D + h * AP_DH;
        __nv_bfloat16* ol = a.out_lo + tok * 3 * a.D + h * AP_DH;
        ap_store_out64<false>(tq + 128, o, live, false, acc, ol);                      // dQ
        ap_store_out64<false>(tq + 64, o + a.D, live, false, acc, ol + a.D);           // dK (row = key i)
        ap_store_out64<false>(tq, o + 2 * a.D, live, false, acc, ol + 2 * a.D);        // dV
      } else {
        __nv_bfloat16* o = a.out + tok * 3 * a.D + h * AP_DH;
        const bool fold = a.dbias != nullptr;
        ap_store_out64<true>(tq + 128, o, live, fold && live, acc);                    // dQ (+ this thread's share of d(b_q))
        ap_store_out64<false>(tq + 64, o + a.D, live, false, acc);                     // dK (row = key i)
        ap_store_out64<true>(tq, o + 2 * a.D, live, fold && row == 127, acc);          // dV; row 127 of dV is the group's d(b_v)
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty);
    }
    if constexpr (BWD && !PAIR) if (a.dbias != nullptr && my_tiles > 0) {
      // d(b_q)[h] = sum over the 127 query lanes of acc; d(b_v)[h] = lane 127's acc.  Every MMA of this CTA has completed (the last o_full
      // wait), so the operand tiles are free: [128 rows][64] fp32 staging over Q | K.
      float* stage = reinterpret_cast<float*>(smem);
#pragma unroll
      for (int j = 0; j < 64; ++j) stage[row * 64 + ((j + row) & 63)] = acc[j];   // rotated columns: conflict-free writes, and reads below
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int t = threadIdx.x - 64;    // 0..127
      if (t < 64) {
        float sq = 0.f;
        for (int rr = 0; rr < 127; ++rr) sq += stage[rr * 64 + ((t + rr) & 63)];
        atomicAdd(a.dbias + h * AP_DH + t, sq);
      } else {
        const int c = t - 64;
        atomicAdd(a.dbias + 2 * a.D + h * AP_DH + c, stage[127 * 64 + ((c + 127) & 63)]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------------------------------
// Software-pipelined backward (plain bf16): ONE CTA per SM that keeps TWO groups in flight in two smem / TMEM slots, with the per-group
// work split over two warp sets that run concurrently on different groups -
//   warps 2..5 "softmax": S, dP (TMEM) -> P', dS (smem)          of group n
//   warps 6..9 "epilogue": dQ, dK, dV (TMEM) -> global, bias sums  of group n - 1
// while the MMA warp alternates [S, dP of n] / [dV, dK, dQ of n - 1] and the TMA warp loads group n + 1. attn_packed_kernel<true> runs these
// phases back to back in one warp set (two CTAs per SM give two chains of ~10 us per group); here a group leaves the SM every
// max(softmax, epilogue) time. Same arithmetic, same smem tiles (one 7-tile set per slot), same folded bias gradients.
// ------------------------------------------------------------------------------------------------------------------------------------
constexpr int APP_THREADS = 320;   // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 softmax, warps 6..9 epilogue

template <int SLOT>
__global__ void __launch_bounds__(APP_THREADS, 1) attn_packed_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                                                                               const AttPArgs a) {
  extern __shared__ __align__(1024) uint8_t ap_smem[];
  constexpr int OFF_Q = 0, OFF_K = 1, OFF_DO = 2, OFF_V = 3, OFF_P = 3, OFF_DS = 5, NSET = 7;
  constexpr uint32_t SLOT_BYTES = NSET * AP_TILE;      // 112 KB per group slot
  constexpr int NS = 127 / SLOT;
  constexpr int ROWS = NS * SLOT;
  uint8_t* smem = ap_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * SLOT_BYTES);
  uint64_t* in_full = bars + 0; uint64_t* in_empty = bars + 2; uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 6; uint64_t* o_full = bars + 8; uint64_t* t_empty = bars + 10;      // [2] each: one per slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0u) { if (threadIdx.x == 0) printf("clipdlm: attention smem base not 1024-byte aligned\n"); __trap(); }
  for (int slot = 0; slot < 2; ++slot) {   // dedicated tiles (P chunk 1, dS): zero for the life of the CTA outside the rows' own blocks
    uint8_t* z0 = smem + slot * SLOT_BYTES + (OFF_P + 1) * AP_TILE;
    for (uint32_t off = threadIdx.x * 16; off < 3u * AP_TILE; off += APP_THREADS * 16) *reinterpret_cast<uint4*>(z0 + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    for (int s2 = 0; s2 < 2; ++s2) {
      mbar_init(in_full + s2, 1); mbar_init(in_empty + s2, 1); mbar_init(s_full + s2, 1);
      mbar_init(p_full + s2, 4); mbar_init(o_full + s2, 1); mbar_init(t_empty + s2, 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int h = (int)(blockIdx.x % a.H);
  const int tile0 = (int)(blockIdx.x / a.H), tstride = (int)(gridDim.x / a.H);
  const int my_tiles = a.tiles > tile0 ? (a.tiles - tile0 + tstride - 1) / tstride : 0;

  if (warp == 0) {
    // ===================================== TMA producer (whole warp in the loop, one elected lane issues) =====================================
    for (int n = 0; n < my_tiles; ++n) {
      const int sl = n & 1;
      const uint32_t ph = (uint32_t)(n >> 1) & 1u;
      uint8_t* base = smem + sl * SLOT_BYTES;
      const int row0 = (tile0 + n * tstride) * ROWS;
      mbar_wait(in_empty + sl, ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(in_full + sl, 4 * AP_TILE);
        tma_load_2d(base + OFF_Q * AP_TILE, &tm_qkv, in_full + sl, h * AP_DH, row0);
        tma_load_2d(base + OFF_K * AP_TILE, &tm_qkv, in_full + sl, a.D + h * AP_DH, row0);
        tma_load_2d(base + OFF_V * AP_TILE, &tm_qkv, in_full + sl, 2 * a.D + h * AP_DH, row0);
        tma_load_2d(base + OFF_DO * AP_TILE, &tm_do, in_full + sl, h * AP_DH, row0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer: [S, dP of n] then [dV, dK, dQ of n - 1] =====================================
    // (whole warp in the loop; each [MMAs + commits] group goes out from one elect.sync-elected lane, no per-instruction waterfall)
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_kv = make_idesc_bf16(128, 64, 0, 1);
      constexpr uint32_t idesc_tv = make_idesc_bf16(128, 64, 1, 1);
      for (int n = 0; n <= my_tiles; ++n) {
        if (n < my_tiles) {
          const int sl = n & 1;
          const uint32_t ph = (uint32_t)(n >> 1) & 1u;
          const uint32_t b0 = smem_u32(smem + sl * SLOT_BYTES), tb = tmem_base + (uint32_t)sl * 256u;
          const uint32_t sq = b0 + OFF_Q * AP_TILE, sk = b0 + OFF_K * AP_TILE, sv = b0 + OFF_V * AP_TILE, sdo = b0 + OFF_DO * AP_TILE;
          mbar_wait(in_full + sl, ph);
          mbar_wait(t_empty + sl, ph ^ 1u);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // S = Q K^T
            umma_bf16(tb, make_smem_desc_sw128(sq, 16, 1024) + (uint64_t)(k * 2), make_smem_desc_sw128(sk, 16, 1024) + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // dP = dO V^T
            umma_bf16(tb + 128, make_smem_desc_sw128(sdo, 16, 1024) + (uint64_t)(k * 2), make_smem_desc_sw128(sv, 16, 1024) + (uint64_t)(k * 2), idesc_s,
                      k > 0 ? 1u : 0u);
          umma_commit(s_full + sl);
          }
          __syncwarp();
        }
        if (n >= 1) {
          const int m = n - 1, sl = m & 1;
          const uint32_t ph = (uint32_t)(m >> 1) & 1u;
          const uint32_t b0 = smem_u32(smem + sl * SLOT_BYTES), tb = tmem_base + (uint32_t)sl * 256u;
          const uint32_t sq = b0 + OFF_Q * AP_TILE, sk = b0 + OFF_K * AP_TILE, sdo = b0 + OFF_DO * AP_TILE, sp = b0 + OFF_P * AP_TILE, sds = b0 + OFF_DS * AP_TILE;
          mbar_wait(p_full + sl, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dV = P^T dO
            umma_bf16(tb, make_smem_desc_sw128(sp, AP_TILE, 1024) + (uint64_t)(k * 128), make_smem_desc_sw128(sdo, 8192, 1024) + (uint64_t)(k * 128), idesc_tv,
                      k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dK = dS^T Q
            umma_bf16(tb + 64, make_smem_desc_sw128(sds, AP_TILE, 1024) + (uint64_t)(k * 128), make_smem_desc_sw128(sq, 8192, 1024) + (uint64_t)(k * 128), idesc_tv,
                      k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dQ = dS K
            umma_bf16(tb + 128, make_smem_desc_sw128(sds + (k >> 2) * AP_TILE, 16, 1024) + (uint64_t)((k & 3) * 2),
                      make_smem_desc_sw128(sk, 8192, 1024) + (uint64_t)(k * 128), idesc_kv, k > 0 ? 1u : 0u);
          umma_commit(o_full + sl);
          umma_commit(in_empty + sl);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int s = row / SLOT;
    const int i = row - s * SLOT;
    const bool dropping = a.drop.thresh16 != 0;
    if (warp < 6) {
      // ===================================== softmax warps: S, dP -> P', dS =====================================
      const int k_slot = s - (32 * quad) / SLOT;
      const float sl2 = a.scale * AP_LOG2E;
      for (int n = 0; n < my_tiles; ++n) {
        const int sl = n & 1;
        const uint32_t ph = (uint32_t)(n >> 1) & 1u;
        const uint32_t tq = tmem_base + (uint32_t)sl * 256u + ((uint32_t)(quad * 32) << 16);
        const uint32_t b0 = smem_u32(smem + sl * SLOT_BYTES);
        const uint32_t sp = b0 + OFF_P * AP_TILE, sds = b0 + OFF_DS * AP_TILE;
        const int tile = tile0 + n * tstride;
        const int r = tile * NS + s;
        const bool live = row < ROWS && r < a.R;
        const unsigned long long rh = (unsigned long long)r * a.H + h;
        const uint32_t keybits = live ? (a.keymask[r] & (SLOT >= 32 ? 0xffffffffu : ((1u << SLOT) - 1u))) : 0u;
        mbar_wait(s_full + sl, ph);
        tc_fence_after();
        float p[SLOT], dp[SLOT];
#pragma unroll
        for (int j = 0; j < SLOT; ++j) { p[j] = 0.f; dp[j] = 0.f; }
        // S row (one x64 load), then the dP window is requested BEFORE the softmax arithmetic on p and awaited after it: its TMEM round trip hides
        // behind the exponentials
        uint32_t wd[64];
        switch (quad) {
          case 0: ap_load_row64<SLOT, 0>(tq, k_slot, p); tmem_ld64_nowait(tq + 128 + ApWin<SLOT, 0>::C0, wd); break;
          case 1: ap_load_row64<SLOT, 1>(tq, k_slot, p); tmem_ld64_nowait(tq + 128 + ApWin<SLOT, 1>::C0, wd); break;
          case 2: ap_load_row64<SLOT, 2>(tq, k_slot, p); tmem_ld64_nowait(tq + 128 + ApWin<SLOT, 2>::C0, wd); break;
          default: ap_load_row64<SLOT, 3>(tq, k_slot, p); tmem_ld64_nowait(tq + 128 + ApWin<SLOT, 3>::C0, wd); break;
        }
        {
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < SLOT; ++j) {
            p[j] = (keybits & (1u << j)) ? p[j] * sl2 : -INFINITY;
            mx = fmaxf(mx, p[j]);
          }
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < SLOT; ++j) { p[j] = ex2_ftz(p[j] - mx); sum += p[j]; }
          const float inv = 1.f / sum;
#pragma unroll
          for (int j = 0; j < SLOT; ++j) p[j] = live ? p[j] * inv : 0.f;
        }
        const uint32_t kb = dropping ? ap_row_keep_bits<SLOT>(a.drop, rh, i) : 0xffffffffu;
        tmem_ld_wait();
        switch (quad) {
          case 0: ap_pick64<SLOT, 0>(wd, k_slot, dp); break;
          case 1: ap_pick64<SLOT, 1>(wd, k_slot, dp); break;
          case 2: ap_pick64<SLOT, 2>(wd, k_slot, dp); break;
          default: ap_pick64<SLOT, 3>(wd, k_slot, dp); break;
        }
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < SLOT; ++j) {
          if (dropping) dp[j] = (kb & (1u << j)) ? dp[j] * a.drop.scale : 0.f;
          dot = fmaf(dp[j], p[j], dot);
        }
        float rho = 0.f;
#pragma unroll
        for (int j = 0; j < SLOT; ++j) {
          dp[j] = live ? p[j] * (dp[j] - dot) * a.scale : 0.f;
          if (dropping) p[j] = (kb & (1u << j)) ? p[j] * a.drop.scale : 0.f;
          rho += p[j];
        }
        ap_zero_row_chunk(sp, row);   // key chunk 0 of P is where TMA landed V
        if (row < ROWS) {
          ap_store_block<SLOT>(sp, row, s, p);
          ap_store_block<SLOT>(sds, row, s, dp);
          if (a.dbias != nullptr) {
            const uint32_t addr = sp + AP_TILE + (uint32_t)row * 128u + ((7u ^ (uint32_t)(row & 7)) << 4) + 12u;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack_bf16x2(0.f, rho)) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + sl);
      }
    } else {
      // ===================================== epilogue warps: dQ, dK, dV -> global (+ bias sums) =====================================
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      const bool fold = a.dbias != nullptr;
      for (int n = 0; n < my_tiles; ++n) {
        const int sl = n & 1;
        const uint32_t ph = (uint32_t)(n >> 1) & 1u;
        const uint32_t tq = tmem_base + (uint32_t)sl * 256u + ((uint32_t)(quad * 32) << 16);
        const int tile = tile0 + n * tstride;
        const int r = tile * NS + s;
        const bool live = row < ROWS && r < a.R;
        mbar_wait(o_full + sl, ph);
        tc_fence_after();
        const size_t tok = (size_t)tile * ROWS + row;
        __nv_bfloat16* o = a.out + tok * 3 * a.D + h * AP_DH;
        ap_store_out64x<true>(tq + 128, o, live, fold && live, acc);                    // dQ (+ this thread's share of d(b_q))
        ap_store_out64x<false>(tq + 64, o + a.D, live, false, acc);                     // dK
        ap_store_out64x<true>(tq, o + 2 * a.D, live, fold && row == 127, acc);          // dV; row 127 of dV is the group's d(b_v)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + sl);
      }
      if (fold && my_tiles > 0) {
        float* stage = reinterpret_cast<float*>(smem);   // every MMA of this CTA has completed: [128 rows][64] fp32 over slot 0's Q | K
#pragma unroll
        for (int j = 0; j < 64; ++j) stage[row * 64 + ((j + row) & 63)] = acc[j];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int t = threadIdx.x - 192;    // 0..127
        if (t < 64) {
          float sq2 = 0.f;
          for (int rr = 0; rr < 127; ++rr) sq2 += stage[rr * 64 + ((t + rr) & 63)];
          atomicAdd(a.dbias + h * AP_DH + t, sq2);
        } else {
          const int c = t - 64;
          atomicAdd(a.dbias + 2 * a.D + h * AP_DH + c, stage[127 * 64 + ((c + 127) & 63)]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int g_packed_bwd_pipe = 1;   // 1 = software-pipelined backward (default), 0 = attn_packed_kernel<true> (two CTAs per SM)

template <int SLOT>
int launch_packed_bwd_pipe(const __nv_bfloat16* qkv, const __nv_bfloat16* dctx, const uint32_t* keymask, int R, int D, int H, __nv_bfloat16* out,
                           float* dbias, const DropoutCfg& drop, cudaStream_t st) {
  constexpr int NS = 127 / SLOT;
  const unsigned long long T = (unsigned long long)R * SLOT;
  CUtensorMap tm_qkv, tm_do;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tm_qkv, qkv, 3ull * D, T, 3ull * D * 2, AP_DH, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_do, dctx, (unsigned long long)D, T, (unsigned long long)D * 2, AP_DH, 128))) return rc;
  AttPArgs a;
  a.keymask = keymask; a.out = out; a.out_lo = nullptr; a.dbias = dbias; a.R = R; a.L = SLOT; a.D = D; a.H = H; a.drop = drop; a.scale = 0.125f;
  a.tiles = (R + NS - 1) / NS;
  const size_t smem = (size_t)14 * AP_TILE + 128;
  static bool attr_set = false;
  if (!attr_set) {
    CLIPDLM_CUDA_OK(cudaFuncSetAttribute(attn_packed_bwd_pipe_kernel<SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  long long per_head = num_sms() / H;   // one CTA per SM, every CTA on one head
  if (per_head < 1) per_head = 1;
  if (per_head > a.tiles) per_head = a.tiles;
  const int grid = (int)(per_head * H);
  CLIPDLM_CUDA_OK(launch_pdl(attn_packed_bwd_pipe_kernel<SLOT>, dim3(grid), dim3(APP_THREADS), smem, st, tm_qkv, tm_do, a));
  return 0;
}

template <bool BWD, int SLOT, bool PAIR, bool EVAL3 = false>
int launch_packed(const __nv_bfloat16* qkv, const __nv_bfloat16* qkv_lo, const __nv_bfloat16* dctx, const __nv_bfloat16* dctx_lo, const uint32_t* keymask,
                  int R, int D, int H, __nv_bfloat16* out, __nv_bfloat16* out_lo, float* dbias, const DropoutCfg& drop, cudaStream_t st) {
  constexpr int NS = 127 / SLOT;
  const unsigned long long T = (unsigned long long)R * SLOT;
  CUtensorMap tm_qkv, tm_do, tm_qkv_lo, tm_do_lo;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tm_qkv, qkv, 3ull * D, T, 3ull * D * 2, AP_DH, 128))) return rc;
  tm_do = tm_qkv; tm_qkv_lo = tm_qkv; tm_do_lo = tm_qkv;
  if (BWD && (rc = make_tmap_2d_bf16(&tm_do, dctx, (unsigned long long)D, T, (unsigned long long)D * 2, AP_DH, 128))) return rc;
  if (PAIR) {
    if ((rc = make_tmap_2d_bf16(&tm_qkv_lo, qkv_lo, 3ull * D, T, 3ull * D * 2, AP_DH, 128))) return rc;
    tm_do_lo = tm_qkv_lo;
    if (BWD && (rc = make_tmap_2d_bf16(&tm_do_lo, dctx_lo, (unsigned long long)D, T, (unsigned long long)D * 2, AP_DH, 128))) return rc;
  }
  AttPArgs a;
  a.keymask = keymask; a.out = out; a.out_lo = out_lo; a.dbias = PAIR ? nullptr : dbias; a.R = R; a.L = SLOT; a.D = D; a.H = H; a.drop = drop;
  a.scale = 0.125f;  // 1 / sqrt(64)
  a.tiles = (R + NS - 1) / NS;
  const size_t smem = (size_t)(BWD ? 7 : 3) * (PAIR ? 2 : 1) * AP_TILE + 64;
  static bool attr_set = false;
  if (!attr_set) {
    CLIPDLM_CUDA_OK(cudaFuncSetAttribute(attn_packed_kernel<BWD, SLOT, PAIR, EVAL3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  // grid: a multiple of H (every CTA keeps one head), at most ap_ctas_per_sm() CTAs per SM (half that in split precision), at most one CTA per (tile, head)
  long long per_head = ((long long)ap_ctas_per_sm<BWD, PAIR, EVAL3>() * num_sms()) / H;
  if (per_head < 1) per_head = 1;
  if (per_head > a.tiles) per_head = a.tiles;
  const int grid = (int)(per_head * H);
  CLIPDLM_CUDA_OK(launch_pdl(attn_packed_kernel<BWD, SLOT, PAIR, EVAL3>, dim3(grid), dim3(AP_THREADS), smem, st, tm_qkv, tm_do, tm_qkv_lo, tm_do_lo, a));
  return 0;
}

}  // namespace

bool attn_packed_supported(int L, int D, int H) { return (L == 16 || L == 18) && D == H * AP_DH && H <= 64; }
void attn_packed_bwd_pipeline(int on) { g_packed_bwd_pipe = on; }

// forward: dctx / dbias unused.  backward: dbias (nullable, plain bf16 only) receives d(q bias) and d(v bias) (+= ; d(k bias) = 0 is not touched).
// qkv_lo != nullptr selects split precision (dctx_lo / out_lo are then required as well).
template <bool BWD>
int launch_attn_packed(const __nv_bfloat16* qkv, const __nv_bfloat16* qkv_lo, const __nv_bfloat16* dctx, const __nv_bfloat16* dctx_lo,
                       const uint32_t* keymask, int R, int L, int D, int H, __nv_bfloat16* out, __nv_bfloat16* out_lo, float* dbias,
                       const DropoutCfg& drop, cudaStream_t st) {
  CLIPDLM_CHECK(attn_packed_supported(L, D, H), "packed attention: unsupported shape L %d D %d H %d", L, D, H);
  const bool pair = qkv_lo != nullptr;
  CLIPDLM_CHECK(!pair || (out_lo != nullptr && (!BWD || dctx_lo != nullptr)), "packed attention: split precision needs every lo operand");
  if (pair) {
    if (L == 16) return launch_packed<BWD, 16, true>(qkv, qkv_lo, dctx, dctx_lo, keymask, R, D, H, out, out_lo, dbias, drop, st);
    return launch_packed<BWD, 18, true>(qkv, qkv_lo, dctx, dctx_lo, keymask, R, D, H, out, out_lo, dbias, drop, st);
  }
  if (BWD && g_packed_bwd_pipe && H <= num_sms()) {
    if (L == 16) return launch_packed_bwd_pipe<16>(qkv, dctx, keymask, R, D, H, out, dbias, drop, st);
    return launch_packed_bwd_pipe<18>(qkv, dctx, keymask, R, D, H, out, dbias, drop, st);
  }
  if (!BWD && drop.thresh16 == 0) {   // eval / denoise forward: three CTAs per SM
    if (L == 16) return launch_packed<false, 16, false, true>(qkv, nullptr, dctx, nullptr, keymask, R, D, H, out, nullptr, dbias, drop, st);
    return launch_packed<false, 18, false, true>(qkv, nullptr, dctx, nullptr, keymask, R, D, H, out, nullptr, dbias, drop, st);
  }
  if (L == 16) return launch_packed<BWD, 16, false>(qkv, nullptr, dctx, nullptr, keymask, R, D, H, out, nullptr, dbias, drop, st);
  return launch_packed<BWD, 18, false>(qkv, nullptr, dctx, nullptr, keymask, R, D, H, out, nullptr, dbias, drop, st);
}
template int launch_attn_packed<false>(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const uint32_t*, int, int, int,
                                       int, __nv_bfloat16*, __nv_bfloat16*, float*, const DropoutCfg&, cudaStream_t);
template int launch_attn_packed<true>(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const uint32_t*, int, int, int,
                                      int, __nv_bfloat16*, __nv_bfloat16*, float*, const DropoutCfg&, cudaStream_t);

}  // namespace clipdlm
