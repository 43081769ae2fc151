// Shared device helpers for the clipdlm sm_100a kernels: PTX wrappers (mbarrier, TMA, tcgen05/TMEM),
// bf16 / split-bf16 ("bf16x3") storage policies, counter-based dropout RNG, reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace clipdlm {

// ------------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define CLIPDLM_CUDA_OK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      clipdlm::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)
#define CLIPDLM_CHECK(cond, ...)                                                                \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      clipdlm::set_last_error(__VA_ARGS__);                                                     \
      return -2;                                                                                \
    }                                                                                           \
  } while (0)

// ------------------------------------------------------------------------------------------------
// small device utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// One lane of a converged warp (elect.sync always names the same lane for the same member mask): ptxas keeps the operands of single-thread
// instructions (tcgen05.mma / commit, TMA) on the uniform datapath under this predicate, where `lane == 0` costs an ELECT / R2UR waterfall each.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// bf16 pack / unpack -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Split ("bf16x3") storage: a value is kept as hi = bf16(x), lo = bf16(x - hi); hi + lo carries ~16
// mantissa bits. Tensors in this mode are two arrays of identical shape. lo == nullptr => plain bf16.
struct BfPtr {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
};
struct CBfPtr {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;
};

// Load 8 consecutive elements (16 B) as floats.
__device__ __forceinline__ void load8(const CBfPtr& p, size_t idx, float (&v)[8]) {
  uint4 h = *reinterpret_cast<const uint4*>(p.hi + idx);
  float2 a = unpack_bf16x2(h.x), b = unpack_bf16x2(h.y), c = unpack_bf16x2(h.z), d = unpack_bf16x2(h.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  if (p.lo != nullptr) {
    uint4 l = *reinterpret_cast<const uint4*>(p.lo + idx);
    a = unpack_bf16x2(l.x); b = unpack_bf16x2(l.y); c = unpack_bf16x2(l.z); d = unpack_bf16x2(l.w);
    v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
  }
}
__device__ __forceinline__ void store8(const BfPtr& p, size_t idx, const float (&v)[8]) {
  uint4 h;
  h.x = pack_bf16x2(v[0], v[1]); h.y = pack_bf16x2(v[2], v[3]);
  h.z = pack_bf16x2(v[4], v[5]); h.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p.hi + idx) = h;
  if (p.lo != nullptr) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i] - bf16_round(v[i]);
    uint4 l;
    l.x = pack_bf16x2(r[0], r[1]); l.y = pack_bf16x2(r[2], r[3]);
    l.z = pack_bf16x2(r[4], r[5]); l.w = pack_bf16x2(r[6], r[7]);
    *reinterpret_cast<uint4*>(p.lo + idx) = l;
  }
}
__device__ __forceinline__ float load1(const CBfPtr& p, size_t idx) {
  float v = __bfloat162float(p.hi[idx]);
  if (p.lo != nullptr) v += __bfloat162float(p.lo[idx]);
  return v;
}
__device__ __forceinline__ void store1(const BfPtr& p, size_t idx, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  p.hi[idx] = h;
  if (p.lo != nullptr) p.lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ------------------------------------------------------------------------------------------------
// Counter-based RNG for dropout: Philox4x32 keyed by (seed), counter = (element index / 8, site).
// 7 rounds: the smallest round count of Philox4x32 that passes BigCrush (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
// SC'11, table 2; 10 is the library default with safety margin) - for dropout masks Crush-resistance is ample, and the mask draw is a
// measurable share of the attention / LayerNorm / FFN-epilogue instruction streams (L = 66 attention forward 309 -> 277 us, backward 719 -> 688 us per launch).
// One call yields 128 bits = 8 x 16-bit lanes; element e keeps iff lane(e % 8) >= thresh16.
// The backward pass regenerates the same mask from (seed, site, index): no mask tensor in HBM.
// ------------------------------------------------------------------------------------------------
constexpr int PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
struct DropoutCfg {
  unsigned long long seed;  // per-step seed
  uint32_t site;            // which dropout site (layer * 4 + kind)
  uint32_t thresh16;        // round(p * 65536); 0 => dropout off
  float scale;              // 1 / (1 - p)
};
// Keep-mask for the 8 elements [8*g, 8*g+8) of a site: bit i set => keep element 8*g+i.
__device__ __forceinline__ uint32_t dropout_keep8(const DropoutCfg& d, unsigned long long g) {
  uint4 r = philox4x32(make_uint4((uint32_t)g, (uint32_t)(g >> 32), d.site, 0x51ed2701u),
                          make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
  uint32_t m = 0;
  m |= ((r.x & 0xffffu) >= d.thresh16) << 0; m |= ((r.x >> 16) >= d.thresh16) << 1;
  m |= ((r.y & 0xffffu) >= d.thresh16) << 2; m |= ((r.y >> 16) >= d.thresh16) << 3;
  m |= ((r.z & 0xffffu) >= d.thresh16) << 4; m |= ((r.z >> 16) >= d.thresh16) << 5;
  m |= ((r.w & 0xffffu) >= d.thresh16) << 6; m |= ((r.w >> 16) >= d.thresh16) << 7;
  return m;
}

// exact-erf GELU (HF activations.py GELUActivation -> nn.functional.gelu) and its derivative, built for the GEMM epilogues
// (8 warps have to finish 128 x 256 outputs inside one K = 768 mainloop, ~6 k cycles): the normal tail
//   h(a) = Phi(-a) = 0.5 erfc(a / sqrt 2),  a = |x|,  is evaluated as 2^(-p(a)) with p a degree-6 polynomial (p(0) = 1, minimax
// fit of -log2 h on [0, 5.6] weighted by h; |abs error of Phi| <= 1.9e-7 in fp32, i.e. round-off level - the A&S 7.1.26 form
// used before has 1.5e-7) - ONE MUFU.EX2 and 7 FMA-pipe instructions, no reciprocal, no branches.  Beyond a = 5.6 the argument
// is clamped (h < 1.1e-8 there).  gelu(x) = max(x, 0) - |x| h(|x|);  gelu'(x) = Phi(x) + x phi(x) with phi through a second EX2.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float phi_tail(float a) {   // a >= 0
  a = fminf(a, 5.6f);
  float p = fmaf(a, 1.775593238e-05f, -6.477629528e-04f);
  p = fmaf(a, p, 7.724055995e-03f);
  p = fmaf(a, p, -5.292675105e-02f);
  p = fmaf(a, p, -4.590827311e-01f);
  p = fmaf(a, p, -1.151116857e+00f);
  p = fmaf(a, p, -1.0f);
  return ex2_ftz(p);
}
__device__ __forceinline__ float gelu_f(float x) {
  const float a = fminf(fabsf(x), 5.6f);   // clamped in the product too (|error| < 7e-8 beyond 5.6): bit-identical to gelu2 below
  return fmaf(-a, phi_tail(a), fmaxf(x, 0.f));
}
__device__ __forceinline__ float dgelu_f(float x) {
  const float h = phi_tail(fabsf(x));
  const float pdf = ex2_ftz(fmaf(x * x, -0.72134752044448170f, -1.32574806473616470f));  // exp(-x^2/2) / sqrt(2 pi)
  const float cdf = x >= 0.f ? 1.0f - h : h;
  return fmaf(x, pdf, cdf);
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: the persistent kernels of the step are launched with the programmatic-stream-serialization
// attribute, announce at their very start that their successor may be scheduled (its CTAs become resident as ours retire and
// run their prologue - barrier init, TMEM allocation, descriptor prefetch - under our tail), and wait for their predecessor's
// memory to be complete and visible before the first global access.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ------------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: two fp32 lanes per instruction, a 64-bit register pair per operand)
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// two bf16 packed in a 32-bit word -> packed fp32 pair (element 0 = low half)
__device__ __forceinline__ f32x2 bf2_to_f2(uint32_t w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t f2_to_bf2(f32x2 v) { float a, b; upk2(v, a, b); return pack_bf16x2(a, b); }
__device__ __forceinline__ float hsum2(f32x2 v) { float a, b; upk2(v, a, b); return a + b; }

// Packed-pair versions of gelu_f / dgelu_f above (same polynomial, evaluated in -a so the final product needs no negation):
// 6.5 / 10 instructions per element instead of 10 / 16 - the GELU epilogues of the K = 768 GEMMs are issue-bound.
__device__ __forceinline__ f32x2 phi_tail2(float x0, float x1, f32x2& an) {   // returns (Phi(-|x0|), Phi(-|x1|)); an = -min(|x|, 5.6)
  an = pk2(fmaxf(-fabsf(x0), -5.6f), fmaxf(-fabsf(x1), -5.6f));
  f32x2 p = fma2(an, pk2(1.775593238e-05f, 1.775593238e-05f), pk2(6.477629528e-04f, 6.477629528e-04f));
  p = fma2(an, p, pk2(7.724055995e-03f, 7.724055995e-03f));
  p = fma2(an, p, pk2(5.292675105e-02f, 5.292675105e-02f));
  p = fma2(an, p, pk2(-4.590827311e-01f, -4.590827311e-01f));
  p = fma2(an, p, pk2(1.151116857e+00f, 1.151116857e+00f));
  p = fma2(an, p, pk2(-1.0f, -1.0f));
  float p0, p1;
  upk2(p, p0, p1);
  return pk2(ex2_ftz(p0), ex2_ftz(p1));
}
__device__ __forceinline__ f32x2 gelu2(float x0, float x1) {
  f32x2 an;
  const f32x2 h = phi_tail2(x0, x1, an);
  return fma2(an, h, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));   // max(x, 0) - |x| Phi(-|x|)
}
__device__ __forceinline__ f32x2 dgelu2(float x0, float x1) {
  f32x2 an;
  const f32x2 h = phi_tail2(x0, x1, an);
  const f32x2 x = pk2(x0, x1);
  const f32x2 arg = fma2(mul2(x, x), pk2(-0.72134752044448170f, -0.72134752044448170f), pk2(-1.32574806473616470f, -1.32574806473616470f));
  float a0, a1;
  upk2(arg, a0, a1);
  const f32x2 pdf = pk2(ex2_ftz(a0), ex2_ftz(a1));
  // Phi(x) = 0.5 + copysign(0.5 - h, x)   (h = Phi(-|x|) <= 0.5)
  float t0, t1;
  upk2(fma2(h, pk2(-1.f, -1.f), pk2(0.5f, 0.5f)), t0, t1);
  t0 = __uint_as_float(__float_as_uint(t0) | (__float_as_uint(x0) & 0x80000000u));
  t1 = __uint_as_float(__float_as_uint(t1) | (__float_as_uint(x1) & 0x80000000u));
  const f32x2 cdf = add2(pk2(t0, t1), pk2(0.5f, 0.5f));
  return fma2(x, pdf, cdf);
}
// gelu(x) and gelu'(x) from ONE evaluation of the Phi tail: the forward epilogue that stores gelu'(u) for the backward
// (CLIPDLM_EPI_STORE_GELU_DERIV) pays the exp(-x^2/2) term on top of gelu2, the backward then only multiplies.
__device__ __forceinline__ void gelu_dgelu2(float x0, float x1, f32x2& g, f32x2& d) {
  f32x2 an;
  const f32x2 h = phi_tail2(x0, x1, an);
  g = fma2(an, h, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  const f32x2 x = pk2(x0, x1);
  const f32x2 arg = fma2(mul2(x, x), pk2(-0.72134752044448170f, -0.72134752044448170f), pk2(-1.32574806473616470f, -1.32574806473616470f));
  float a0, a1;
  upk2(arg, a0, a1);
  const f32x2 pdf = pk2(ex2_ftz(a0), ex2_ftz(a1));
  float t0, t1;
  upk2(fma2(h, pk2(-1.f, -1.f), pk2(0.5f, 0.5f)), t0, t1);
  t0 = __uint_as_float(__float_as_uint(t0) | (__float_as_uint(x0) & 0x80000000u));
  t1 = __uint_as_float(__float_as_uint(t1) | (__float_as_uint(x1) & 0x80000000u));
  d = fma2(x, pdf, add2(pk2(t0, t1), pk2(0.5f, 0.5f)));
}

// ------------------------------------------------------------------------------------------------
// mbarrier (shared::cta) with a bounded spin: a protocol bug traps instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz: protocol deadlock
      printf("clipdlm: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// thread-block cluster (CTA pair) helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// Arrive on an mbarrier that may live in the peer CTA. Default .release.cta semantics on purpose: a .cluster-scope release
// compiles to MEMBAR.ALL.GPU + CCTL per arrive (measured: 4x slower pipeline); the data these barriers order travels
// through the async proxy (TMA complete_tx) or TMEM (tcgen05.wait::ld + tcgen05 fences), not through generic memory.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, completing on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// L2 prefetch of a 2-D tile (no shared-memory destination, no barrier): issued one work item ahead, so that the real load hits L2.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1) : "memory");
}

// 1-D bulk copy global -> shared (contiguous bytes, multiple of 16, both addresses 16-byte aligned), completing on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 8 consecutive bf16 (pair) elements from shared memory as floats
__device__ __forceinline__ void lds8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float (&v)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi);
  float2 a = unpack_bf16x2(h.x), b = unpack_bf16x2(h.y), c = unpack_bf16x2(h.z), d = unpack_bf16x2(h.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  if (lo != nullptr) {
    const uint4 l = *reinterpret_cast<const uint4*>(lo);
    a = unpack_bf16x2(l.x); b = unpack_bf16x2(l.y); c = unpack_bf16x2(l.z); d = unpack_bf16x2(l.w);
    v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
  }
}
// TMA store of a shared-memory tile (written with generic-proxy stores + fence.proxy.async) into a 2-D tensor; bulk async-group
// bookkeeping: commit after the store(s), wait_group.read N before the source buffer is overwritten again.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// CTA-pair variants: the data lands in THIS CTA's shared memory, the transaction bytes are signalled on the mbarrier at
// shared::cluster address `bar_cluster` (the leader CTA's barrier, which the MMA-issuing thread waits on).
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_cg2(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// L2 cache-policy operands for TMA loads (.L2::cache_hint): the fixed encodings createpolicy.fractional.L2::evict_{first,last} (fraction 1.0)
// produces. evict_last keeps a small, heavily re-read operand (the 47 MB embedding table of the lm_head GEMMs) resident while gigabytes of
// logits stream through L2; evict_first marks an operand that is read exactly once.
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2_hint(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// CTA-pair (cta_group::2) variants: one warp of EACH CTA of the pair executes alloc / dealloc.
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (bf16 inputs, fp32 accumulate); one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// CTA-pair MMA: issued by one thread of the leader CTA; A (M = 256: 128 rows from each CTA) and B (N split in halves, one
// per CTA) are read from the same shared-memory offsets in both CTAs, each CTA's TMEM receives its 128 accumulator rows.
__device__ __forceinline__ void umma_bf16_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all prior MMAs of this thread completed) on the mbarrier at the same shared-memory offset in both CTAs.
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i gets lane base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 64-bit shared-memory matrix descriptor for tcgen05.mma, 128-byte swizzle (layout_type = 2, version = 1).
// Field layout follows the published UMMA descriptor (start>>4 @0, LBO>>4 @16, SBO>>4 @32, version @46,
// layout @61).  K-major : rows of 128 B, 8-row groups SBO = 1024 B apart (LBO ignored, set to 1).
//                MN-major: 64-element (128 B) MN runs, k-rows 128 B apart, 8-k groups SBO apart, MN groups LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// 32-bit instruction descriptor, kind::f16: D=f32, A=B=bf16, given majors (0 = K, 1 = MN), M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace clipdlm
