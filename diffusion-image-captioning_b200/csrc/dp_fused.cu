// Data-parallel gradient exchange fused with the optimizer: ONE kernel per step does
//     reduce-scatter(gradients)  ->  AdamW on the rank's slice  ->  all-gather(updated fp32 weights + bf16 shadow)
// over NVLink / NVSwitch peer memory, instead of  ncclAllReduce(177 MB) + a full-size AdamW pass on every rank.
//
// The flat gradient / parameter / shadow buffers of all ranks live in symmetric memory (same size, same layout on every GPU, mapped
// into every process). Rank r owns the element slice [r * slice, (r + 1) * slice):
//   * multicast path (NVSwitch multimem, "NVLS"): the gradient slice is read with  multimem.ld_reduce.add.v4.f32  on the multicast
//     address - the switch returns the sum over all GPUs in one load - and the updated weights / shadow are written once with
//     multimem.st, which the switch replicates to every GPU;
//   * peer path (no multicast object): the slice is summed with plain loads from every peer's buffer and written with plain
//     stores to every peer's buffer.
// Adam moments exist only for the owned slice's elements (ZeRO-1 style: optimizer HBM traffic and the moment update are 1 / world per
// rank). The caller brackets the kernel with device-side barriers over the same symmetric-memory group: all backward passes
// complete before, all remote writes visible after (parallel.py).
//
// Replaces (absent in the single-GPU reference): ncclAllReduce on the flat gradient + clipdlm_adamw.
#include "common.cuh"
#include "../../include/clipdlm.h"

namespace clipdlm {

int num_sms();

struct DpPeers {
  float* p[CLIPDLM_MAX_PEERS];
  const float* g[CLIPDLM_MAX_PEERS];
  __nv_bfloat16* sh_hi[CLIPDLM_MAX_PEERS];
  __nv_bfloat16* sh_lo[CLIPDLM_MAX_PEERS];
};

__device__ __forceinline__ float4 multimem_ld_reduce_add_f32x4(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_f32x4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_b32x2(void* mc, const uint2& v) {   // 8 bytes of packed bf16, type-agnostic store
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}

template <bool MC>
__global__ void __launch_bounds__(256) adamw_dp_kernel(DpPeers peers, float* p_mc, const float* g_mc, __nv_bfloat16* sh_hi_mc,
                                                       __nv_bfloat16* sh_lo_mc, float* __restrict__ m, float* __restrict__ v,
                                                       long long i4_begin, long long i4_end, int world, int have_lo, float decay, float beta1,
                                                       float beta2, float eps, float step_size, float inv_sqrt_bc2, float grad_scale) {
  for (long long i = i4_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i4_end; i += (long long)gridDim.x * blockDim.x) {
    float4 gg;
    if (MC) {
      gg = multimem_ld_reduce_add_f32x4(g_mc + 4 * i);
    } else {
      gg = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(peers.g[r]) + i);
        gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
      }
    }
    float4 pp = reinterpret_cast<const float4*>(peers.p[0])[i];   // peers.p[0] is this rank's own copy (every copy is identical)
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = reinterpret_cast<float*>(&pp); float* G = reinterpret_cast<float*>(&gg);
    float* M = reinterpret_cast<float*>(&mm); float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // identical arithmetic to adamw_kernel (hbm_kernels.cu)
      const float gr = G[j] * grad_scale;
      P[j] *= decay;
      M[j] = beta1 * M[j] + (1.f - beta1) * gr;
      V[j] = beta2 * V[j] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(V[j]) * inv_sqrt_bc2 + eps;
      P[j] -= step_size * (M[j] / denom);
    }
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    uint2 h, l;
    h.x = pack_bf16x2(P[0], P[1]); h.y = pack_bf16x2(P[2], P[3]);
    l.x = pack_bf16x2(P[0] - bf16_round(P[0]), P[1] - bf16_round(P[1]));
    l.y = pack_bf16x2(P[2] - bf16_round(P[2]), P[3] - bf16_round(P[3]));
    if (MC) {
      multimem_st_f32x4(p_mc + 4 * i, pp);
      multimem_st_b32x2(sh_hi_mc + 4 * i, h);
      if (have_lo) multimem_st_b32x2(sh_lo_mc + 4 * i, l);
    } else {
      for (int r = 0; r < world; ++r) {
        reinterpret_cast<float4*>(peers.p[r])[i] = pp;
        reinterpret_cast<uint2*>(peers.sh_hi[r])[i] = h;
        if (have_lo) reinterpret_cast<uint2*>(peers.sh_lo[r])[i] = l;
      }
    }
  }
}

}  // namespace clipdlm

using namespace clipdlm;

extern "C" {

int clipdlm_dp_slice(int64_t n, int32_t rank, int32_t world, int64_t* begin, int64_t* end) {
  CLIPDLM_CHECK(n > 0 && n % 4 == 0 && world >= 1 && rank >= 0 && rank < world && begin && end, "dp_slice: bad arguments");
  const long long per = ((n / 4 + world - 1) / world + 1) / 2 * 2 * 4;   // elements per rank, a multiple of 8 (16-byte bf16 stores stay aligned)
  long long b = (long long)rank * per, e = b + per;
  if (b > n) b = n;
  if (e > n) e = n;
  *begin = b; *end = e;
  return 0;
}

int clipdlm_adamw_dp(const clipdlm_dp_buffers_t* d, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     float weight_decay, int32_t step, float grad_scale, clipdlm_stream stream) {
  CLIPDLM_CHECK(d != nullptr && m && v && n > 0 && n % 4 == 0 && step >= 1, "adamw_dp: bad arguments");
  CLIPDLM_CHECK(d->world >= 1 && d->world <= CLIPDLM_MAX_PEERS && d->rank >= 0 && d->rank < d->world, "adamw_dp: bad rank %d / world %d", d->rank, d->world);
  const bool mc = d->p_mc != nullptr && d->g_mc != nullptr && d->shadow_hi_mc != nullptr;
  bool have_lo = d->shadow_lo[0] != nullptr;
  CLIPDLM_CHECK(!mc || !have_lo || d->shadow_lo_mc != nullptr, "adamw_dp: split-precision shadow without a multicast mapping");
  DpPeers peers;
  memset(&peers, 0, sizeof(peers));
  for (int r = 0; r < d->world; ++r) {
    // slot 0 = this rank's own buffers, then the others in rank order
    const int src = r == 0 ? d->rank : (r <= d->rank ? r - 1 : r);
    CLIPDLM_CHECK(d->p[src] && d->g[src] && d->shadow_hi[src], "adamw_dp: null peer pointer for rank %d", src);
    peers.p[r] = d->p[src]; peers.g[r] = d->g[src];
    peers.sh_hi[r] = (__nv_bfloat16*)d->shadow_hi[src]; peers.sh_lo[r] = (__nv_bfloat16*)d->shadow_lo[src];
  }
  int64_t b = 0, e = 0;
  int rc = clipdlm_dp_slice(n, d->rank, d->world, &b, &e);
  if (rc) return rc;
  if (e <= b) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float decay = (float)(1.0 - (double)lr * (double)weight_decay);
  const long long n4 = (e - b) / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  if (mc)
    adamw_dp_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(peers, d->p_mc, d->g_mc, (__nv_bfloat16*)d->shadow_hi_mc, (__nv_bfloat16*)d->shadow_lo_mc,
                                                            m, v, b / 4, e / 4, d->world, have_lo ? 1 : 0, decay, beta1, beta2, eps, step_size,
                                                            inv_sqrt_bc2, grad_scale);
  else
    adamw_dp_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(peers, nullptr, nullptr, nullptr, nullptr, m, v, b / 4, e / 4, d->world,
                                                             have_lo ? 1 : 0, decay, beta1, beta2, eps, step_size, inv_sqrt_bc2, grad_scale);
  CLIPDLM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
