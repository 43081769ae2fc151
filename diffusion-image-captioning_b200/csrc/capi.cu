// extern "C" surface of libclipdlm.so (declared in include/clipdlm.h): thin, exception-free forwarding
// to the kernel dispatchers. No torch types cross this boundary.
#include "common.cuh"
#include "../../include/clipdlm.h"
#include <stdarg.h>
#include <string.h>

namespace clipdlm {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int gemm_dispatch(const clipdlm_gemm_t* g, cudaStream_t st);
int lse_combine_dispatch(const float* pmax, const float* psum, const int* parg, int n_tiles, int M, const float* tgt_logit, float* lse,
                         int* argmax, double* loss_acc, double scale, cudaStream_t st);
void gemm_debug_mn_desc(uint32_t lbo, uint32_t sbo);

}  // namespace clipdlm

using namespace clipdlm;

extern "C" {

const char* clipdlm_last_error(void) { return g_last_error; }
int clipdlm_version(void) { return 100; }

int clipdlm_device_ok(void) {
  int dev = 0, major = 0;
  CLIPDLM_CUDA_OK(cudaGetDevice(&dev));
  CLIPDLM_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}

int clipdlm_gemm(const clipdlm_gemm_t* g, clipdlm_stream stream) { return gemm_dispatch(g, (cudaStream_t)stream); }

void clipdlm_gemm_debug_mn_desc(uint32_t lbo_bytes, uint32_t sbo_bytes) { gemm_debug_mn_desc(lbo_bytes, sbo_bytes); }

int clipdlm_lse_combine(const float* part_max, const float* part_sum, const int32_t* part_arg, int32_t n_tiles, int32_t M,
                        const float* tgt_logit, float* lse, int32_t* argmax, double* loss_acc, double scale, clipdlm_stream stream) {
  return lse_combine_dispatch(part_max, part_sum, part_arg, n_tiles, M, tgt_logit, lse, argmax, loss_acc, scale, (cudaStream_t)stream);
}

}  // extern "C"
