/* A non-Python host of libclipdlm.so: one train step and a 3-step denoise loop of a small CLIP-Diffusion-LM model driven from plain
 * C through include/clipdlm.h — no torch, no Python in the process. This is the binding a C / C++ / JNI / cgo host would write
 * (INTEGRATION.md section 2); the Python package does exactly these calls through ctypes.
 *
 *   gcc -O2 -std=c99 examples/c_host.c -Iinclude -I/usr/local/cuda/include -Ldiffusion-image-captioning_b200 -lclipdlm \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/diffusion-image-captioning_b200 -o examples/c_host
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "clipdlm.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
#define LK(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "clipdlm error %d at %s:%d: %s\n", r_, __FILE__, __LINE__, clipdlm_last_error()); return 1; } } while (0)

static uint32_t rng_state = 12345u;
static float urand(void) { rng_state = rng_state * 1664525u + 1013904223u; return (float)(rng_state >> 8) / 16777216.0f - 0.5f; } /* U(-0.5, 0.5) */

static float* dev_f32(const float* host, size_t n) {
  float* d = NULL;
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) return NULL;
  if (host) cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice); else cudaMemset(d, 0, n * sizeof(float));
  return d;
}

int main(void) {
  if (clipdlm_device_ok() != 1) { fprintf(stderr, "needs a compute-capability 10.x GPU (B200)\n"); return 2; }
  enum { LAYERS = 2, D = 768, HEADS = 12, FFN = 3072, VOCAB = 1000, ML = 16, CLIP = 512, MAXPOS = 64, B = 4, S = 3, L = ML + 2 };
  const char* env_dbg = getenv("C_HOST_GEMM_DEBUG_FLAGS"); /* e.g. 1024: row-wise global stores instead of TMA stores (see clipdlm_gemm_debug_flags) */
  if (env_dbg) clipdlm_gemm_debug_flags((uint32_t)strtoul(env_dbg, NULL, 0));
  const char* env_drop = getenv("C_HOST_DROPOUT"); /* e.g. 0.1: train with dropout (tools/sanitize_c_host.sh runs both settings under compute-sanitizer) */
  const float pdrop = env_drop ? (float)atof(env_drop) : 0.0f;
  clipdlm_config_t cfg = {LAYERS, D, HEADS, FFN, VOCAB, ML, CLIP, MAXPOS, /*fusion concat*/ 0, /*precision bf16*/ 0, 1e-12f, pdrop, pdrop};
  const int64_t n = clipdlm_param_count(&cfg);
  if (n <= 0) { fprintf(stderr, "bad config: %s\n", clipdlm_last_error()); return 1; }

  /* ---- parameters: N(0, 0.02)-like weights, LayerNorm weights = 1 (HF init), frozen embedding table ---- */
  float* hp = (float*)malloc((size_t)n * sizeof(float));
  for (int64_t i = 0; i < n; ++i) hp[i] = 0.07f * urand();
  int ln_slots[2 + 2 * LAYERS], ns = 0;
  ln_slots[ns++] = CLIPDLM_P_EMB_LN_W; ln_slots[ns++] = CLIPDLM_P_VLN_W;
  for (int l = 0; l < LAYERS; ++l) {
    ln_slots[ns++] = CLIPDLM_P_LAYER0 + l * CLIPDLM_P_PER_LAYER + CLIPDLM_PL_LN1_W;
    ln_slots[ns++] = CLIPDLM_P_LAYER0 + l * CLIPDLM_P_PER_LAYER + CLIPDLM_PL_LN2_W;
  }
  for (int k = 0; k < ns; ++k) {
    const int64_t off = clipdlm_param_offset(&cfg, ln_slots[k]), cnt = clipdlm_param_size(&cfg, ln_slots[k]);
    for (int64_t i = 0; i < cnt; ++i) hp[off + i] = 1.0f;
  }
  const int vpad = (VOCAB + 255) / 256 * 256;
  float* hemb = (float*)calloc((size_t)vpad * D, sizeof(float));
  for (int64_t i = 0; i < (int64_t)VOCAB * D; ++i) hemb[i] = 0.07f * urand();

  float *params = dev_f32(hp, n), *grads = dev_f32(NULL, n), *adam_m = dev_f32(NULL, n), *adam_v = dev_f32(NULL, n), *emb = dev_f32(hemb, (size_t)vpad * D);
  void *shadow = NULL, *emb_bf = NULL;
  CK(cudaMalloc(&shadow, (size_t)n * 2));
  CK(cudaMalloc(&emb_bf, (size_t)vpad * D * 2));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  LK(clipdlm_to_bf16(params, shadow, NULL, n, st));
  LK(clipdlm_to_bf16(emb, emb_bf, NULL, (int64_t)vpad * D, st));

  /* ---- batch: token ids, CLIP features, noise, diffusion coefficients ---- */
  int32_t hids[B * ML];
  for (int i = 0; i < B * ML; ++i) { rng_state = rng_state * 1664525u + 1013904223u; hids[i] = (int32_t)((rng_state >> 10) % VOCAB); }
  float himg[B * CLIP], hnoise[B * ML * D];
  for (int i = 0; i < B * CLIP; ++i) himg[i] = 0.1f * urand();
  for (int i = 0; i < B * ML * D; ++i) hnoise[i] = 2.0f * urand();
  float hca[S] = {0.99f, 0.7f, 0.1f}, hcb[S];
  for (int s = 0; s < S; ++s) hcb[s] = sqrtf(1.0f - hca[s] * hca[s]);
  int32_t* ids = NULL;
  CK(cudaMalloc((void**)&ids, sizeof(hids)));
  CK(cudaMemcpy(ids, hids, sizeof(hids), cudaMemcpyHostToDevice));
  float *img = dev_f32(himg, B * CLIP), *txt = dev_f32(NULL, B * CLIP), *noise = dev_f32(hnoise, B * ML * D), *ca = dev_f32(hca, S), *cb = dev_f32(hcb, S);
  double* losses = NULL;
  CK(cudaMalloc((void**)&losses, 2 * sizeof(double)));
  CK(cudaMemset(losses, 0, 2 * sizeof(double)));

  /* ---- one train step: forward (embed + q_sample + fusion + encoder), loss + full backward, AdamW ---- */
  const int R = S * B;
  size_t ws_bytes = clipdlm_workspace_bytes(&cfg, R, B, 1);
  void* ws = NULL;
  CK(cudaMalloc(&ws, ws_bytes));
  clipdlm_buffers_t bufs = {params, grads, shadow, NULL, emb, emb_bf, NULL, ws, ws_bytes};
  clipdlm_engine_t* eng = clipdlm_engine_create(&cfg, &bufs, R, B, 1);
  if (!eng) { fprintf(stderr, "engine_create: %s\n", clipdlm_last_error()); return 1; }
  if (getenv("C_HOST_FUSED") && atoi(getenv("C_HOST_FUSED"))) {
    /* the options the Python host switches on by default in plain bf16: factored softmax-CE gradient of the lm_head (exponent shift 0 without
       CLIPDLM_OPT_EXP_SHIFT_PTR) and gelu'(u) stored by lin1's epilogue + lin1's bias gradient summed in the lin2 gradient GEMM */
    LK(clipdlm_engine_set_option(eng, CLIPDLM_OPT_FUSED_SOFTMAX_GRAD, 1));
    LK(clipdlm_engine_set_option(eng, CLIPDLM_OPT_GELU_DERIV_STORE, 2));
  }
  clipdlm_pass_t pass;
  memset(&pass, 0, sizeof(pass));
  pass.R = R; pass.B = B; pass.mode = 1; pass.train = 1;
  pass.ids = ids; pass.noise = noise; pass.coef_a = ca; pass.coef_b = cb; pass.image_clip = img; pass.text_clip = txt; pass.drop_seed = 1;
  LK(clipdlm_engine_forward(eng, &pass, st));
  clipdlm_loss_cfg_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.loss_kind = 0; lc.use_embed_loss = 1; lc.use_prob_loss = 1; lc.batch_size = B; lc.R_total = R; lc.rounding_weight = 0.5f; lc.backward = 1;
  LK(clipdlm_engine_loss_backward(eng, &lc, losses, st));
  LK(clipdlm_adamw(params, grads, adam_m, adam_v, shadow, NULL, n, 1e-4f, 0.9f, 0.999f, 1e-8f, 0.01f, 1, 1.0f, 1, st));
  double hl[2];
  CK(cudaMemcpyAsync(hl, losses, sizeof(hl), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  printf("train step: x_t_loss %.5f  rounding CE %.5f  (%lld kernel launches)\n", hl[0], hl[1], (long long)clipdlm_engine_launch_count(eng));
  if (!(hl[0] > 0.0 && hl[0] < 1e4 && hl[1] > 0.0 && hl[1] < 1e4)) { fprintf(stderr, "implausible losses\n"); return 1; }
  clipdlm_engine_destroy(eng);
  CK(cudaFree(ws));

  /* ---- denoise loop (CLIP-DDPM.py:613-621): x <- model(x[:, :16]) three times, then argmax of the rounding projection ---- */
  ws_bytes = clipdlm_workspace_bytes(&cfg, B, B, 0);
  CK(cudaMalloc(&ws, ws_bytes));
  bufs.workspace = ws; bufs.workspace_bytes = ws_bytes;
  eng = clipdlm_engine_create(&cfg, &bufs, B, B, 0);
  if (!eng) { fprintf(stderr, "engine_create: %s\n", clipdlm_last_error()); return 1; }
  float* hx = (float*)malloc((size_t)B * L * D * sizeof(float));
  for (int i = 0; i < B * L * D; ++i) hx[i] = 2.0f * urand();
  float *cur = dev_f32(hx, (size_t)B * L * D), *nxt = dev_f32(NULL, (size_t)B * L * D);
  for (int step = 0; step < 3; ++step) {
    memset(&pass, 0, sizeof(pass));
    pass.R = B; pass.B = B; pass.mode = 0; pass.reuse_proj = step > 0;
    pass.x_in = cur; pass.x_in_stride = (int64_t)L * D; pass.image_clip = img; pass.text_clip = txt; pass.x_out = nxt;
    LK(clipdlm_engine_forward(eng, &pass, st));
    float* t = cur; cur = nxt; nxt = t;
  }
  int32_t* argmax = NULL;
  CK(cudaMalloc((void**)&argmax, B * ML * sizeof(int32_t)));
  LK(clipdlm_engine_lm_head(eng, NULL, 0, argmax, st));
  int32_t hout[B * ML];
  CK(cudaMemcpyAsync(hout, argmax, sizeof(hout), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  printf("denoise ids[0]:");
  for (int j = 0; j < ML; ++j) { printf(" %d", hout[j]); if (hout[j] < 0 || hout[j] >= VOCAB) { fprintf(stderr, "\nid out of range\n"); return 1; } }
  printf("\nc_host ok\n");
  clipdlm_engine_destroy(eng);
  return 0;
}
