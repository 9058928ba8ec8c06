/* Minimal C host of libiodine_b200.so: plan for the CLEVR6 architecture, caller-owned workspace, one reconstruct()
 * with host buffers.  Weights are left at zero here -- a real host fills IodineWeights with device pointers to the
 * reference's state_dict tensors.
 *
 *   gcc examples/host_c.c -Iinclude -Liodine_b200/lib -liodine_b200 -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/iodine_b200/lib -o host_c
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "iodine_b200.h"

/* the two CUDA runtime calls a C host needs, declared here to keep the example free of cuda_runtime.h */
extern int cudaMalloc(void** p, size_t n);
extern int cudaMemset(void* p, int v, size_t n);

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      fprintf(stderr, "%s failed: %s\n", #call, iodine_last_error());      \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void) {
  IodineShape s;
  memset(&s, 0, sizeof s);
  s.B = 2; s.K = 7; s.L = 64; s.H = 128; s.W = 128; s.T = 5; s.img_c = 3;
  s.dec_layers = 4; s.dec_chan = 64; s.dec_k = 3;
  s.ref_layers = 4; s.ref_chan = 64; s.ref_k = 3; s.ref_stride = 2;
  s.mlp_units = 256; s.layernorm = 1; s.sigma = 0.10f; s.precision = IODINE_FP16;

  IodinePlan* plan = NULL;
  CHECK(iodine_plan_create(&s, &plan));
  size_t ws_bytes = 0;
  CHECK(iodine_plan_workspace_bytes(plan, &ws_bytes));
  void* ws = NULL;
  if (cudaMalloc(&ws, ws_bytes) != 0) { fprintf(stderr, "cudaMalloc(%zu) failed\n", ws_bytes); return 1; }
  CHECK(iodine_plan_set_workspace(plan, ws, ws_bytes));

  /* weights: one device pointer per state_dict key; zeros stand in for a checkpoint here */
  size_t wbytes = 8u << 20;
  void* zeros = NULL;
  if (cudaMalloc(&zeros, wbytes) != 0 || cudaMemset(zeros, 0, wbytes) != 0) return 1;
  IodineWeights w;
  memset(&w, 0, sizeof w);
  for (int i = 0; i < s.dec_layers; ++i) { w.dec_w[i] = zeros; w.dec_b[i] = zeros; }
  for (int i = 0; i < s.ref_layers; ++i) { w.ref_w[i] = zeros; w.ref_b[i] = zeros; }
  w.dec_out_w = w.dec_out_b = w.mlp_w = w.mlp_b = zeros;
  w.lstm_w_ih = w.lstm_w_hh = w.lstm_b_ih = w.lstm_b_hh = zeros;
  w.mean_w = w.mean_b = w.logvar_w = w.logvar_b = w.init_mean = w.init_logvar = zeros;
  CHECK(iodine_plan_set_weights(plan, &w, NULL));

  size_t HW = (size_t)s.H * s.W, BK = (size_t)s.B * s.K;
  float* x = calloc((size_t)s.B * 3 * HW, sizeof(float));
  float* eps = calloc((size_t)(s.T + 1) * BK * s.L, sizeof(float));
  float* pred = malloc((size_t)s.B * 3 * HW * sizeof(float));
  float* mask = malloc(BK * HW * sizeof(float));
  float terms[2 * 16];
  CHECK(iodine_reconstruct_host(plan, x, eps, pred, mask, NULL, NULL, terms, NULL));
  printf("mask[0] = %g (1/K = %g), sum_b log-lik of the last step = %g\n", mask[0], 1.0 / s.K, terms[2 * (s.T - 1)]);
  iodine_plan_destroy(plan);
  return 0;
}
