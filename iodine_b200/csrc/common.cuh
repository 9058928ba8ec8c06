// common.cuh -- shared declarations for the iodine_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/iodine_b200.h"

namespace iod {

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
#define IOD_CHECK_CUDA(expr)                                                      \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      iod::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                     cudaGetErrorString(_e));                                     \
      return 1;                                                                   \
    }                                                                             \
  } while (0)
#define IOD_REQUIRE(cond, ...)                                                    \
  do {                                                                            \
    if (!(cond)) {                                                                \
      iod::set_error(__VA_ARGS__);                                                \
      return 1;                                                                   \
    }                                                                             \
  } while (0)
#define IOD_LAUNCH_CHECK(plan)                                                    \
  do {                                                                            \
    (plan)->launches++;                                                           \
    IOD_CHECK_CUDA(cudaGetLastError());                                           \
  } while (0)

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float elu_f(float v) { return v > 0.f ? v : expm1f(v); }
// ELU'(pre) recovered from the post-activation a = ELU(pre): a>0 -> 1, else exp(pre)=a+1
__device__ __forceinline__ float elu_grad_from_act(float a) { return a > 0.f ? 1.f : a + 1.f; }
__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// border class of coordinate y for a k-tap zero-padded conv (p = k/2): 0..p-1 = distance
// from the low border, p = interior, p+1..2p = high border (2p = last row/col).
__host__ __device__ __forceinline__ int border_class(int y, int n, int p) {
  if (y < p) return y;
  if (y >= n - p) return 2 * p - (n - 1 - y);
  return p;
}

// ------------------------------------------------------------------ the plan
struct ConvPack {       // fp32 direct-conv weights repacked [chunk][tap][ci_in_chunk][co]
  float* w = nullptr;   // forward
  float* wt = nullptr;  // data-gradient (transposed + flipped)
  float* b = nullptr;   // bias [co]
};

struct Plan {
  IodineShape s;
  int device = 0;
  int num_sms = 148;
  int BK = 0, HW = 0, M = 0, C = 0, Cr = 0;
  // K-split (IodineShape.slot_ranks > 1): s.K is rewritten to the LOCAL slot count at create, so that every
  // per-slot loop and buffer of the plan covers this rank's slots only; the pixel mixture and the recombination
  // are the two places that see all K_total slots of an image (through out4_all)
  int K_total = 0;            // ARCH.SLOTS
  int ks_ranks = 1, ks_rank = 0;
  int n_class = 0;            // dec_k * dec_k border classes of decoder layer 1
  int ref_h[IODINE_MAX_LAYERS + 1];
  int ref_w[IODINE_MAX_LAYERS + 1];
  uint64_t launches = 0;
  bool weights_set = false;
  void* comm = nullptr;       // ncclComm_t installed by iodine_plan_set_comm (nullptr: single rank)
  int comm_rank = 0, comm_nranks = 1;
  int profiling = 0;          // iodine_plan_profile: 0 off, 1 decoder C->C convolutions, 2 pixel-mixture kernel
  std::vector<cudaEvent_t> prof_events;   // pairs (start, stop)
  size_t prof_used = 0;

  // ---- CUDA-graph replay of encode() (plan.cu): the ~125 launches of one call are captured once per
  //      (x, eps, outputs) pointer tuple and replayed on a plan-owned stream
  struct EncodeGraph {
    cudaGraphExec_t exec = nullptr;
    const void* key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool warm = false;                                                      // key was seen by one eager call
    uint64_t used = 0;                                                      // graph_clock of the last use (LRU)
    uint64_t kernels = 0;                                                   // kernel launches inside the graph
  };
  static constexpr int N_GRAPHS = 4;
  EncodeGraph graph[N_GRAPHS];
  uint64_t graph_clock = 0;
  cudaStream_t gstream = nullptr;
  cudaEvent_t gev_in = nullptr, gev_out = nullptr;
  bool graphs = true;

  // ---- weights (plan-owned device memory, allocated at create)
  float* wsum = nullptr;      // [n_class][C][L]   layer-1 tap sums per border class
  float* wsumT = nullptr;     // [n_class][L][C]   the same, transposed (coalesced over C for the forward GEMV)
  float* ptab = nullptr;      // [H][W][C]         layer-1 coord-conv + bias table
  ConvPack dec[IODINE_MAX_LAYERS];   // layers 1..n-1 (index = layer), C->C
  float* out_w = nullptr;     // [tap][ci][4]      decoder.conv forward
  float* out_wt = nullptr;    // [tap][4][co]      decoder.conv data-gradient (flipped)
  float* out_b = nullptr;     // [4]
  float* ref_w0 = nullptr;    // [tap][17][Cr]     refine layer 0
  float* ref_wp[IODINE_MAX_LAYERS];  // [tap][ci][Cr] refine layers >= 1
  float* ref_b[IODINE_MAX_LAYERS];
  float *mlp_w = nullptr, *mlp_b = nullptr;
  float *w_ih = nullptr, *w_hh = nullptr, *b_ih = nullptr, *b_hh = nullptr;
  float *head_w = nullptr, *head_b = nullptr;   // [2L, M] (mean rows then logvar rows), [2L]
  float *init_mean = nullptr, *init_logvar = nullptr;
  void* train = nullptr;      // training state (train.cu: TrainState)
  float* tape_u = nullptr;    // training tape hooks of launch_head: the MLP's inner ELU [BK,M] ...
  float* tape_gates = nullptr;  // ... and the summed LSTM gate pre-activations [BK,4M]
  void* tc = nullptr;         // tensor-core path state (conv_tc.cu: TcState), IODINE_BF16 / IODINE_FP16
  void* rtc = nullptr;        // refinement-encoder tensor-core state (refine_tc.cu: RtcState)

  // ---- workspace carve-up (caller memory)
  void* ws = nullptr;
  size_t ws_bytes = 0, ws_need = 0;
  void* act[IODINE_MAX_LAYERS];   // [BK,H,W,C] fp32 (or bf16 in IODINE_BF16) post-activations
  void* gbuf[2];                  // dgrad ping-pong, same type as act
  float* out4 = nullptr;          // [BK,H,W,4] this rank's slots (K-split: a window of out4_all)
  float* out4_all = nullptr;      // [ks_ranks][B,K,H,W,4]: every rank's out4 after the all-gather (== out4 without K-split)
  float* seed4 = nullptr;         // [BK,H,W,4] fp32; IODINE_BF16: the same bytes hold [BK,H,W,8] bf16
                                  // (4 gradient channels + 4 zeros = one 8-channel plane)
  float* auxs = nullptr;          // [BK,H,W,12] per-slot raw aux channels
  float* lik = nullptr;           // [B,H,W] raw pixel likelihood
  float* enc20 = nullptr;         // [BK,H,W,20] normalised refinement input (17 + 3 pad)
  float* rbuf[2];                 // refine conv ping-pong (fp32 path)
  void* enc16 = nullptr;          // [BK,2,H,W,8] 16-bit chunk-planar refinement input, 15 data channels (16-bit modes)
  void* r16[2];                   // refine conv ping-pong, 16-bit chunk-planar (16-bit modes)
  float* z = nullptr;             // [BK,L]
  float* u = nullptr;             // [BK,n_class,C]
  float* eu = nullptr;            // [BK,n_class,C] exp(u) (tensor-core modes: tc_layer1_kernel multiplies instead of evaluating)
  int* ubig = nullptr;            // [BK] 1 where a slot's |u| exceeds the range in which exp(u) * exp(p) is safe
  float* G = nullptr;             // [BK,n_class,C] class-wise pixel sums of dJ/d(pre-act 1)
  float* dz = nullptr;            // [BK,L]
  double* stats = nullptr;        // [BK,4,2] (sum, sumsq) for grad_means/grad_mask/lik/loo
  float* lnp = nullptr;           // [BK,8] their layer-norm parameters: 4 means, 4 x 1/(std + 1e-5) (post_grads_kernel)
  double* accum = nullptr;        // [2] (sum ll, sum kl) of the step
  float* pool = nullptr;          // [BK,Cr]
  float* xin = nullptr;           // [BK, M+4L]
  float* gates = nullptr;         // [BK, 4M]
  float* st_mean = nullptr;       // [BK,L] scratch state for encode()/reconstruct()
  float* st_logvar = nullptr;
  float* st_h = nullptr;          // [BK,M]
  float* st_c = nullptr;
  float* st_z = nullptr;          // [BK,L]
  float* st_terms = nullptr;      // [T,2]
  // staging for *_host entry points
  float* hx = nullptr;            // [B,3,H,W]
  float* heps = nullptr;          // [T+1,B,K,L]
  float* hpred = nullptr;         // [B,3,H,W]
  float* hmask = nullptr;         // [B,K,1,H,W]
  float* hmean = nullptr;         // [B,K,3,H,W]
  uint8_t* hamax = nullptr;       // [B,H,W] argmax over the K masks (iodine_evaluate_host)
  // logger side channel: image 0 of the last elbo() evaluation (iodine.py:225-239)
  float* log_pred = nullptr;      // [3,H,W]
  float* log_mask = nullptr;      // [K,H,W]
  float* log_mean = nullptr;      // [K,3,H,W]
};

// collapsed first decoder layer, tensor-core modes: |u|, |p| up to which exp(u + p) is formed as exp(u) * exp(p)
#define IOD_L1_EXP_RANGE 40.f

// split-K factor of the LSTM gate GEMM (head.cu); the gates buffer holds that many partial sums
constexpr int LSTM_KSPLIT = 8;

// tensor-core modes keep the decoder activations chunk-planar: 16 bytes per pixel and plane = 8 channels of 16-bit
// values (bf16 / fp16) or 4 channels of fp32 values rounded to tf32 (IODINE_TF32)
inline bool tc_mode(const Plan* p) {
  return p->s.precision == IODINE_BF16 || p->s.precision == IODINE_FP16 || p->s.precision == IODINE_TF32;
}
inline bool tf_mode(const Plan* p) { return p->s.precision == IODINE_TF32; }
inline size_t act_elem_bytes(const Plan* p) { return !tc_mode(p) ? 4 : tf_mode(p) ? 4 : 2; }
// operand format of the 16-bit tensor-core kernels of a mode: IEEE half for IODINE_FP16 and for the refinement
// encoder of IODINE_TF32 (same 10-bit mantissa; its inputs are layer-normalised / bounded), bfloat16 for IODINE_BF16
inline int half_is_f16(const Plan* p) { return p->s.precision != IODINE_BF16; }

// 16-bit pair packing selected at run time (helper kernels) -- f16 != 0: IEEE half, else bfloat16
__device__ __forceinline__ uint32_t pack_h2(float a, float b, int f16) {
  if (f16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u, int f16) {
  if (f16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

// ------------------------------------------------------------------ launchers (conv_f32.cu)
// stride-1 C->C conv, NHWC fp32.  mode 0: out = ELU(conv + bias); mode 1: out = conv * ELU'(act)
// mode 2: like 1 but instead of storing, accumulate class-wise pixel sums into G[n][class][co]
int launch_conv_cc(Plan* p, const float* in, const float* wpack, const float* bias,
                   const float* act_prev, float* out, float* G, int mode, cudaStream_t st);
int launch_conv_out4(Plan* p, const float* in, float* out4, cudaStream_t st);
int launch_dgrad_in4(Plan* p, const float* seed4, const float* act_prev, float* gout,
                     cudaStream_t st, bool store = false);
int launch_refine_convs(Plan* p, const float* x, cudaStream_t st, float* const* outs = nullptr);

// ------------------------------------------------------------------ launchers (mixture.cu)
// fused_aux: write the refinement input in the form refine_l0f_kernel consumes (refine_tc.cu) instead of raw fp32 auxs
int launch_mixture(Plan* p, const float* x, bool want_grads, cudaStream_t st, bool fused_aux = false);
int launch_recombine(Plan* p, float* pred, float* mask, float* mean, int n_images, cudaStream_t st,
                     uint8_t* amax = nullptr);   // amax[B,H,W]: argmax over the K masks (evaluator tail)
// slot (b, k) of an image, k in [0, K_total), inside out4_all [ks_ranks][B][Kl][HW] (float4 units): rank-major, as
// ncclAllGather leaves it; without K-split (Kl == K_total) this is the plain (b * K + k) * HW
__host__ __device__ __forceinline__ size_t out4_slot(int b, int k, int B, int Kl, int HW) {
  const int r = k / Kl;
  return ((size_t)(r * B + b) * Kl + (k - r * Kl)) * HW;
}
int launch_export_aux(Plan* p, const float* x, float* aux_out, cudaStream_t st);
int launch_assemble16(Plan* p, const float* x, cudaStream_t st);   // 16-bit modes: writes p->enc16

// ------------------------------------------------------------------ launchers (head.cu)
int launch_sample_l1(Plan* p, const float* mu, const float* logvar, const float* eps,
                     const float* z_in, float* act0_f32, cudaStream_t st);
int launch_post_grads(Plan* p, const float* mu, const float* logvar, const float* eps,
                      float* latent_out, cudaStream_t st);
int launch_head(Plan* p, float* mu, float* logvar, float* h, float* c, cudaStream_t st);
int launch_setup_weights(Plan* p, const IodineWeights* w, cudaStream_t st);

// ------------------------------------------------------------------ plan.cu pieces the training step reuses
int plan_check_ready(Plan* p);
void plan_prof_mark(Plan* p, cudaStream_t st, int cls);   // iodine_plan_profile bracket (cls 2: the pixel-mixture kernel)
int plan_decoder_forward(Plan* p, const float* mu, const float* lv, const float* eps, const float* z_in, cudaStream_t st);
int plan_allreduce_sum(Plan* p, float* buf, size_t count, cudaStream_t st);   // no-op without a communicator
int launch_assemble(Plan* p, const float* x, cudaStream_t st);
int launch_init_state(Plan* p, float* mu, float* lv, float* h, float* c, cudaStream_t st);
void train_free(Plan* p);

// ------------------------------------------------------------------ conv_tc.cu (tcgen05 path)
int tc_supported(const Plan* p);      // 1 if the bf16 tensor-core path can run this shape
int tc_alloc(Plan* p);
void tc_free(Plan* p);
int tc_on_workspace(Plan* p);         // (re)build TMA descriptors over the workspace
int tc_setup_weights(Plan* p, const IodineWeights* w, cudaStream_t st);
// layer l (1..n-1) forward: out = ELU(conv(in) + b); dgrad: out = convT(in) * ELU'(act_prev),
// or, when G != nullptr, the class-wise pixel sums of that product (nothing stored).
int tc_launch_conv(Plan* p, int layer, bool dgrad, const void* in, const void* act_prev,
                   void* out, float* G, cudaStream_t st);
int tc_launch_out4(Plan* p, const void* in, float* out4, cudaStream_t st);
int tc_launch_dgrad_in4(Plan* p, const float* seed4, const void* act_prev, void* gout, cudaStream_t st);
int tc_export_f32(Plan* p, const void* src_bf16, float* dst, size_t n, cudaStream_t st);
int tc_export_seed(Plan* p, const void* seed8, float* dst, cudaStream_t st);
// collapsed first decoder layer into the chunk-planar bf16 layout, and the class-wise pixel
// sums of the last data-gradient (its inverse)
int tc_launch_layer1(Plan* p, void* act0, cudaStream_t st);
int tc_launch_class_sum(Plan* p, const void* g, cudaStream_t st);
int tc_rs_worklist(const Plan* p, int which, const int4** itab, const int32_t** coff, int* grid, const void** zero_row);
// wgrad_tc.cu: tcgen05 weight gradients of the decoder's 3x3 layers (tensor-core modes, C = 64, W = 128)
int wgrad_tc_supported(const Plan* p);
int wgrad_to_h16(Plan* p, const void* src, void* dst, int channels, cudaStream_t st);   // tf32 planes -> fp16 planes
int launch_wgrad_tc(Plan* p, const void* act_prev, const void* g, float* dw, float coef, cudaStream_t st);
int launch_wgrad_tc_out4(Plan* p, const void* act_last, const void* seed8, float* dw, float coef, cudaStream_t st);
int tc_dsum_fused(const Plan* p);   // 1: tc_launch_conv(..., G) of the last data-gradient layer produces the class sums itself

// ------------------------------------------------------------------ refine_tc.cu (tcgen05 refinement encoder)
int rtc_supported(const Plan* p);
int rtc_alloc(Plan* p);
void rtc_free(Plan* p);
bool rtc_enabled(const Plan* p);       // 16-bit mode and every refine layer fits the tensor-core kernel
bool rtc_fused_aux(const Plan* p);     // layer 0 normalises / packs the aux stack itself (no assemble16 pass, no enc16)
int rtc_setup_weights(Plan* p, const IodineWeights* w, cudaStream_t st);
int rtc_launch_refine_convs(Plan* p, cudaStream_t st);   // p->enc16 -> p->pool

}  // namespace iod
