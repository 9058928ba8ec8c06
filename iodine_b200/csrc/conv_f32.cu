// conv_f32.cu -- exact fp32 (FFMA) convolutions of the decoder and the refinement encoder.
//
// Replaces: MultiLayerConv.forward (reference lib/modeling/iodine.py:586-594), Decoder.conv
// (iodine.py:422,435) and the autograd convolution_backward that (B*elbo).backward()
// (iodine.py:90) triggers for the decoder -- data-gradient only, the weight gradients the
// reference also computes are never read at inference and are not computed here.
//
// Layout: activations are NHWC ([slot-image, y, x, channel]); weights are repacked once by
// head.cu::setup kernels into [chunk][tap][ci][co].  These kernels are the IODINE_FP32
// precision path and the checker for the tcgen05 path (conv_tc.cu).
#include "common.cuh"

namespace iod {

// =====================================================================================
// stride-1 CIN -> COUT conv (COUT multiple of 8), tile 4 rows x 32 cols, 256 threads.
// MODE 0: out = ELU(acc + bias)                       (forward)
// MODE 1: out = acc * ELU'(act_prev)                  (data-gradient, weights pre-flipped)
// MODE 2: G[n][class][co] += sum_pixels acc*ELU'(act) (data-gradient into layer 1, reduced
//         over the border classes of the broadcast-collapsed first layer; nothing stored)
// =====================================================================================
template <int CIN, int KS>
struct CcCfg {
  static constexpr int CK = (CIN < 16) ? CIN : (KS == 3 ? 16 : 8);
  static constexpr int TH = 4, TW = 32;
  static constexpr int HWC = TW + KS - 1;           // halo cols
  static constexpr int HR = TH + KS - 1;            // halo rows
  static constexpr int HP = HR * HWC;
  // pad the per-channel stride so the transposing fill is bank-conflict-light
  static constexpr int HPS = HP + ((CK / 4 >= 4) ? ((10 - HP % 8) % 8) : ((12 - HP % 8) % 8));
};

template <int CIN, int COUT, int KS, int MODE>
__global__ void __launch_bounds__(256)
conv_cc_kernel(const float* __restrict__ in, const float* __restrict__ wpack,
               const float* __restrict__ bias, const float* __restrict__ actp,
               float* __restrict__ out, float* __restrict__ G, int H, int W) {
  using Cfg = CcCfg<CIN, KS>;
  constexpr int CK = Cfg::CK, TH = Cfg::TH, TW = Cfg::TW, HWC = Cfg::HWC, HP = Cfg::HP,
                HPS = Cfg::HPS;
  constexpr int COG = COUT / 8;          // co groups of 8
  constexpr int WPC = 8 / COG;           // warps per co group
  constexpr int RPT = TH / WPC;          // rows per thread
  static_assert(COG >= 2 && COG <= 8 && RPT >= 1, "unsupported COUT");
  constexpr int P = KS / 2;
  constexpr int NCHUNK = CIN / CK;

  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                    // [CK][HPS]
  float* s_w = smem + CK * HPS;          // [KS*KS][CK][COUT]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cog = warp % COG, row0 = (warp / COG) * RPT;
  const int n = blockIdx.z, ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const float* in_n = in + (size_t)n * H * W * CIN;

  float acc[RPT][8];
#pragma unroll
  for (int j = 0; j < RPT; ++j)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[j][q] = 0.f;

  for (int ch = 0; ch < NCHUNK; ++ch) {
    __syncthreads();
    // ---- input halo tile, transposed to [ci][pixel]
    for (int i = tid; i < HP * (CK / 4); i += 256) {
      const int pix = i / (CK / 4), cg = i % (CK / 4);
      const int hy = pix / HWC, hx = pix % HWC;
      const int y = ty0 - P + hy, x = tx0 - P + hx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < H && x >= 0 && x < W)
        v = *reinterpret_cast<const float4*>(in_n + ((size_t)y * W + x) * CIN + ch * CK + cg * 4);
      s_in[(cg * 4 + 0) * HPS + pix] = v.x;
      s_in[(cg * 4 + 1) * HPS + pix] = v.y;
      s_in[(cg * 4 + 2) * HPS + pix] = v.z;
      s_in[(cg * 4 + 3) * HPS + pix] = v.w;
    }
    // ---- weight chunk [tap][ci][co], straight copy
    {
      const float4* src = reinterpret_cast<const float4*>(wpack + (size_t)ch * KS * KS * CK * COUT);
      float4* dst = reinterpret_cast<float4*>(s_w);
      for (int i = tid; i < KS * KS * CK * COUT / 4; i += 256) dst[i] = src[i];
    }
    __syncthreads();

#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        float col[RPT + KS - 1];
#pragma unroll
        for (int r = 0; r < RPT + KS - 1; ++r)
          col[r] = s_in[ci * HPS + (row0 + r) * HWC + lane + dx];
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          const float4 w0 = *reinterpret_cast<const float4*>(
              s_w + ((dy * KS + dx) * CK + ci) * COUT + cog * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(
              s_w + ((dy * KS + dx) * CK + ci) * COUT + cog * 8 + 4);
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const float v = col[j + dy];
            acc[j][0] = fmaf(v, w0.x, acc[j][0]);
            acc[j][1] = fmaf(v, w0.y, acc[j][1]);
            acc[j][2] = fmaf(v, w0.z, acc[j][2]);
            acc[j][3] = fmaf(v, w0.w, acc[j][3]);
            acc[j][4] = fmaf(v, w1.x, acc[j][4]);
            acc[j][5] = fmaf(v, w1.y, acc[j][5]);
            acc[j][6] = fmaf(v, w1.z, acc[j][6]);
            acc[j][7] = fmaf(v, w1.w, acc[j][7]);
          }
        }
      }
    }
  }

  // ------------------------------------------------------------------ epilogue
  const int x = tx0 + lane;
  const int co0 = cog * 8;
  if (MODE == 0) {
    float bv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) bv[q] = bias[co0 + q];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int y = ty0 + row0 + j;
      if (y < H && x < W) {
        float* o = out + (((size_t)n * H + y) * W + x) * COUT + co0;
        float4 a, b;
        a.x = elu_f(acc[j][0] + bv[0]); a.y = elu_f(acc[j][1] + bv[1]);
        a.z = elu_f(acc[j][2] + bv[2]); a.w = elu_f(acc[j][3] + bv[3]);
        b.x = elu_f(acc[j][4] + bv[4]); b.y = elu_f(acc[j][5] + bv[5]);
        b.z = elu_f(acc[j][6] + bv[6]); b.w = elu_f(acc[j][7] + bv[7]);
        *reinterpret_cast<float4*>(o) = a;
        *reinterpret_cast<float4*>(o + 4) = b;
      }
    }
  } else if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int y = ty0 + row0 + j;
      if (y < H && x < W) {
        const size_t idx = (((size_t)n * H + y) * W + x) * COUT + co0;
        const float4 a0 = *reinterpret_cast<const float4*>(actp + idx);
        const float4 a1 = *reinterpret_cast<const float4*>(actp + idx + 4);
        float4 a, b;
        a.x = acc[j][0] * elu_grad_from_act(a0.x); a.y = acc[j][1] * elu_grad_from_act(a0.y);
        a.z = acc[j][2] * elu_grad_from_act(a0.z); a.w = acc[j][3] * elu_grad_from_act(a0.w);
        b.x = acc[j][4] * elu_grad_from_act(a1.x); b.y = acc[j][5] * elu_grad_from_act(a1.y);
        b.z = acc[j][6] * elu_grad_from_act(a1.z); b.w = acc[j][7] * elu_grad_from_act(a1.w);
        *reinterpret_cast<float4*>(out + idx) = a;
        *reinterpret_cast<float4*>(out + idx + 4) = b;
      }
    }
  } else {
    const int ncls = KS * KS;
    const int cx = border_class(x < W ? x : 0, W, P);
    const bool xin = x < W;
    float si[8];
    int cy_run = -1;
#pragma unroll
    for (int q = 0; q < 8; ++q) si[q] = 0.f;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int y = ty0 + row0 + j;
      const bool yin = y < H;                     // warp-uniform
      const int cy = border_class(yin ? y : 0, H, P);
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = 0.f;
      if (yin && xin) {
        const size_t idx = (((size_t)n * H + y) * W + x) * COUT + co0;
        const float4 a0 = *reinterpret_cast<const float4*>(actp + idx);
        const float4 a1 = *reinterpret_cast<const float4*>(actp + idx + 4);
        v[0] = acc[j][0] * elu_grad_from_act(a0.x); v[1] = acc[j][1] * elu_grad_from_act(a0.y);
        v[2] = acc[j][2] * elu_grad_from_act(a0.z); v[3] = acc[j][3] * elu_grad_from_act(a0.w);
        v[4] = acc[j][4] * elu_grad_from_act(a1.x); v[5] = acc[j][5] * elu_grad_from_act(a1.y);
        v[6] = acc[j][6] * elu_grad_from_act(a1.z); v[7] = acc[j][7] * elu_grad_from_act(a1.w);
        if (cx != P) {                            // image-border columns: few lanes, direct
#pragma unroll
          for (int q = 0; q < 8; ++q)
            atomicAdd(G + ((size_t)n * ncls + cy * KS + cx) * COUT + co0 + q, v[q]);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = 0.f;
        }
      }
      if (yin) {
        if (cy != cy_run && cy_run >= 0) {        // flush the run of rows with equal class
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float s = warp_sum(si[q]);
            if (lane == 0) atomicAdd(G + ((size_t)n * ncls + cy_run * KS + P) * COUT + co0 + q, s);
            si[q] = 0.f;
          }
        }
        cy_run = cy;
#pragma unroll
        for (int q = 0; q < 8; ++q) si[q] += v[q];
      }
    }
    if (cy_run >= 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float s = warp_sum(si[q]);
        if (lane == 0) atomicAdd(G + ((size_t)n * ncls + cy_run * KS + P) * COUT + co0 + q, s);
      }
    }
  }
}

template <int CIN, int COUT, int KS, int MODE>
static int launch_cc_t(Plan* p, const float* in, const float* w, const float* bias,
                       const float* actp, float* out, float* G, cudaStream_t st) {
  using Cfg = CcCfg<CIN, KS>;
  const size_t smem = (size_t)(Cfg::CK * Cfg::HPS + KS * KS * Cfg::CK * COUT) * sizeof(float);
  auto kern = conv_cc_kernel<CIN, COUT, KS, MODE>;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((p->s.W + Cfg::TW - 1) / Cfg::TW, (p->s.H + Cfg::TH - 1) / Cfg::TH, p->BK);
  kern<<<grid, 256, smem, st>>>(in, w, bias, actp, out, G, p->s.H, p->s.W);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

template <int C, int KS>
static int launch_cc_mode(Plan* p, const float* in, const float* w, const float* bias,
                          const float* actp, float* out, float* G, int mode, cudaStream_t st) {
  switch (mode) {
    case 0: return launch_cc_t<C, C, KS, 0>(p, in, w, bias, actp, out, G, st);
    case 1: return launch_cc_t<C, C, KS, 1>(p, in, w, bias, actp, out, G, st);
    default: return launch_cc_t<C, C, KS, 2>(p, in, w, bias, actp, out, G, st);
  }
}

int launch_conv_cc(Plan* p, const float* in, const float* wpack, const float* bias,
                   const float* act_prev, float* out, float* G, int mode, cudaStream_t st) {
  const int C = p->C, KS = p->s.dec_k;
#define IOD_CASE(c, k) \
  if (C == c && KS == k) return launch_cc_mode<c, k>(p, in, wpack, bias, act_prev, out, G, mode, st);
  IOD_CASE(16, 3) IOD_CASE(32, 3) IOD_CASE(64, 3)
  IOD_CASE(16, 5) IOD_CASE(32, 5) IOD_CASE(64, 5)
#undef IOD_CASE
  set_error("conv_cc: unsupported dec_chan=%d dec_k=%d", C, KS);
  return 1;
}

// 4 -> C data-gradient of decoder.conv: g = convT(seed4, Wout) * ELU'(act_last)
// store: always write the gradient (the training step reduces it itself), also for a one-layer decoder
int launch_dgrad_in4(Plan* p, const float* seed4, const float* act_prev, float* gout,
                     cudaStream_t st, bool store) {
  const int C = p->C, KS = p->s.dec_k;
  const int mode = (p->s.dec_layers == 1 && !store) ? 2 : 1;
#define IOD_CASE(c, k)                                                                     \
  if (C == c && KS == k) {                                                                 \
    if (mode == 1) return launch_cc_t<4, c, k, 1>(p, seed4, p->out_wt, nullptr, act_prev, gout, nullptr, st); \
    return launch_cc_t<4, c, k, 2>(p, seed4, p->out_wt, nullptr, act_prev, nullptr, p->G, st); \
  }
  IOD_CASE(16, 3) IOD_CASE(32, 3) IOD_CASE(64, 3)
  IOD_CASE(16, 5) IOD_CASE(32, 5) IOD_CASE(64, 5)
#undef IOD_CASE
  set_error("dgrad_in4: unsupported dec_chan=%d dec_k=%d", C, KS);
  return 1;
}

// =====================================================================================
// decoder.conv forward: C -> 4, tile 32x32 pixels, thread = 4 rows x 1 col x 4 outputs
// =====================================================================================
template <int CIN, int KS>
__global__ void __launch_bounds__(256)
conv_out4_kernel(const float* __restrict__ in, const float* __restrict__ wpk,  // [tap][ci][4]
                 const float* __restrict__ bias, float* __restrict__ out4, int H, int W) {
  constexpr int CK = 8, TH = 32, TW = 32, HWC = TW + KS - 1, HR = TH + KS - 1, HP = HR * HWC;
  constexpr int HPS = HP + ((12 - HP % 8) % 8);
  constexpr int P = KS / 2, RPT = 4;
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                 // [CK][HPS]
  float* s_w = smem + CK * HPS;       // [tap][CK][4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = warp * RPT;
  const int n = blockIdx.z, ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const float* in_n = in + (size_t)n * H * W * CIN;
  float acc[RPT][4];
#pragma unroll
  for (int j = 0; j < RPT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  for (int ch = 0; ch < CIN / CK; ++ch) {
    __syncthreads();
    for (int i = tid; i < HP * (CK / 4); i += 256) {
      const int pix = i / (CK / 4), cg = i % (CK / 4);
      const int hy = pix / HWC, hx = pix % HWC;
      const int y = ty0 - P + hy, x = tx0 - P + hx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < H && x >= 0 && x < W)
        v = *reinterpret_cast<const float4*>(in_n + ((size_t)y * W + x) * CIN + ch * CK + cg * 4);
      s_in[(cg * 4 + 0) * HPS + pix] = v.x;
      s_in[(cg * 4 + 1) * HPS + pix] = v.y;
      s_in[(cg * 4 + 2) * HPS + pix] = v.z;
      s_in[(cg * 4 + 3) * HPS + pix] = v.w;
    }
    for (int i = tid; i < KS * KS * CK; i += 256) {
      const int tap = i / CK, ci = i % CK;
      reinterpret_cast<float4*>(s_w)[i] =
          *reinterpret_cast<const float4*>(wpk + ((size_t)tap * CIN + ch * CK + ci) * 4);
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        float col[RPT + KS - 1];
#pragma unroll
        for (int r = 0; r < RPT + KS - 1; ++r)
          col[r] = s_in[ci * HPS + (row0 + r) * HWC + lane + dx];
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          const float4 w = reinterpret_cast<const float4*>(s_w)[(dy * KS + dx) * CK + ci];
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const float v = col[j + dy];
            acc[j][0] = fmaf(v, w.x, acc[j][0]);
            acc[j][1] = fmaf(v, w.y, acc[j][1]);
            acc[j][2] = fmaf(v, w.z, acc[j][2]);
            acc[j][3] = fmaf(v, w.w, acc[j][3]);
          }
        }
      }
    }
  }
  const int x = tx0 + lane;
  const float b0 = bias[0], b1 = bias[1], b2 = bias[2], b3 = bias[3];
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int y = ty0 + row0 + j;
    if (y < H && x < W)
      reinterpret_cast<float4*>(out4)[((size_t)n * H + y) * W + x] =
          make_float4(acc[j][0] + b0, acc[j][1] + b1, acc[j][2] + b2, acc[j][3] + b3);
  }
}

template <int CIN, int KS>
static int launch_out4_t(Plan* p, const float* in, float* out4, cudaStream_t st) {
  constexpr int CK = 8, HP = (32 + KS - 1) * (32 + KS - 1), HPS = HP + ((12 - HP % 8) % 8);
  const size_t smem = (size_t)(CK * HPS + KS * KS * CK * 4) * sizeof(float);
  auto kern = conv_out4_kernel<CIN, KS>;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((p->s.W + 31) / 32, (p->s.H + 31) / 32, p->BK);
  kern<<<grid, 256, smem, st>>>(in, p->out_w, p->out_b, out4, p->s.H, p->s.W);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

int launch_conv_out4(Plan* p, const float* in, float* out4, cudaStream_t st) {
  const int C = p->C, KS = p->s.dec_k;
#define IOD_CASE(c, k) if (C == c && KS == k) return launch_out4_t<c, k>(p, in, out4, st);
  IOD_CASE(16, 3) IOD_CASE(32, 3) IOD_CASE(64, 3)
  IOD_CASE(16, 5) IOD_CASE(32, 5) IOD_CASE(64, 5)
#undef IOD_CASE
  set_error("conv_out4: unsupported dec_chan=%d dec_k=%d", C, KS);
  return 1;
}

// =====================================================================================
// refinement encoder: strided CIN -> COUT conv + ELU, NHWC, 128 output pixels per block
// (RefinementNetwork.forward, iodine.py:480; MultiLayerConv with stride, 583)
// =====================================================================================
template <int COUT, int KS, int S, int TW>
__global__ void __launch_bounds__(256)
conv_strided_kernel(const float* __restrict__ in, const float* __restrict__ wpk,  // [tap][cin][COUT]
                    const float* __restrict__ bias, float* __restrict__ out,
                    int CIN, int Hin, int Win, int Hout, int Wout) {
  constexpr int CK = 8, TH = 128 / TW;
  constexpr int HWC = (TW - 1) * S + KS, HR = (TH - 1) * S + KS, HP = HR * HWC;
  constexpr int HPS = HP + ((12 - HP % 8) % 8);
  constexpr int P = KS / 2;
  constexpr int COG = COUT / 8, WPC = 8 / COG, PPT = 4 / WPC;   // pixels per thread
  static_assert(PPT >= 1, "COUT too small");
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                 // [CK][HPS]
  float* s_w = smem + CK * HPS;       // [tap][CK][COUT]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cog = warp % COG, pgrp = warp / COG;
  const int n = blockIdx.z, ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const float* in_n = in + (size_t)n * Hin * Win * CIN;
  const int iy0 = ty0 * S - P, ix0 = tx0 * S - P;

  int poff[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int pidx = (pgrp * PPT + j) * 32 + lane;      // 0..127
    poff[j] = (pidx / TW) * S * HWC + (pidx % TW) * S;
  }
  float acc[PPT][8];
#pragma unroll
  for (int j = 0; j < PPT; ++j)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[j][q] = 0.f;

  const int nchunk = (CIN + CK - 1) / CK;
  for (int ch = 0; ch < nchunk; ++ch) {
    __syncthreads();
    for (int i = tid; i < HP * (CK / 4); i += 256) {
      const int pix = i / (CK / 4), cg = i % (CK / 4);
      const int hy = pix / HWC, hx = pix % HWC;
      const int y = iy0 + hy, x = ix0 + hx;
      const int c = ch * CK + cg * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < Hin && x >= 0 && x < Win && c < CIN)
        v = *reinterpret_cast<const float4*>(in_n + ((size_t)y * Win + x) * CIN + c);
      s_in[(cg * 4 + 0) * HPS + pix] = v.x;
      s_in[(cg * 4 + 1) * HPS + pix] = v.y;
      s_in[(cg * 4 + 2) * HPS + pix] = v.z;
      s_in[(cg * 4 + 3) * HPS + pix] = v.w;
    }
    for (int i = tid; i < KS * KS * CK * COUT / 4; i += 256) {
      const int co4 = i % (COUT / 4), ci = (i / (COUT / 4)) % CK, tap = i / (COUT / 4) / CK;
      const int c = ch * CK + ci;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < CIN) v = *reinterpret_cast<const float4*>(wpk + ((size_t)tap * CIN + c) * COUT + co4 * 4);
      reinterpret_cast<float4*>(s_w)[i] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float4 w0 = *reinterpret_cast<const float4*>(
              s_w + ((dy * KS + dx) * CK + ci) * COUT + cog * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(
              s_w + ((dy * KS + dx) * CK + ci) * COUT + cog * 8 + 4);
#pragma unroll
          for (int j = 0; j < PPT; ++j) {
            const float v = s_in[ci * HPS + poff[j] + dy * HWC + dx];
            acc[j][0] = fmaf(v, w0.x, acc[j][0]);
            acc[j][1] = fmaf(v, w0.y, acc[j][1]);
            acc[j][2] = fmaf(v, w0.z, acc[j][2]);
            acc[j][3] = fmaf(v, w0.w, acc[j][3]);
            acc[j][4] = fmaf(v, w1.x, acc[j][4]);
            acc[j][5] = fmaf(v, w1.y, acc[j][5]);
            acc[j][6] = fmaf(v, w1.z, acc[j][6]);
            acc[j][7] = fmaf(v, w1.w, acc[j][7]);
          }
        }
      }
    }
  }
  float bv[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) bv[q] = bias[cog * 8 + q];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int pidx = (pgrp * PPT + j) * 32 + lane;
    const int y = ty0 + pidx / TW, x = tx0 + pidx % TW;
    if (y < Hout && x < Wout) {
      float* o = out + (((size_t)n * Hout + y) * Wout + x) * COUT + cog * 8;
      float4 a, b;
      a.x = elu_f(acc[j][0] + bv[0]); a.y = elu_f(acc[j][1] + bv[1]);
      a.z = elu_f(acc[j][2] + bv[2]); a.w = elu_f(acc[j][3] + bv[3]);
      b.x = elu_f(acc[j][4] + bv[4]); b.y = elu_f(acc[j][5] + bv[5]);
      b.z = elu_f(acc[j][6] + bv[6]); b.w = elu_f(acc[j][7] + bv[7]);
      *reinterpret_cast<float4*>(o) = a;
      *reinterpret_cast<float4*>(o + 4) = b;
    }
  }
}

template <int COUT, int KS, int S, int TW>
static int launch_strided_t(Plan* p, const float* in, const float* w, const float* b, float* out,
                            int CIN, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
  constexpr int CK = 8, TH = 128 / TW;
  constexpr int HWC = (TW - 1) * S + KS, HR = (TH - 1) * S + KS, HP = HR * HWC;
  constexpr int HPS = HP + ((12 - HP % 8) % 8);
  const size_t smem = (size_t)(CK * HPS + KS * KS * CK * COUT) * sizeof(float);
  auto kern = conv_strided_kernel<COUT, KS, S, TW>;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((Wout + TW - 1) / TW, (Hout + TH - 1) / TH, p->BK);
  kern<<<grid, 256, smem, st>>>(in, w, b, out, CIN, Hin, Win, Hout, Wout);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

template <int COUT, int KS, int S>
static int launch_strided_tw(Plan* p, const float* in, const float* w, const float* b, float* out,
                             int CIN, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
  if (Wout > 16) return launch_strided_t<COUT, KS, S, 32>(p, in, w, b, out, CIN, Hin, Win, Hout, Wout, st);
  if (Wout > 8) return launch_strided_t<COUT, KS, S, 16>(p, in, w, b, out, CIN, Hin, Win, Hout, Wout, st);
  return launch_strided_t<COUT, KS, S, 8>(p, in, w, b, out, CIN, Hin, Win, Hout, Wout, st);
}

static int launch_strided(Plan* p, const float* in, const float* w, const float* b, float* out,
                          int CIN, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
  const int C = p->Cr, KS = p->s.ref_k, S = p->s.ref_stride;
#define IOD_CASE(c, k, s) \
  if (C == c && KS == k && S == s) return launch_strided_tw<c, k, s>(p, in, w, b, out, CIN, Hin, Win, Hout, Wout, st);
  IOD_CASE(16, 3, 2) IOD_CASE(32, 3, 2) IOD_CASE(64, 3, 2)
  IOD_CASE(16, 5, 2) IOD_CASE(32, 5, 2) IOD_CASE(64, 5, 2)
  IOD_CASE(16, 3, 1) IOD_CASE(32, 3, 1) IOD_CASE(64, 3, 1)
#undef IOD_CASE
  set_error("refine conv: unsupported ref_chan=%d ref_k=%d ref_stride=%d", C, KS, S);
  return 1;
}

// global average pool over the last refine activation (F.adaptive_avg_pool2d, iodine.py:481)
__global__ void avgpool_kernel(const float* __restrict__ in, float* __restrict__ pool, int HWo, int C) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < HWo; ++i) s += in[((size_t)n * HWo + i) * C + c];
    pool[(size_t)n * C + c] = s / (float)HWo;
  }
}

// All refine conv layers.  Layer 0 reads the assembled 20-channel (17 + 3 zero pad) input
// that mixture.cu::assemble wrote into rbuf[1]'s tail region (p->enc20).
// outs: optional per-layer output buffers (the training tape keeps every activation); default = ping-pong
int launch_refine_convs(Plan* p, const float* enc20, cudaStream_t st, float* const* outs) {
  const float* cur = enc20;
  int cin = 20;
  for (int l = 0; l < p->s.ref_layers; ++l) {
    float* dst = outs ? outs[l] : p->rbuf[l & 1];
    const float* w = (l == 0) ? p->ref_w0 : p->ref_wp[l];
    if (launch_strided(p, cur, w, p->ref_b[l], dst, cin, p->ref_h[l], p->ref_w[l], p->ref_h[l + 1],
                       p->ref_w[l + 1], st))
      return 1;
    cur = dst;
    cin = p->Cr;
  }
  const int HWo = p->ref_h[p->s.ref_layers] * p->ref_w[p->s.ref_layers];
  avgpool_kernel<<<p->BK, 64, 0, st>>>(cur, p->pool, HWo, p->Cr);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

}  // namespace iod
