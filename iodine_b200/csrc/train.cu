// train.cu -- the TRAINING step: IODINE.forward (reference lib/modeling/iodine.py:115-158) followed by
// loss.backward() (lib/engine/train.py:60-65), every parameter gradient computed by hand-written kernels.
//
// Structure of the reference's graph (oracle/train_restatement.py restates it and is pinned against the reference's
// autograd): loss = -sum_{i=0..T} w_i elbo_i, w_i = (i+1)/(T+1), elbo_i = J_i / B.  Gaussian.update detaches the
// previous posterior (iodine.py:642-643) and get_input_encoding detaches both of its outputs (343), hence
//   * the DECODER weights only see the direct terms c_i dJ_i/dW, c_i = -w_i/B: the seeds and the data-gradient chain
//     of step i are exactly the inference ones, so each layer's weight gradient is accumulated inside step i, right
//     after the data-gradient kernel has produced dJ_i/d(pre-activation) of that layer -- nothing of the decoder is
//     kept across steps;
//   * posterior.init_mean / init_logvar only see step 0;
//   * the REFINER is reached through delta_i (posterior_{i+1} = const + delta_i): incoming gradient
//     c_{i+1} dJ_{i+1}/d(posterior_{i+1}) -- which the loop computes anyway -- plus the LSTM state chain, i.e. one
//     backward sweep over the T refiner calls after the loop, over a tape of their activations (the 17-channel input
//     is a constant: no data-gradient into it).
// With a communicator installed (iodine_plan_set_comm) the ranks hold different images of one global batch: every
// batch mean divides by the GLOBAL size and ONE ncclAllReduce of the flat gradient buffer (all parameters + the loss
// + the ELBO table, ~4.4 MB at the CLEVR6 sizes) replaces DataParallel's gradient reduction to GPU 0.
#include <string.h>

#include "common.cuh"

namespace iod {

// ------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------
struct TrainState {
  void* ws = nullptr;
  size_t need = 0;
  std::vector<float*> enc, u, xin, gates, pool;   // [T] tape of the refiner calls
  std::vector<std::vector<float*>> ract;          // [T][ref_layers] conv activations
  float* hs = nullptr;        // [T+1][BK*M] LSTM h before call t (hs[t]) / after call t (hs[t+1])
  float* cs = nullptr;        // [T+1][BK*M]
  float* pg = nullptr;        // [T+1][BK*2L] dJ_i/d(posterior_i): (dmu | dlogvar) per slot
  float *mu = nullptr, *lv = nullptr;              // [BK*L] running posterior
  float *Gx = nullptr, *Gy = nullptr;              // [BK][n_class][C] first moments of dJ/d(pre-activation 0)
  float *dgates = nullptr, *dx = nullptr, *dpool = nullptr;
  float* dh[2] = {nullptr, nullptr};
  float* dc[2] = {nullptr, nullptr};
  float* rg[2] = {nullptr, nullptr};               // refiner conv gradients, ping-pong
  void *h16_act = nullptr, *h16_g = nullptr;       // IODINE_TF32: fp16 copies of the weight-gradient operands
  float* gflat = nullptr;                          // every parameter gradient + loss + ELBO table
  size_t n_flat = 0;
  // offsets (floats) into gflat, state_dict order
  size_t o_dec_w[IODINE_MAX_LAYERS], o_dec_b[IODINE_MAX_LAYERS], o_out_w, o_out_b;
  size_t o_ref_w[IODINE_MAX_LAYERS], o_ref_b[IODINE_MAX_LAYERS], o_mlp_w, o_mlp_b;
  size_t o_wih, o_whh, o_bih, o_bhh, o_head_w, o_head_b, o_init, o_loss, o_terms;
};

static TrainState* train_state(Plan* p) { return reinterpret_cast<TrainState*>(p->train); }
void train_free(Plan* p) {
  delete train_state(p);
  p->train = nullptr;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t train_carve(Plan* p, TrainState* ts, char* base) {
  size_t off = 0;
  auto take = [&](size_t floats) -> float* {
    float* r = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off = align_up(off + floats * sizeof(float), 1024);
    return r;
  };
  const IodineShape& s = p->s;
  const size_t BK = p->BK, HW = p->HW, L = s.L, M = p->M, Cr = p->Cr, C = p->C, T = s.T;
  ts->enc.assign(T, nullptr); ts->u.assign(T, nullptr); ts->xin.assign(T, nullptr);
  ts->gates.assign(T, nullptr); ts->pool.assign(T, nullptr);
  ts->ract.assign(T, std::vector<float*>(s.ref_layers, nullptr));
  size_t biggest = 1;
  for (size_t t = 0; t < T; ++t) {
    ts->enc[t] = take(BK * HW * 20);
    for (int l = 0; l < s.ref_layers; ++l) {
      const size_t n = BK * (size_t)p->ref_h[l + 1] * p->ref_w[l + 1] * Cr;
      ts->ract[t][l] = take(n);
      if (n > biggest) biggest = n;
    }
    ts->u[t] = take(BK * M);
    ts->xin[t] = take(BK * (M + 4 * L));
    ts->gates[t] = take(BK * 4 * M);
    ts->pool[t] = take(BK * Cr);
  }
  ts->hs = take((T + 1) * BK * M);
  ts->cs = take((T + 1) * BK * M);
  ts->pg = take((T + 1) * BK * 2 * L);
  ts->mu = take(BK * L);
  ts->lv = take(BK * L);
  ts->Gx = take(BK * p->n_class * C);
  ts->Gy = take(BK * p->n_class * C);
  ts->dgates = take(BK * 4 * M);
  ts->dx = take(BK * M);
  ts->dpool = take(BK * Cr);
  for (int i = 0; i < 2; ++i) { ts->dh[i] = take(BK * M); ts->dc[i] = take(BK * M); ts->rg[i] = take(biggest); }
  if (tf_mode(p) && wgrad_tc_supported(p)) {
    ts->h16_act = take(BK * HW * C / 2);
    ts->h16_g = take(BK * HW * C / 2);
  }
  // flat gradient buffer
  size_t o = 0;
  const size_t kk = (size_t)s.dec_k * s.dec_k, rkk = (size_t)s.ref_k * s.ref_k;
  for (int l = 0; l < s.dec_layers; ++l) { ts->o_dec_w[l] = o; o += C * (l == 0 ? L + 2 : C) * kk; }
  for (int l = 0; l < s.dec_layers; ++l) { ts->o_dec_b[l] = o; o += C; }
  ts->o_out_w = o; o += 4 * C * kk;
  ts->o_out_b = o; o += 4;
  for (int l = 0; l < s.ref_layers; ++l) { ts->o_ref_w[l] = o; o += Cr * (l == 0 ? 17 : Cr) * rkk; }
  for (int l = 0; l < s.ref_layers; ++l) { ts->o_ref_b[l] = o; o += Cr; }
  ts->o_mlp_w = o; o += M * Cr;
  ts->o_mlp_b = o; o += M;
  ts->o_wih = o; o += 4 * M * (M + 4 * L);
  ts->o_whh = o; o += 4 * M * M;
  ts->o_bih = o; o += 4 * M;
  ts->o_bhh = o; o += 4 * M;
  ts->o_head_w = o; o += 2 * L * M;            // mean_update.weight rows, then logvar_update.weight rows
  ts->o_head_b = o; o += 2 * L;
  ts->o_init = o; o += 2 * L;                  // init_mean, init_logvar
  ts->o_loss = o; o += 1;
  ts->o_terms = o; o += 2 * (T + 1);
  ts->n_flat = o;
  ts->gflat = take(o);
  return off;
}

// ------------------------------------------------------------------------------------------------
// layouts: the decoder's activations / gradients are NHWC fp32 (IODINE_FP32) or chunk-planar (tensor-core modes:
// planes of 8 16-bit channels or of 4 tf32 channels, 16 bytes per pixel and plane); the refiner's tape is NHWC fp32
// ------------------------------------------------------------------------------------------------
enum Lay { LAY_NHWC = 0, LAY_P16 = 1, LAY_PTF = 2 };
struct TView {
  const void* base;
  int lay;        // Lay
  int f16;        // LAY_P16: 1 = IEEE half, 0 = bfloat16
  int C;          // channels (LAY_NHWC: valid channels)
  int pitch;      // LAY_NHWC: floats per pixel (multiple of 4, >= C; padding holds zeros); planar: the REAL channels (<= C)
};
// four channels c0..c0+3 (c0 % 4 == 0) of pixel `pix` of slot-image n; channels >= C read as zero
__device__ __forceinline__ float4 load4(const TView& v, int n, int HW, int pix, int c0) {
  if (c0 >= v.C) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (v.lay == LAY_NHWC)
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(v.base) + ((size_t)n * HW + pix) * v.pitch + c0));
  if (v.lay == LAY_PTF)
    return __ldg(reinterpret_cast<const float4*>(v.base) + ((size_t)n * (v.C >> 2) + (c0 >> 2)) * HW + pix);
  const uint2 q = __ldg(reinterpret_cast<const uint2*>(v.base) + (((size_t)n * (v.C >> 3) + (c0 >> 3)) * HW + pix) * 2 + ((c0 >> 2) & 1));
  const float2 a = unpack_h2(q.x, v.f16), b = unpack_h2(q.y, v.f16);
  return make_float4(a.x, a.y, b.x, b.y);
}
// the 4-channel seed: fp32 [n][pix][4] (IODINE_FP32 / IODINE_TF32) or 16-bit [n][pix][8] with 4 real channels
__device__ __forceinline__ float4 load_seed(const void* seed, int half /*0 f32, 1 bf16, 2 f16*/, size_t sp) {
  if (!half) return __ldg(reinterpret_cast<const float4*>(seed) + sp);
  const uint2 q = __ldg(reinterpret_cast<const uint2*>(seed) + sp * 2);
  const float2 a = unpack_h2(q.x, half == 2), b = unpack_h2(q.y, half == 2);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------
// convolution weight gradient:  dW[co][ci][dy][dx] += coef * sum_{n,oy,ox} g[n,oy,ox,co] * in[n, oy*S+dy-P, ox*S+dx-P, ci]
// (torch.nn.grad.conv2d_weight; zero padding P = KS/2).  One block = one tap, one 64 x (16*TCI) tile of (co, ci)
// and one share of the output pixels; a 4 x TCI register tile per thread, 16 pixels staged per round.
// ------------------------------------------------------------------------------------------------
struct WgParams {
  TView in, g;
  float* dw;                 // [COUT][CIN][KS][KS] (PyTorch)
  float coef;
  int N, Hin, Win, Hout, Wout, CIN, COUT, KS, S;
  int px_per_block;          // multiple of 16
  int ci_tiles;
};

template <int TCI>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgParams q) {
  __shared__ __align__(16) float sg[16][64 + 4];
  __shared__ __align__(16) float si[16][16 * TCI + 4];
  const int tap = blockIdx.x, dy = tap / q.KS, dx = tap - dy * q.KS, P = q.KS / 2;
  const int co_t = blockIdx.z / q.ci_tiles, ci_t = blockIdx.z - co_t * q.ci_tiles;
  const int co0 = co_t * 64, ci0 = ci_t * 16 * TCI;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int HWo = q.Hout * q.Wout, HWi = q.Hin * q.Win;
  const long long total = (long long)q.N * HWo;
  const long long p_lo = (long long)blockIdx.y * q.px_per_block;
  const long long p_hi = (p_lo + q.px_per_block < total) ? p_lo + q.px_per_block : total;
  float acc[4][TCI];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TCI; ++j) acc[i][j] = 0.f;
  const int lp = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;      // loader role: pixel of the round, channel group
  auto fetch = [&](long long p0, float4& gv, float4& iv) {
    gv = make_float4(0.f, 0.f, 0.f, 0.f);
    iv = gv;
    const long long op = p0 + lp;
    if (op < p_hi) {
      const int n = (int)(op / HWo), r = (int)(op - (long long)n * HWo);
      const int oy = r / q.Wout, ox = r - oy * q.Wout;
      if (co0 + lc < q.COUT) gv = load4(q.g, n, HWo, r, co0 + lc);
      const int iy = oy * q.S + dy - P, ix = ox * q.S + dx - P;
      if (lc < 16 * TCI && iy >= 0 && iy < q.Hin && ix >= 0 && ix < q.Win) iv = load4(q.in, n, HWi, iy * q.Win + ix, ci0 + lc);
    }
  };
  float4 gv, iv;
  fetch(p_lo, gv, iv);
  for (long long p0 = p_lo; p0 < p_hi; p0 += 16) {
    __syncthreads();
    *reinterpret_cast<float4*>(&sg[lp][lc]) = gv;
    if (lc < 16 * TCI) *reinterpret_cast<float4*>(&si[lp][lc]) = iv;
    __syncthreads();
    if (p0 + 16 < p_hi) fetch(p0 + 16, gv, iv);        // next round's loads fly under this round's FMAs
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sg[kk][ty * 4]);
      float b[TCI];
#pragma unroll
      for (int j = 0; j < TCI; ++j) b[j] = si[kk][tx * TCI + j];
#pragma unroll
      for (int j = 0; j < TCI; ++j) {
        acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
        acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
        acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
        acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= q.COUT) continue;
#pragma unroll
    for (int j = 0; j < TCI; ++j) {
      const int ci = ci0 + tx * TCI + j;
      if (ci < q.CIN) atomicAdd(q.dw + (((size_t)co * q.CIN + ci) * q.KS + dy) * q.KS + dx, q.coef * acc[i][j]);
    }
  }
}

static int launch_conv_wgrad(Plan* p, const TView& in, const TView& g, float* dw, float coef, int N, int Hin, int Win,
                             int Hout, int Wout, int CIN, int COUT, int KS, int S, cudaStream_t st) {
  WgParams q;
  q.in = in; q.g = g; q.dw = dw; q.coef = coef;
  q.N = N; q.Hin = Hin; q.Win = Win; q.Hout = Hout; q.Wout = Wout; q.CIN = CIN; q.COUT = COUT; q.KS = KS; q.S = S;
  const bool narrow = CIN <= 32;
  const int tci = narrow ? 32 : 64;
  q.ci_tiles = (CIN + tci - 1) / tci;
  const int tiles = q.ci_tiles * ((COUT + 63) / 64);
  const long long total = (long long)N * Hout * Wout;
  // enough blocks for ~3 per SM, at least 256 pixels each
  long long splits = (3LL * p->num_sms + (long long)KS * KS * tiles - 1) / ((long long)KS * KS * tiles);
  if (splits < 1) splits = 1;
  long long per = (total + splits - 1) / splits;
  if (per < 256) per = 256;
  per = (per + 15) / 16 * 16;
  q.px_per_block = (int)per;
  dim3 grid(KS * KS, (unsigned)((total + per - 1) / per), tiles);
  if (narrow) conv_wgrad_kernel<2><<<grid, 256, 0, st>>>(q);
  else conv_wgrad_kernel<4><<<grid, 256, 0, st>>>(q);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// Weight gradient of the refiner's FIRST conv (17 input channels in a 20-float NHWC record, 3x3, stride 2, 64 output
// channels): the generic kernel wastes half of its 64 x 32 tile on it and re-reads g once per tap (2.0 ms per call).
// Here one block accumulates ALL nine taps: a round = 16 consecutive output pixels of one output row; the g tile
// [16][64] and the 3 x 33-pixel input patch they touch are staged once, thread = (output channel, quarter of the 45
// (tap, 4-channel group) pairs), 45-48 running sums in registers, 4 FMAs per shared-memory float4.
__global__ void __launch_bounds__(256)
refine_wgrad_l0_kernel(const float* __restrict__ enc20, const float* __restrict__ g, float* __restrict__ dw, int N, int Hin,
                       int Win, int Hout, int Wout, int rounds_per_block) {
  __shared__ __align__(16) float sg[16][64];
  __shared__ __align__(16) float sp[3][33][20];
  const int co = threadIdx.x & 63, tq = threadIdx.x >> 6;
  const int segs = Wout / 16;
  const long long total = (long long)N * Hout * segs;
  const long long r_lo = (long long)blockIdx.x * rounds_per_block;
  const long long r_hi = (r_lo + rounds_per_block < total) ? r_lo + rounds_per_block : total;
  float acc[12][4];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  // loader roles: g tile = 256 float4 (one per thread); patch = 3 * 33 * 5 = 495 float4 (two per thread)
  float4 gv, pv0, pv1;
  auto fetch = [&](long long rr) {
    const int n = (int)(rr / ((long long)Hout * segs));
    const int r = (int)(rr - (long long)n * Hout * segs);
    const int oy = r / segs, ox0 = (r - oy * segs) * 16;
    gv = __ldg(reinterpret_cast<const float4*>(g + (((size_t)n * Hout + oy) * Wout + ox0 + (threadIdx.x >> 4)) * 64 + (threadIdx.x & 15) * 4));
    auto patch = [&](int idx) -> float4 {
      if (idx >= 495) return make_float4(0.f, 0.f, 0.f, 0.f);
      const int c4 = idx % 5, j = (idx / 5) % 33, dy = idx / 165;
      const int iy = 2 * oy - 1 + dy, ix = 2 * ox0 - 1 + j;
      if (iy < 0 || iy >= Hin || ix < 0 || ix >= Win) return make_float4(0.f, 0.f, 0.f, 0.f);
      return __ldg(reinterpret_cast<const float4*>(enc20 + (((size_t)n * Hin + iy) * Win + ix) * 20 + c4 * 4));
    };
    pv0 = patch(threadIdx.x);
    pv1 = patch(threadIdx.x + 256);
  };
  if (r_lo < r_hi) fetch(r_lo);
  for (long long rr = r_lo; rr < r_hi; ++rr) {
    __syncthreads();
    *reinterpret_cast<float4*>(&sg[threadIdx.x >> 4][(threadIdx.x & 15) * 4]) = gv;
    reinterpret_cast<float4*>(&sp[0][0][0])[threadIdx.x] = pv0;
    if (threadIdx.x + 256 < 495) reinterpret_cast<float4*>(&sp[0][0][0])[threadIdx.x + 256] = pv1;
    __syncthreads();
    if (rr + 1 < r_hi) fetch(rr + 1);              // the next round's loads fly under this round's FMAs
#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {
      const float gval = sg[kk][co];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int q = tq + 4 * i;                  // (tap, 4-channel group) pair of this thread
        if (q < 45) {
          const int tap = q / 5, c4 = q - tap * 5, dy = tap / 3, dx = tap - dy * 3;
          const float4 v = *reinterpret_cast<const float4*>(&sp[dy][2 * kk + dx][c4 * 4]);
          acc[i][0] = fmaf(gval, v.x, acc[i][0]);
          acc[i][1] = fmaf(gval, v.y, acc[i][1]);
          acc[i][2] = fmaf(gval, v.z, acc[i][2]);
          acc[i][3] = fmaf(gval, v.w, acc[i][3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const int q = tq + 4 * i;
    if (q >= 45) continue;
    const int tap = q / 5, c4 = q - tap * 5;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ci = c4 * 4 + e;
      if (ci < 17) atomicAdd(dw + ((size_t)co * 17 + ci) * 9 + tap, acc[i][e]);
    }
  }
}

// decoder.conv (C -> 4) weight and bias gradient: dW[o][ci][dy][dx] += coef * sum g4[n,oy,ox,o] * act[n,oy+dy-P,ox+dx-P,ci].
// Thread = (ci, phase); it walks the INPUT pixels of its phase, reads act once and the KS*KS seed pixels around it
// (the same address for all ci threads: broadcast loads), KS*KS*4 running sums in registers.
template <int KS>
__global__ void __launch_bounds__(256)
out4_wgrad_kernel(TView act, const void* __restrict__ seed, int seed_half, float* __restrict__ dw, float* __restrict__ db,
                  float coef, int N, int H, int W, int C, int px_per_block) {
  constexpr int KK = KS * KS, P = KS / 2;
  extern __shared__ float s_red[];                 // [KK*4][C]
  const int ci = threadIdx.x % C, phase = threadIdx.x / C, nphase = 256 / C;
  const int HW = H * W;
  const long long total = (long long)N * HW;
  const long long p_lo = (long long)blockIdx.x * px_per_block;
  const long long p_hi = (p_lo + px_per_block < total) ? p_lo + px_per_block : total;
  float acc[KK][4];
#pragma unroll
  for (int t = 0; t < KK; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < KK * 4 * C; i += 256) s_red[i] = 0.f;
  __syncthreads();
  if (phase < nphase) {
    for (long long ip = p_lo + phase; ip < p_hi; ip += nphase) {
      const int n = (int)(ip / HW), r = (int)(ip - (long long)n * HW);
      const int y = r / W, x = r - y * W;
      float v;
      if (act.lay == LAY_NHWC) v = __ldg(reinterpret_cast<const float*>(act.base) + ((size_t)n * HW + r) * act.pitch + ci);
      else if (act.lay == LAY_PTF) v = __ldg(reinterpret_cast<const float*>(act.base) + (((size_t)n * (C >> 2) + (ci >> 2)) * HW + r) * 4 + (ci & 3));
      else {
        const uint16_t h = __ldg(reinterpret_cast<const uint16_t*>(act.base) + (((size_t)n * (C >> 3) + (ci >> 3)) * HW + r) * 8 + (ci & 7));
        v = act.f16 ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
      }
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        const int oy = y - dy + P;
        if (oy < 0 || oy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const int ox = x - dx + P;
          if (ox < 0 || ox >= W) continue;
          const float4 g = load_seed(seed, seed_half, (size_t)n * HW + oy * W + ox);
          acc[dy * KS + dx][0] = fmaf(g.x, v, acc[dy * KS + dx][0]);
          acc[dy * KS + dx][1] = fmaf(g.y, v, acc[dy * KS + dx][1]);
          acc[dy * KS + dx][2] = fmaf(g.z, v, acc[dy * KS + dx][2]);
          acc[dy * KS + dx][3] = fmaf(g.w, v, acc[dy * KS + dx][3]);
        }
      }
      if (ci == 0) {                                  // bias: sum of the seed over all pixels
        const float4 g = load_seed(seed, seed_half, (size_t)n * HW + r);
        bs[0] += g.x; bs[1] += g.y; bs[2] += g.z; bs[3] += g.w;
      }
    }
#pragma unroll
    for (int t = 0; t < KK; ++t)
#pragma unroll
      for (int o = 0; o < 4; ++o) atomicAdd(&s_red[(t * 4 + o) * C + ci], acc[t][o]);
    if (ci == 0)
      for (int o = 0; o < 4; ++o) atomicAdd(db + o, coef * bs[o]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KK * 4 * C; i += 256) {
    const int c = i % C, o = (i / C) % 4, t = i / (4 * C);
    atomicAdd(dw + ((size_t)o * C + c) * KK + t, coef * s_red[i]);
  }
}

static int launch_out4_wgrad(Plan* p, const TView& act, const void* seed, int seed_half, float* dw, float* db, float coef,
                             cudaStream_t st) {
  const int C = p->C, KS = p->s.dec_k;
  const long long total = (long long)p->BK * p->HW;
  long long per = (total + 4LL * p->num_sms - 1) / (4LL * p->num_sms);
  if (per < 64) per = 64;
  const unsigned grid = (unsigned)((total + per - 1) / per);
  const size_t smem = (size_t)KS * KS * 4 * C * sizeof(float);
  if (KS == 3) out4_wgrad_kernel<3><<<grid, 256, smem, st>>>(act, seed, seed_half, dw, db, coef, p->BK, p->s.H, p->s.W, C, (int)per);
  else out4_wgrad_kernel<5><<<grid, 256, smem, st>>>(act, seed, seed_half, dw, db, coef, p->BK, p->s.H, p->s.W, C, (int)per);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// bias gradient of a convolution: db[c] += coef * sum over slot-images and pixels of g[.,.,c]
__global__ void __launch_bounds__(256)
chan_sum_kernel(TView g, float* __restrict__ db, float coef, int N, int HW, int px_per_block) {
  __shared__ float red[256][4];
  const int groups = (g.C + 3) / 4;                  // <= 16
  const int grp = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = 256 / groups;
  const long long total = (long long)N * HW;
  const long long p_lo = (long long)blockIdx.x * px_per_block;
  const long long p_hi = (p_lo + px_per_block < total) ? p_lo + px_per_block : total;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes)
    for (long long ip = p_lo + lane; ip < p_hi; ip += lanes) {
      const int n = (int)(ip / HW), r = (int)(ip - (long long)n * HW);
      const float4 v = load4(g, n, HW, r, grp * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  red[threadIdx.x][0] = s.x; red[threadIdx.x][1] = s.y; red[threadIdx.x][2] = s.z; red[threadIdx.x][3] = s.w;
  __syncthreads();
  if (threadIdx.x < groups * 4) {
    const int gq = threadIdx.x / 4, e = threadIdx.x % 4;
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[l * groups + gq][e];
    if (gq * 4 + e < g.C) atomicAdd(db + gq * 4 + e, coef * t);
  }
}

// the same for chunk-planar tensors: block = (pixel share, plane); a thread reads whole 16-byte plane values
// (consecutive threads = consecutive pixels), 8 (16-bit) or 4 (tf32) running sums
__global__ void __launch_bounds__(256)
plane_sum_kernel(TView g, float* __restrict__ db, float coef, int N, int HW, int px_per_block, int cmax) {
  __shared__ float red[8][8];
  const int PW = g.lay == LAY_PTF ? 4 : 8, planes = g.C / PW, k = blockIdx.y;
  const long long total = (long long)N * HW;
  const long long p_lo = (long long)blockIdx.x * px_per_block;
  const long long p_hi = (p_lo + px_per_block < total) ? p_lo + px_per_block : total;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4* base = reinterpret_cast<const uint4*>(g.base);
  for (long long ip = p_lo + threadIdx.x; ip < p_hi; ip += 256) {
    const long long n = ip / HW;
    const uint4 v = __ldg(base + (n * planes + k) * HW + (ip - n * HW));
    if (g.lay == LAY_PTF) {
      s[0] += __uint_as_float(v.x); s[1] += __uint_as_float(v.y); s[2] += __uint_as_float(v.z); s[3] += __uint_as_float(v.w);
    } else {
      const float2 a = unpack_h2(v.x, g.f16), b = unpack_h2(v.y, g.f16), c = unpack_h2(v.z, g.f16), d = unpack_h2(v.w, g.f16);
      s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = warp_sum(s[e]);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x >> 5][e] = s[e];
  __syncthreads();
  if (threadIdx.x < PW && k * PW + threadIdx.x < cmax) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(db + k * PW + threadIdx.x, coef * t);
  }
}

static int launch_chan_sum(Plan* p, const TView& g, float* db, float coef, int N, int HW, cudaStream_t st) {
  if (g.lay != LAY_NHWC) {
    const int planes = g.C / (g.lay == LAY_PTF ? 4 : 8);
    const long long total = (long long)N * HW;
    long long nb = (4LL * p->num_sms + planes - 1) / planes;
    long long per = (total + nb - 1) / nb;
    if (per < 1024) per = 1024;
    dim3 grid((unsigned)((total + per - 1) / per), planes);
    plane_sum_kernel<<<grid, 256, 0, st>>>(g, db, coef, N, HW, (int)per, g.pitch);
    IOD_LAUNCH_CHECK(p);
    return 0;
  }
  const long long total = (long long)N * HW;
  long long per = (total + 2LL * p->num_sms - 1) / (2LL * p->num_sms);
  if (per < 64) per = 64;
  chan_sum_kernel<<<(unsigned)((total + per - 1) / per), 256, 0, st>>>(g, db, coef, N, HW, (int)per);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// first decoder layer (spatial broadcast collapsed into border classes, head.cu): its input is z tiled over the
// pixels plus the two coordinate planes, so its weight gradient needs only, per slot-image and border class, the
// pixel sums of dJ/d(pre-activation 0) and their first moments in x and y.
// ------------------------------------------------------------------------------------------------
// G0/Gx/Gy[n][class][c] += sum over the pixels of the class of g, g*x, g*y.  Thread = (group of 8 channels, column);
// a block walks `rows_per_block` rows of one slot-image in maximal runs of equal row class.
__global__ void __launch_bounds__(256)
class_moments_kernel(TView g, float* __restrict__ G0, float* __restrict__ Gx, float* __restrict__ Gy, int H, int W,
                     int KS, int rows_per_block) {
  const int C = g.C, P = KS / 2, HW = H * W;
  const int n = blockIdx.z, k8 = blockIdx.y;                 // k8: group of 8 channels
  const int y_lo = blockIdx.x * rows_per_block;
  const int y_hi = (y_lo + rows_per_block < H) ? y_lo + rows_per_block : H;
  const int lane = threadIdx.x & 31;
  float s0[8], sy[8];
  for (int xb = 0; xb < W; xb += 256) {                      // (uniform trip count: the flush below shuffles)
    const int x = xb + threadIdx.x;
    const bool live = x < W;
    const int cx = live ? border_class(x, W, P) : P;
#pragma unroll
    for (int e = 0; e < 8; ++e) { s0[e] = 0.f; sy[e] = 0.f; }
    int cy_run = border_class(y_lo, H, P);
    for (int y = y_lo; y <= y_hi; ++y) {
      const int cy = (y < y_hi) ? border_class(y, H, P) : -1;
      if (cy != cy_run) {                                    // flush the run (block-uniform branch)
        float* o0 = G0 + ((size_t)n * KS * KS + cy_run * KS) * C + k8 * 8;
        float* ox = Gx + ((size_t)n * KS * KS + cy_run * KS) * C + k8 * 8;
        float* oy = Gy + ((size_t)n * KS * KS + cy_run * KS) * C + k8 * 8;
        const bool border = live && cx != P;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (border && k8 * 8 + e < C) {                    // a border column keeps its own sums
            atomicAdd(o0 + (size_t)cx * C + e, s0[e]);
            atomicAdd(ox + (size_t)cx * C + e, s0[e] * (float)x);
            atomicAdd(oy + (size_t)cx * C + e, sy[e]);
          }
          // interior columns: one atomic per warp and value
          const float a0 = warp_sum(border || !live ? 0.f : s0[e]);
          const float ax = warp_sum(border || !live ? 0.f : s0[e] * (float)x);
          const float ay = warp_sum(border || !live ? 0.f : sy[e]);
          if (lane == 0 && k8 * 8 + e < C) {
            atomicAdd(o0 + (size_t)P * C + e, a0);
            atomicAdd(ox + (size_t)P * C + e, ax);
            atomicAdd(oy + (size_t)P * C + e, ay);
          }
          s0[e] = 0.f; sy[e] = 0.f;
        }
        cy_run = cy;
      }
      if (y == y_hi) break;
      if (live) {
        const float4 a = load4(g, n, HW, y * W + x, k8 * 8), b = load4(g, n, HW, y * W + x, k8 * 8 + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) { s0[e] += v[e]; sy[e] = fmaf(v[e], (float)y, sy[e]); }
      }
    }
  }
}

// dW0[co][ci][dy][dx] (ci < L: z channels; ci = L, L+1: x / y coordinate planes) and db0[co].
// block = (tap, co); threads = input channels.
__global__ void __launch_bounds__(256)
wgrad_l0_kernel(const float* __restrict__ G0, const float* __restrict__ Gx, const float* __restrict__ Gy,
                const float* __restrict__ z, float* __restrict__ dw, float* __restrict__ db, float coef, int BK, int C,
                int L, int KS, int H, int W) {
  extern __shared__ float sS[];                      // [chunk] per slot-image: class sums valid for this tap
  const int tap = blockIdx.x, co = blockIdx.y, dy = tap / KS, dx = tap - dy * KS, P = KS / 2, NC = KS * KS;
  // classes (cy, cx) whose pixels keep tap (dy, dx) inside the image
  auto valid = [&](int cy, int cx) {
    const int dy_lo = (cy < P) ? P - cy : 0, dy_hi = (cy > P) ? KS - 1 - (cy - P) : KS - 1;
    const int dx_lo = (cx < P) ? P - cx : 0, dx_hi = (cx > P) ? KS - 1 - (cx - P) : KS - 1;
    return dy >= dy_lo && dy <= dy_hi && dx >= dx_lo && dx <= dx_hi;
  };
  float accz = 0.f;                                  // thread ci < L
  float t0 = 0.f, tx = 0.f, ty = 0.f, tall = 0.f;    // totals over n (threads stride over n)
  constexpr int CH = 1024;
  for (int n0 = 0; n0 < BK; n0 += CH) {
    const int nn = (BK - n0 < CH) ? BK - n0 : CH;
    __syncthreads();
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
      const size_t o = (size_t)(n0 + i) * NC * C + co;
      float s = 0.f, sx = 0.f, sy = 0.f, sa = 0.f;
      for (int cls = 0; cls < NC; ++cls) {
        const float g0 = G0[o + (size_t)cls * C];
        sa += g0;
        if (valid(cls / KS, cls % KS)) { s += g0; sx += Gx[o + (size_t)cls * C]; sy += Gy[o + (size_t)cls * C]; }
      }
      sS[i] = s;
      t0 += s; tx += sx; ty += sy; tall += sa;
    }
    __syncthreads();
    for (int ci = threadIdx.x; ci < L; ci += blockDim.x) {   // (L <= 256: one pass)
      float a = 0.f;
      for (int i = 0; i < nn; ++i) a = fmaf(z[(size_t)(n0 + i) * L + ci], sS[i], a);
      accz += a;
    }
  }
  if (threadIdx.x < L) dw[(((size_t)co * (L + 2) + threadIdx.x) * KS + dy) * KS + dx] += coef * accz;
  // coordinate channels and bias: block totals
  t0 = warp_sum(t0); tx = warp_sum(tx); ty = warp_sum(ty); tall = warp_sum(tall);
  __shared__ float tot[4][8];
  if ((threadIdx.x & 31) == 0) {
    tot[0][threadIdx.x >> 5] = t0; tot[1][threadIdx.x >> 5] = tx; tot[2][threadIdx.x >> 5] = ty; tot[3][threadIdx.x >> 5] = tall;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a0 = 0.f, ax = 0.f, ay = 0.f, aa = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a0 += tot[0][w]; ax += tot[1][w]; ay += tot[2][w]; aa += tot[3][w]; }
    // coordinate of the INPUT pixel (x + dx - P): -1 + 2 (x + dx - P) / (W - 1)   (iodine.py:526-530)
    const float bx = (W > 1) ? 2.f / (float)(W - 1) : 0.f, by = (H > 1) ? 2.f / (float)(H - 1) : 0.f;
    const float cx0 = -1.f + bx * (float)(dx - P), cy0 = -1.f + by * (float)(dy - P);
    dw[(((size_t)co * (L + 2) + L) * KS + dy) * KS + dx] += coef * (cx0 * a0 + bx * ax);
    dw[(((size_t)co * (L + 2) + L + 1) * KS + dy) * KS + dx] += coef * (cy0 * a0 + by * ay);
    if (tap == 0) db[co] += coef * aa;
  }
}

// ------------------------------------------------------------------------------------------------
// dense helpers of the refiner's backward pass
// ------------------------------------------------------------------------------------------------
// C[m][n] = alpha * sum_k A(m,k) B(k,n) + beta * C[m][n],  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
             long long sbn, float* __restrict__ Cm, int ldc, int Mm, int Nn, int Kk, float alpha, float beta) {
  __shared__ float sa[16][64 + 4];
  __shared__ float sb[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < Kk; k0 += 16) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      // choose the faster-running index by the smaller stride so that loads coalesce either way
      int kk, r;
      if (sak <= sam) { kk = i % 16; r = i / 16; } else { r = i % 64; kk = i / 64; }
      sa[kk][r] = (m0 + r < Mm && k0 + kk < Kk) ? A[(long long)(m0 + r) * sam + (long long)(k0 + kk) * sak] : 0.f;
      int kb, c;
      if (sbk <= sbn) { kb = i % 16; c = i / 16; } else { c = i % 64; kb = i / 64; }
      sb[kb][c] = (n0 + c < Nn && k0 + kb < Kk) ? B[(long long)(k0 + kb) * sbk + (long long)(n0 + c) * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sb[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= Mm) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Nn) continue;
      float* c = Cm + (size_t)m * ldc + n;
      *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
    }
  }
}
static int gemm(Plan* p, const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* Cm,
                int ldc, int Mm, int Nn, int Kk, float alpha, float beta, cudaStream_t st) {
  dim3 grid((Nn + 63) / 64, (Mm + 63) / 64);
  sgemm_kernel<<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, Cm, ldc, Mm, Nn, Kk, alpha, beta);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// out[c] += alpha * sum_r X[r*ld + c]     (bias gradients of the dense layers; one block per 32 columns)
__global__ void __launch_bounds__(256)
col_sum_kernel(const float* __restrict__ X, int ld, int rows, int cols, float* __restrict__ out, float* __restrict__ out2,
               float alpha) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int r = w; r < rows; r += 8) s += X[(size_t)r * ld + c];
  red[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < cols) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    out[c] += alpha * t;
    if (out2) out2[c] += alpha * t;
  }
}
static int col_sum(Plan* p, const float* X, int ld, int rows, int cols, float* out, float* out2, float alpha, cudaStream_t st) {
  col_sum_kernel<<<(cols + 31) / 32, 256, 0, st>>>(X, ld, rows, cols, out, out2, alpha);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// LSTMCell backward, pointwise part (gate order i, f, g, o; iodine.py:488): dc_in already holds the heads' share
__global__ void lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c0, const float* __restrict__ c1,
                                const float* __restrict__ dh, const float* __restrict__ dc_in, float* __restrict__ dgates,
                                float* __restrict__ dc_prev, int N, int M) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * M; i += gridDim.x * blockDim.x) {
    const int n = i / M, j = i - n * M;
    const float* g = gates + (size_t)n * 4 * M;
    const float si = sigmoid_f(g[j]), sf = sigmoid_f(g[M + j]), tg = tanhf(g[2 * M + j]), so = sigmoid_f(g[3 * M + j]);
    const float tc = tanhf(c1[i]);
    const float dhv = dh[i];
    const float dc = dc_in[i] + dhv * so * (1.f - tc * tc);
    float* d = dgates + (size_t)n * 4 * M;
    d[j] = dc * tg * si * (1.f - si);
    d[M + j] = dc * c0[i] * sf * (1.f - sf);
    d[2 * M + j] = dc * si * (1.f - tg * tg);
    d[3 * M + j] = dhv * tc * so * (1.f - so);
    dc_prev[i] = dc * sf;
  }
}

// the two ELUs between the MLP's Linear and the LSTM input (iodine.py:485, 565): u = ELU(lin) is on the tape,
// d(lin) = d * ELU'(u) * ELU'(lin), ELU'(u) = u > 0 ? 1 : exp(u), ELU'(lin) = u > 0 ? 1 : u + 1
__global__ void mlp_bwd_kernel(float* __restrict__ d, const float* __restrict__ u, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float uv = u[i];
    if (uv <= 0.f) d[i] *= expf(uv) * (uv + 1.f);
  }
}

// adaptive_avg_pool2d backward fused with the ELU' of the last refine conv: g[n][pix][c] = dpool[n][c] / HWo * ELU'(act)
__global__ void pool_bwd_kernel(const float* __restrict__ dpool, const float* __restrict__ act, float* __restrict__ g,
                                int N, int HWo, int C) {
  const size_t total = (size_t)N * HWo * C;
  const float inv = 1.f / (float)HWo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t n = i / ((size_t)HWo * C);
    g[i] = dpool[n * C + c] * inv * elu_grad_from_act(act[i]);
  }
}

// data-gradient of a strided refine conv (conv_transpose2d, output_padding implied by the input size) times ELU' of
// the activation it feeds:  gin[n,iy,ix,ci] = ELU'(actp) * sum_{dy,dx,co} g[n,(iy+P-dy)/S,(ix+P-dx)/S,co] W[tap][ci][co]
// over the taps whose source position is integral and inside.  W in the forward pack [tap][ci][co] (co contiguous).
__global__ void __launch_bounds__(256)
refine_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ wp, const float* __restrict__ actp,
                    float* __restrict__ gin, int N, int Hin, int Win, int Hout, int Wout, int C, int KS, int S) {
  const int P = KS / 2;
  const size_t total = (size_t)N * Hin * Win * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C);
    size_t r = i / C;
    const int ix = (int)(r % Win); r /= Win;
    const int iy = (int)(r % Hin);
    const int n = (int)(r / Hin);
    float acc = 0.f;
    for (int dy = 0; dy < KS; ++dy) {
      const int ty = iy + P - dy;
      if (ty < 0 || ty % S) continue;
      const int oy = ty / S;
      if (oy >= Hout) continue;
      for (int dx = 0; dx < KS; ++dx) {
        const int tx = ix + P - dx;
        if (tx < 0 || tx % S) continue;
        const int ox = tx / S;
        if (ox >= Wout) continue;
        const float4* gp = reinterpret_cast<const float4*>(g + (((size_t)n * Hout + oy) * Wout + ox) * C);
        const float4* wr = reinterpret_cast<const float4*>(wp + ((size_t)(dy * KS + dx) * C + ci) * C);
        float a0 = 0.f, a1 = 0.f;
        for (int q = 0; q < C / 4; q += 2) {
          const float4 g0 = __ldg(gp + q), w0 = __ldg(wr + q), g1 = __ldg(gp + q + 1), w1 = __ldg(wr + q + 1);
          a0 = fmaf(g0.x, w0.x, a0); a0 = fmaf(g0.y, w0.y, a0); a0 = fmaf(g0.z, w0.z, a0); a0 = fmaf(g0.w, w0.w, a0);
          a1 = fmaf(g1.x, w1.x, a1); a1 = fmaf(g1.y, w1.y, a1); a1 = fmaf(g1.z, w1.z, a1); a1 = fmaf(g1.w, w1.w, a1);
        }
        acc += a0 + a1;
      }
    }
    gin[i] = acc * elu_grad_from_act(actp[i]);
  }
}

// The same as an implicit GEMM (the gather kernel above runs at ~5 TFLOP/s).  Input pixels are split into the S*S
// parity classes of (iy, ix): within a class every pixel receives the same taps, so a class is a dense GEMM
// [pixels of the class] x [taps * co] x [ci].  Block = 64 pixels x 64 ci, 4 x 4 register tiles, 16 co per round.
__global__ void __launch_bounds__(256)
refine_dgrad_gemm_kernel(const float* __restrict__ g, const float* __restrict__ wp, const float* __restrict__ actp,
                         float* __restrict__ gin, int N, int Hin, int Win, int Hout, int Wout, int C, int KS, int S) {
  __shared__ __align__(16) float sa[16][64 + 4];
  __shared__ __align__(16) float sw[16][64 + 4];
  const int P = KS / 2;
  const int cls = blockIdx.y, py = cls / S, px = cls - py * S;
  const int Hc = (Hin - py + S - 1) / S, Wc = (Win - px + S - 1) / S;
  const long long total = (long long)N * Hc * Wc;
  const long long m0 = (long long)blockIdx.x * 64;
  if (m0 >= total) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // loader role: pixel lm of the tile, 4 consecutive co of the 16-wide round
  const int lm = threadIdx.x >> 2, lsub = (threadIdx.x & 3) * 4;
  const long long lpix = m0 + lm;
  int ln = 0, liy = 0, lix = 0;
  const bool lvalid = lpix < total;
  if (lvalid) {
    ln = (int)(lpix / ((long long)Hc * Wc));
    const int r = (int)(lpix - (long long)ln * Hc * Wc);
    liy = (r / Wc) * S + py;
    lix = (r % Wc) * S + px;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int dy = 0; dy < KS; ++dy) {
    if ((py + P - dy) % S != 0) continue;              // (py + P - dy may be negative: C++ % keeps the sign, 0 stays 0)
    for (int dx = 0; dx < KS; ++dx) {
      if ((px + P - dx) % S != 0) continue;
      const int ty_ = liy + P - dy, tx_ = lix + P - dx;
      const int oy = ty_ / S, ox = tx_ / S;
      const bool inside = lvalid && ty_ >= 0 && tx_ >= 0 && oy < Hout && ox < Wout;
      const float* grow = g + (((size_t)ln * Hout + (inside ? oy : 0)) * Wout + (inside ? ox : 0)) * C;
      const float* wrow = wp + ((size_t)(dy * KS + dx) * C + lm) * C;      // W[tap][ci = lm][co]
      for (int c0 = 0; c0 < C; c0 += 16) {
        const float4 av = inside ? __ldg(reinterpret_cast<const float4*>(grow + c0 + lsub)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 wv = (lm < C) ? __ldg(reinterpret_cast<const float4*>(wrow + c0 + lsub)) : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        sa[lsub + 0][lm] = av.x; sa[lsub + 1][lm] = av.y; sa[lsub + 2][lm] = av.z; sa[lsub + 3][lm] = av.w;
        sw[lsub + 0][lm] = wv.x; sw[lsub + 1][lm] = wv.y; sw[lsub + 2][lm] = wv.z; sw[lsub + 3][lm] = wv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&sw[kk][tx * 4]);
          const float av4[4] = {a.x, a.y, a.z, a.w}, bv4[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av4[i], bv4[j], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long pix = m0 + ty * 4 + i;
    if (pix >= total) continue;
    const int n = (int)(pix / ((long long)Hc * Wc));
    const int r = (int)(pix - (long long)n * Hc * Wc);
    const int iy = (r / Wc) * S + py, ix = (r % Wc) * S + px;
    const size_t o = (((size_t)n * Hin + iy) * Win + ix) * C + tx * 4;
    if (tx * 4 < C) {
      const float4 a = *reinterpret_cast<const float4*>(actp + o);
      float4 out;
      out.x = acc[i][0] * elu_grad_from_act(a.x); out.y = acc[i][1] * elu_grad_from_act(a.y);
      out.z = acc[i][2] * elu_grad_from_act(a.z); out.w = acc[i][3] * elu_grad_from_act(a.w);
      *reinterpret_cast<float4*>(gin + o) = out;
    }
  }
}

// loss = -sum_i (i+1)/(T+1) * (ll_i - kl_i) / Bg   (iodine.py:149-158), from the [T+1][2] table of batch sums
__global__ void loss_kernel(const float* __restrict__ terms, float* __restrict__ loss, int T, float inv_bg) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i <= T; ++i) s += (float)(i + 1) / (float)(T + 1) * (terms[2 * i] - terms[2 * i + 1]) * inv_bg;
    *loss = -s;
  }
}
__global__ void terms2_kernel(const double* __restrict__ accum, float* __restrict__ out) {
  if (threadIdx.x < 2) out[threadIdx.x] = (float)accum[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------------
static TView dec_view(const Plan* p, const void* buf) {
  TView v;
  v.base = buf; v.C = p->C; v.pitch = p->C;
  v.lay = !tc_mode(p) ? LAY_NHWC : tf_mode(p) ? LAY_PTF : LAY_P16;
  v.f16 = p->s.precision == IODINE_FP16;
  return v;
}
static TView nhwc_view(const float* buf, int C, int pitch) {
  TView v;
  v.base = buf; v.C = C; v.pitch = pitch; v.lay = LAY_NHWC; v.f16 = 0;
  return v;
}

// decoder data-gradient chain of one ELBO evaluation with the weight gradients of every layer interleaved
static int decoder_backward_train(Plan* p, TrainState* ts, float coef, cudaStream_t st) {
  const IodineShape& s = p->s;
  const int n = s.dec_layers, C = p->C, KS = s.dec_k;
  const size_t gsz = (size_t)p->BK * p->n_class * C * sizeof(float);
  IOD_CHECK_CUDA(cudaMemsetAsync(p->G, 0, gsz, st));
  IOD_CHECK_CUDA(cudaMemsetAsync(ts->Gx, 0, gsz, st));
  IOD_CHECK_CUDA(cudaMemsetAsync(ts->Gy, 0, gsz, st));
  const int seed_half = (tc_mode(p) && !tf_mode(p)) ? (s.precision == IODINE_FP16 ? 2 : 1) : 0;
  // decoder.conv (C -> 4): weight + bias gradient from the seeds, then its data-gradient
  const bool wtc = wgrad_tc_supported(p) != 0, tf = tf_mode(p);
  if (wtc) {
    const void *a16 = p->act[n - 1], *s16 = p->seed4;
    if (tf) {                                        // fp16 copies of the operands (wgrad_tc.cu)
      if (wgrad_to_h16(p, p->act[n - 1], ts->h16_act, C, st) || wgrad_to_h16(p, p->seed4, ts->h16_g, 4, st)) return 1;
      a16 = ts->h16_act; s16 = ts->h16_g;
    }
    if (launch_wgrad_tc_out4(p, a16, s16, ts->gflat + ts->o_out_w, coef, st)) return 1;
    TView sv;                                        // the seed: one 16-bit plane of 8 channels, 4 of them real
    sv.base = s16; sv.lay = LAY_P16; sv.f16 = s.precision != IODINE_BF16; sv.C = 8; sv.pitch = 4;
    if (launch_chan_sum(p, sv, ts->gflat + ts->o_out_b, coef, p->BK, p->HW, st)) return 1;
  } else if (launch_out4_wgrad(p, dec_view(p, p->act[n - 1]), p->seed4, seed_half, ts->gflat + ts->o_out_w,
                               ts->gflat + ts->o_out_b, coef, st))
    return 1;
  if (tc_mode(p)) {
    if (tc_launch_dgrad_in4(p, p->seed4, p->act[n - 1], p->gbuf[0], st)) return 1;
  } else {
    if (launch_dgrad_in4(p, p->seed4, (const float*)p->act[n - 1], (float*)p->gbuf[0], st, true)) return 1;
  }
  int cur = 0;
  for (int l = n - 1; l >= 1; --l) {
    // gbuf[cur] = dJ/d(pre-activation l)
    const TView gv = dec_view(p, p->gbuf[cur]);
    if (wtc) {
      const void *a16 = p->act[l - 1], *g16 = p->gbuf[cur];
      if (tf) {
        if (wgrad_to_h16(p, p->act[l - 1], ts->h16_act, C, st) || wgrad_to_h16(p, p->gbuf[cur], ts->h16_g, C, st)) return 1;
        a16 = ts->h16_act; g16 = ts->h16_g;
      }
      if (launch_wgrad_tc(p, a16, g16, ts->gflat + ts->o_dec_w[l], coef, st)) return 1;
    } else if (launch_conv_wgrad(p, dec_view(p, p->act[l - 1]), gv, ts->gflat + ts->o_dec_w[l], coef, p->BK, s.H, s.W, s.H, s.W, C,
                                 C, KS, 1, st))
      return 1;
    if (launch_chan_sum(p, gv, ts->gflat + ts->o_dec_b[l], coef, p->BK, p->HW, st)) return 1;
    if (tc_mode(p)) {
      if (tc_launch_conv(p, l, true, p->gbuf[cur], p->act[l - 1], p->gbuf[cur ^ 1], nullptr, st)) return 1;
    } else {
      if (launch_conv_cc(p, (const float*)p->gbuf[cur], p->dec[l].wt, nullptr, (const float*)p->act[l - 1],
                         (float*)p->gbuf[cur ^ 1], nullptr, 1, st))
        return 1;
    }
    cur ^= 1;
  }
  // gbuf[cur] = dJ/d(pre-activation 0): class sums (what the inference path fuses into the last data-gradient) plus
  // the first moments the coordinate channels need
  {
    const int rpb = 16;
    dim3 grid((s.H + rpb - 1) / rpb, (C + 7) / 8, p->BK);
    class_moments_kernel<<<grid, 256, 0, st>>>(dec_view(p, p->gbuf[cur]), p->G, ts->Gx, ts->Gy, s.H, s.W, KS, rpb);
    IOD_LAUNCH_CHECK(p);
  }
  {
    dim3 grid(KS * KS, C);
    const int chunk = p->BK < 1024 ? p->BK : 1024;
    wgrad_l0_kernel<<<grid, 256, chunk * sizeof(float), st>>>(p->G, ts->Gx, ts->Gy, p->z, ts->gflat + ts->o_dec_w[0],
                                                              ts->gflat + ts->o_dec_b[0], coef, p->BK, C, s.L, KS, s.H, s.W);
    IOD_LAUNCH_CHECK(p);
  }
  return 0;
}

// backward of refiner call t (oracle/train_restatement.py: refine_backward)
static int refiner_backward(Plan* p, TrainState* ts, int t, float alpha, int cur, cudaStream_t st) {
  const IodineShape& s = p->s;
  const int N = p->BK, M = p->M, L = s.L, Cr = p->Cr, I = M + 4 * L;
  const size_t nm = (size_t)N * M;
  const float* g = ts->pg + (size_t)(t + 1) * N * 2 * L;          // dJ_{t+1}/d(posterior_{t+1}), [N][2L]
  const float* c1 = ts->cs + (size_t)(t + 1) * nm;
  const float* c0 = ts->cs + (size_t)t * nm;
  const float* h0 = ts->hs + (size_t)t * nm;
  float* gf = ts->gflat;
  // heads read the CELL state (iodine.py:488-492): dW = alpha g^T c1, db = alpha colsum(g), dc += alpha g [Wm; Wl]
  if (gemm(p, g, 1, 2 * L, c1, M, 1, gf + ts->o_head_w, M, 2 * L, M, N, alpha, 1.f, st)) return 1;
  if (col_sum(p, g, 2 * L, N, 2 * L, gf + ts->o_head_b, nullptr, alpha, st)) return 1;
  if (gemm(p, g, 2 * L, 1, p->head_w, M, 1, ts->dc[cur], M, N, M, 2 * L, alpha, 1.f, st)) return 1;
  lstm_bwd_kernel<<<(N * M + 255) / 256, 256, 0, st>>>(ts->gates[t], c0, c1, ts->dh[cur], ts->dc[cur], ts->dgates,
                                                         ts->dc[cur ^ 1], N, M);
  IOD_LAUNCH_CHECK(p);
  // LSTM weights: dW_ih = dgates^T xin, dW_hh = dgates^T h0, both biases = colsum(dgates)
  if (gemm(p, ts->dgates, 1, 4 * M, ts->xin[t], I, 1, gf + ts->o_wih, I, 4 * M, I, N, 1.f, 1.f, st)) return 1;
  if (gemm(p, ts->dgates, 1, 4 * M, h0, M, 1, gf + ts->o_whh, M, 4 * M, M, N, 1.f, 1.f, st)) return 1;
  if (col_sum(p, ts->dgates, 4 * M, N, 4 * M, gf + ts->o_bih, gf + ts->o_bhh, 1.f, st)) return 1;
  // state and input gradients: dh_prev = dgates W_hh ; d(xin[:, :M]) = dgates W_ih[:, :M] (the latent half is detached, 343)
  if (gemm(p, ts->dgates, 4 * M, 1, p->w_hh, M, 1, ts->dh[cur ^ 1], M, N, M, 4 * M, 1.f, 0.f, st)) return 1;
  if (gemm(p, ts->dgates, 4 * M, 1, p->w_ih, I, 1, ts->dx, M, N, M, 4 * M, 1.f, 0.f, st)) return 1;
  mlp_bwd_kernel<<<(N * M + 255) / 256, 256, 0, st>>>(ts->dx, ts->u[t], N * M);
  IOD_LAUNCH_CHECK(p);
  if (gemm(p, ts->dx, 1, M, ts->pool[t], Cr, 1, gf + ts->o_mlp_w, Cr, M, Cr, N, 1.f, 1.f, st)) return 1;
  if (col_sum(p, ts->dx, M, N, M, gf + ts->o_mlp_b, nullptr, 1.f, st)) return 1;
  if (gemm(p, ts->dx, M, 1, p->mlp_w, Cr, 1, ts->dpool, Cr, N, Cr, M, 1.f, 0.f, st)) return 1;
  // pool + conv stack
  const int nl = s.ref_layers;
  const int HWo = p->ref_h[nl] * p->ref_w[nl];
  {
    const size_t total = (size_t)N * HWo * Cr;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pool_bwd_kernel<<<blocks, 256, 0, st>>>(ts->dpool, ts->ract[t][nl - 1], ts->rg[0], N, HWo, Cr);
    IOD_LAUNCH_CHECK(p);
  }
  int rc = 0;
  for (int l = nl - 1; l >= 0; --l) {
    const int Hin = p->ref_h[l], Win = p->ref_w[l], Hout = p->ref_h[l + 1], Wout = p->ref_w[l + 1];
    const TView gv = nhwc_view(ts->rg[rc], Cr, Cr);
    const TView iv = (l == 0) ? nhwc_view(ts->enc[t], 17, 20) : nhwc_view(ts->ract[t][l - 1], Cr, Cr);
    if (l == 0 && Cr == 64 && s.ref_k == 3 && s.ref_stride == 2 && Wout % 16 == 0 && !getenv("IODINE_REFINE_WGRAD_GENERIC")) {
      const long long rounds = (long long)N * Hout * (Wout / 16);
      long long per = (rounds + 2LL * p->num_sms - 1) / (2LL * p->num_sms);
      if (per < 8) per = 8;
      refine_wgrad_l0_kernel<<<(unsigned)((rounds + per - 1) / per), 256, 0, st>>>(ts->enc[t], ts->rg[rc], gf + ts->o_ref_w[0], N,
                                                                                    Hin, Win, Hout, Wout, (int)per);
      IOD_LAUNCH_CHECK(p);
    } else if (launch_conv_wgrad(p, iv, gv, gf + ts->o_ref_w[l], 1.f, N, Hin, Win, Hout, Wout, l == 0 ? 17 : Cr, Cr, s.ref_k,
                                 s.ref_stride, st))
      return 1;
    if (launch_chan_sum(p, gv, gf + ts->o_ref_b[l], 1.f, N, Hout * Wout, st)) return 1;
    if (l > 0) {
      static const bool gather = getenv("IODINE_REFINE_DGRAD_GATHER") != nullptr;
      if (gather) {
        const size_t total = (size_t)N * Hin * Win * Cr;
        const unsigned blocks = (unsigned)((total + 255) / 256 < 65535 * 8 ? (total + 255) / 256 : 65535 * 8);
        refine_dgrad_kernel<<<blocks, 256, 0, st>>>(ts->rg[rc], p->ref_wp[l], ts->ract[t][l - 1], ts->rg[rc ^ 1], N, Hin, Win,
                                                     Hout, Wout, Cr, s.ref_k, s.ref_stride);
      } else {
        const int S = s.ref_stride;
        const long long per_class = (long long)N * ((Hin + S - 1) / S) * ((Win + S - 1) / S);
        dim3 grid((unsigned)((per_class + 63) / 64), S * S);
        refine_dgrad_gemm_kernel<<<grid, 256, 0, st>>>(ts->rg[rc], p->ref_wp[l], ts->ract[t][l - 1], ts->rg[rc ^ 1], N, Hin, Win,
                                                        Hout, Wout, Cr, s.ref_k, S);
      }
      IOD_LAUNCH_CHECK(p);
      rc ^= 1;
    }
  }
  return 0;
}

}  // namespace iod

using namespace iod;

extern "C" {

IODINE_API int iodine_plan_train_workspace_bytes(IodinePlan* plan, size_t* bytes_out) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p && bytes_out, "iodine_plan_train_workspace_bytes: null argument");
  IOD_REQUIRE(p->ks_ranks == 1, "training is not available on K-split plans");
  if (!p->train) p->train = new TrainState();
  TrainState* ts = train_state(p);
  ts->need = train_carve(p, ts, nullptr);
  *bytes_out = ts->need;
  return 0;
}

IODINE_API int iodine_plan_set_train_workspace(IodinePlan* plan, void* workspace, size_t bytes) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p && workspace, "iodine_plan_set_train_workspace: null argument");
  size_t need = 0;
  if (iodine_plan_train_workspace_bytes(plan, &need)) return 1;
  IOD_REQUIRE(bytes >= need, "training workspace too small: %zu < %zu", bytes, need);
  IOD_REQUIRE(((uintptr_t)workspace & 1023) == 0, "training workspace must be 1024-byte aligned");
  TrainState* ts = train_state(p);
  ts->ws = workspace;
  train_carve(p, ts, (char*)workspace);
  return 0;
}

IODINE_API int iodine_train_step(IodinePlan* plan, const float* x, const float* eps, int32_t global_batch,
                                 const IodineGrads* grads_out, float* loss_out, float* elbo_terms_out, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (plan_check_ready(p)) return 1;
  TrainState* ts = train_state(p);
  IOD_REQUIRE(ts && ts->ws, "training workspace not set (iodine_plan_set_train_workspace)");
  IOD_REQUIRE(x && eps && grads_out, "null tensor argument");
  cudaStream_t st = (cudaStream_t)stream;
  const IodineShape& s = p->s;
  const int T = s.T, N = p->BK, M = p->M, L = s.L;
  const size_t nl = (size_t)N * L, nm = (size_t)N * M;
  const float Bg = (float)(global_batch > 0 ? global_batch : s.B);
  float* gf = ts->gflat;
  IOD_CHECK_CUDA(cudaMemsetAsync(gf, 0, ts->n_flat * sizeof(float), st));
  if (launch_init_state(p, ts->mu, ts->lv, ts->hs, ts->cs, st)) return 1;
  // the forward sweep runs the refiner on the fp32 kernels into the tape (plan-field redirection, restored below)
  float *sv_enc20 = p->enc20, *sv_xin = p->xin, *sv_pool = p->pool;
  int rc = 0;
  for (int i = 0; i <= T && !rc; ++i) {
    const float coef = -((float)(i + 1) / (float)(T + 1)) / Bg;
    const float* eps_i = eps + (size_t)i * nl;
    if (i < T) { p->xin = ts->xin[i]; p->pool = ts->pool[i]; p->enc20 = ts->enc[i]; }
    else { p->xin = sv_xin; p->pool = sv_pool; p->enc20 = sv_enc20; }
    rc = plan_decoder_forward(p, ts->mu, ts->lv, eps_i, nullptr, st);
    if (!rc) rc = launch_mixture(p, x, true, st);
    if (!rc && i == T) rc = launch_recombine(p, p->log_pred, p->log_mask, p->log_mean, 1, st);   // logger side channel: the last elbo()
    if (!rc) rc = decoder_backward_train(p, ts, coef, st);
    float* pg_i = ts->pg + (size_t)i * N * 2 * L;
    if (!rc) rc = launch_post_grads(p, ts->mu, ts->lv, eps_i, pg_i, st);
    if (rc) break;
    terms2_kernel<<<1, 32, 0, st>>>(p->accum, gf + ts->o_terms + 2 * i);
    p->launches++;
    if (i == 0) rc = col_sum(p, pg_i, 2 * L, N, 2 * L, gf + ts->o_init, nullptr, coef, st);   // init_unit tiles over (B, K)
    if (i == T || rc) break;
    rc = launch_assemble(p, x, st);
    if (!rc) rc = launch_refine_convs(p, ts->enc[i], st, ts->ract[i].data());
    if (rc) break;
    if (cudaMemcpyAsync(ts->hs + (size_t)(i + 1) * nm, ts->hs + (size_t)i * nm, nm * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(ts->cs + (size_t)(i + 1) * nm, ts->cs + (size_t)i * nm, nm * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("train step: state copy failed");
      rc = 1;
      break;
    }
    p->tape_u = ts->u[i]; p->tape_gates = ts->gates[i];
    rc = launch_head(p, ts->mu, ts->lv, ts->hs + (size_t)(i + 1) * nm, ts->cs + (size_t)(i + 1) * nm, st);
    p->tape_u = nullptr; p->tape_gates = nullptr;
  }
  p->enc20 = sv_enc20; p->xin = sv_xin; p->pool = sv_pool;
  p->tape_u = nullptr; p->tape_gates = nullptr;
  if (rc) return 1;
  // backward sweep over the refiner calls
  IOD_CHECK_CUDA(cudaMemsetAsync(ts->dh[0], 0, nm * sizeof(float), st));
  IOD_CHECK_CUDA(cudaMemsetAsync(ts->dc[0], 0, nm * sizeof(float), st));
  int cur = 0;
  for (int t = T - 1; t >= 0; --t) {
    const float alpha = -((float)(t + 2) / (float)(T + 1)) / Bg;                  // c_{t+1}
    if (refiner_backward(p, ts, t, alpha, cur, st)) return 1;
    cur ^= 1;
  }
  loss_kernel<<<1, 32, 0, st>>>(gf + ts->o_terms, gf + ts->o_loss, T, 1.f / Bg);
  IOD_LAUNCH_CHECK(p);
  // slot-shard data parallelism: this rank's share of every gradient, of the loss and of the ELBO table -> the sums
  if (plan_allreduce_sum(p, gf, ts->n_flat, st)) return 1;
  // hand the gradients out in the state_dict's shapes
  const size_t C = p->C, Cr = p->Cr, kk = (size_t)s.dec_k * s.dec_k, rkk = (size_t)s.ref_k * s.ref_k;
  auto give = [&](float* dst, size_t off, size_t n) -> int {
    if (!dst) return 0;
    IOD_CHECK_CUDA(cudaMemcpyAsync(dst, gf + off, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  for (int l = 0; l < s.dec_layers; ++l)
    if (give(grads_out->dec_w[l], ts->o_dec_w[l], C * (l == 0 ? L + 2 : C) * kk) || give(grads_out->dec_b[l], ts->o_dec_b[l], C)) return 1;
  if (give(grads_out->dec_out_w, ts->o_out_w, 4 * C * kk) || give(grads_out->dec_out_b, ts->o_out_b, 4)) return 1;
  for (int l = 0; l < s.ref_layers; ++l)
    if (give(grads_out->ref_w[l], ts->o_ref_w[l], Cr * (l == 0 ? 17 : Cr) * rkk) || give(grads_out->ref_b[l], ts->o_ref_b[l], Cr)) return 1;
  if (give(grads_out->mlp_w, ts->o_mlp_w, (size_t)M * Cr) || give(grads_out->mlp_b, ts->o_mlp_b, M)) return 1;
  if (give(grads_out->lstm_w_ih, ts->o_wih, (size_t)4 * M * (M + 4 * L)) || give(grads_out->lstm_w_hh, ts->o_whh, (size_t)4 * M * M)) return 1;
  if (give(grads_out->lstm_b_ih, ts->o_bih, 4 * M) || give(grads_out->lstm_b_hh, ts->o_bhh, 4 * M)) return 1;
  if (give(grads_out->mean_w, ts->o_head_w, (size_t)L * M) || give(grads_out->logvar_w, ts->o_head_w + (size_t)L * M, (size_t)L * M)) return 1;
  if (give(grads_out->mean_b, ts->o_head_b, L) || give(grads_out->logvar_b, ts->o_head_b + L, L)) return 1;
  if (give(grads_out->init_mean, ts->o_init, L) || give(grads_out->init_logvar, ts->o_init + L, L)) return 1;
  if (give(loss_out, ts->o_loss, 1) || give(elbo_terms_out, ts->o_terms, 2 * (size_t)(T + 1))) return 1;
  return 0;
}

}  // extern "C"
