// head.cu -- everything on the path that is not a big convolution or the pixel mixture:
//   * weight repacking (once per load_state_dict)
//   * Gaussian.sample (reference lib/modeling/iodine.py:620-634) fused with the
//     broadcast-collapsed first decoder layer (SpatialBroadcast 512-540 + MultiLayerConv
//     layer 0, 583/592): the [BK, L+2, H, W] broadcast tensor is never materialised
//   * the posterior gradients / KL / latent vector (iodine.py:253-275, 653-659, 382-384)
//   * the refinement head: MLP + double ELU (485, 565), LSTMCell (488), both heads on the
//     CELL state (491-492) and Gaussian.update (642-643)
#include "common.cuh"

namespace iod {

// =====================================================================================
// weight repack kernels (run once)
// =====================================================================================
// OIHW [CO][CI][k][k] -> [chunk][tap][ci_in_chunk][co] (forward) and the data-gradient
// version (roles of ci/co swapped, taps flipped): wt[chunk][tap][co_in_chunk][ci] with
// tap' = (k-1-dy, k-1-dx).  ci_off/ci_cnt select a channel window of the source (the first
// decoder layer's L+2 inputs are handled separately).
__global__ void pack_conv_kernel(const float* __restrict__ w, float* __restrict__ fwd,
                                 float* __restrict__ bwd, int CO, int CI, int KS, int CK) {
  const int total = CO * CI * KS * KS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int dx = i % KS, dy = (i / KS) % KS, ci = (i / (KS * KS)) % CI, co = i / (KS * KS * CI);
    const float v = w[i];
    if (fwd) {
      const int ckf = CK < CI ? CK : CI;
      const int ch = ci / ckf, cc = ci % ckf;
      fwd[(((size_t)ch * KS * KS + dy * KS + dx) * ckf + cc) * CO + co] = v;
    }
    if (bwd) {  // input of the dgrad conv = co axis, output = ci axis
      const int ckb = CK < CO ? CK : CO;
      const int ch = co / ckb, cc = co % ckb;
      const int tap = (KS - 1 - dy) * KS + (KS - 1 - dx);
      bwd[(((size_t)ch * KS * KS + tap) * ckb + cc) * CI + ci] = v;
    }
  }
}

// decoder.conv [4][C][k][k] -> out_w [tap][ci][4]
__global__ void pack_out4_kernel(const float* __restrict__ w, float* __restrict__ out_w, int C, int KS) {
  const int total = 4 * C * KS * KS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int dx = i % KS, dy = (i / KS) % KS, ci = (i / (KS * KS)) % C, o = i / (KS * KS * C);
    out_w[((size_t)(dy * KS + dx) * C + ci) * 4 + o] = w[i];
  }
}

// refine conv OIHW [CO][CI][k][k] -> [tap][CIP][CO] with CIP >= CI (zero padded)
__global__ void pack_refine_kernel(const float* __restrict__ w, float* __restrict__ o, int CO, int CI,
                                   int CIP, int KS) {
  const int total = KS * KS * CIP * CO;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % CO, ci = (i / CO) % CIP, tap = i / (CO * CIP);
    o[i] = (ci < CI) ? w[((size_t)co * CI + ci) * KS * KS + tap] : 0.f;
  }
}

// Layer-1 collapse.  For border class (cy,cx) the valid taps are those that stay inside the
// image; wsum[cls][co][ci] = sum of W1[co][ci][dy][dx] over them (ci < L).
__global__ void pack_wsum_kernel(const float* __restrict__ w1, float* __restrict__ wsum, float* __restrict__ wsumT,
                                 int C, int L, int KS) {
  const int P = KS / 2, CI = L + 2;
  const int total = KS * KS * C * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % L, co = (i / L) % C, cls = i / (L * C);
    const int cy = cls / KS, cx = cls % KS;
    const int dy_lo = (cy < P) ? P - cy : 0, dy_hi = (cy > P) ? KS - 1 - (cy - P) : KS - 1;
    const int dx_lo = (cx < P) ? P - cx : 0, dx_hi = (cx > P) ? KS - 1 - (cx - P) : KS - 1;
    float s = 0.f;
    for (int dy = dy_lo; dy <= dy_hi; ++dy)
      for (int dx = dx_lo; dx <= dx_hi; ++dx) s += w1[(((size_t)co * CI + ci) * KS + dy) * KS + dx];
    wsum[i] = s;
    wsumT[((size_t)cls * L + ci) * C + co] = s;
  }
}

// ptab[y][x][co] = bias + zero-padded conv of the two coordinate planes (channels L, L+1 of
// the broadcast input: x = linspace(-1,1,W) along W, y along H; iodine.py:526-530).
__global__ void pack_ptab_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                 float* __restrict__ ptab, int C, int L, int KS, int H, int W) {
  const int P = KS / 2, CI = L + 2;
  const int total = H * W * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % C, x = (i / C) % W, y = i / (C * W);
    float s = b1[co];
    for (int dy = 0; dy < KS; ++dy) {
      const int yy = y + dy - P;
      if (yy < 0 || yy >= H) continue;
      const float cyv = (H > 1) ? -1.f + 2.f * (float)yy / (float)(H - 1) : -1.f;
      for (int dx = 0; dx < KS; ++dx) {
        const int xx = x + dx - P;
        if (xx < 0 || xx >= W) continue;
        const float cxv = (W > 1) ? -1.f + 2.f * (float)xx / (float)(W - 1) : -1.f;
        s += w1[(((size_t)co * CI + L) * KS + dy) * KS + dx] * cxv +
             w1[(((size_t)co * CI + L + 1) * KS + dy) * KS + dx] * cyv;
      }
    }
    ptab[i] = s;
  }
}

__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

static int copy_to(Plan* p, const float* src, float* dst, size_t n, cudaStream_t st) {
  IOD_REQUIRE(src != nullptr, "set_weights: null weight pointer");
  IOD_CHECK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  (void)p;
  return 0;
}

int launch_setup_weights(Plan* p, const IodineWeights* w, cudaStream_t st) {
  const IodineShape& s = p->s;
  const int C = p->C, L = s.L, KS = s.dec_k, Cr = p->Cr, M = p->M;
  const int ck = (KS == 3) ? 16 : 8;
  IOD_REQUIRE(w->dec_w[0] && w->dec_b[0], "set_weights: decoder layer 0 missing");
  pack_wsum_kernel<<<64, 256, 0, st>>>(w->dec_w[0], p->wsum, p->wsumT, C, L, KS);
  IOD_LAUNCH_CHECK(p);
  pack_ptab_kernel<<<256, 256, 0, st>>>(w->dec_w[0], w->dec_b[0], p->ptab, C, L, KS, s.H, s.W);
  IOD_LAUNCH_CHECK(p);
  for (int l = 1; l < s.dec_layers; ++l) {
    IOD_REQUIRE(w->dec_w[l] && w->dec_b[l], "set_weights: decoder layer %d missing", l);
    pack_conv_kernel<<<64, 256, 0, st>>>(w->dec_w[l], p->dec[l].w, p->dec[l].wt, C, C, KS, ck);
    IOD_LAUNCH_CHECK(p);
    if (copy_to(p, w->dec_b[l], p->dec[l].b, C, st)) return 1;
  }
  IOD_REQUIRE(w->dec_out_w && w->dec_out_b, "set_weights: decoder.conv missing");
  pack_out4_kernel<<<16, 256, 0, st>>>(w->dec_out_w, p->out_w, C, KS);
  IOD_LAUNCH_CHECK(p);
  // data-gradient of decoder.conv: input channels = 4 (one chunk), output = C
  pack_conv_kernel<<<16, 256, 0, st>>>(w->dec_out_w, nullptr, p->out_wt, 4, C, KS, 4);
  IOD_LAUNCH_CHECK(p);
  if (copy_to(p, w->dec_out_b, p->out_b, 4, st)) return 1;
  for (int l = 0; l < s.ref_layers; ++l) {
    IOD_REQUIRE(w->ref_w[l] && w->ref_b[l], "set_weights: refine layer %d missing", l);
    if (l == 0)
      pack_refine_kernel<<<32, 256, 0, st>>>(w->ref_w[0], p->ref_w0, Cr, 17, 20, s.ref_k);
    else
      pack_refine_kernel<<<32, 256, 0, st>>>(w->ref_w[l], p->ref_wp[l], Cr, Cr, Cr, s.ref_k);
    IOD_LAUNCH_CHECK(p);
    if (copy_to(p, w->ref_b[l], p->ref_b[l], Cr, st)) return 1;
  }
  if (copy_to(p, w->mlp_w, p->mlp_w, (size_t)M * Cr, st)) return 1;
  if (copy_to(p, w->mlp_b, p->mlp_b, M, st)) return 1;
  if (copy_to(p, w->lstm_w_ih, p->w_ih, (size_t)4 * M * (M + 4 * L), st)) return 1;
  if (copy_to(p, w->lstm_w_hh, p->w_hh, (size_t)4 * M * M, st)) return 1;
  if (copy_to(p, w->lstm_b_ih, p->b_ih, 4 * M, st)) return 1;
  if (copy_to(p, w->lstm_b_hh, p->b_hh, 4 * M, st)) return 1;
  if (copy_to(p, w->mean_w, p->head_w, (size_t)L * M, st)) return 1;
  if (copy_to(p, w->logvar_w, p->head_w + (size_t)L * M, (size_t)L * M, st)) return 1;
  if (copy_to(p, w->mean_b, p->head_b, L, st)) return 1;
  if (copy_to(p, w->logvar_b, p->head_b + L, L, st)) return 1;
  if (copy_to(p, w->init_mean, p->init_mean, L, st)) return 1;
  if (copy_to(p, w->init_logvar, p->init_logvar, L, st)) return 1;
  return 0;
}

// =====================================================================================
// sample + first decoder layer
// =====================================================================================
// z = mu + exp(logvar/2) * eps (or z given), u[n][cls][co] = sum_ci wsumT[cls][ci][co] z[n][ci]
// (thread = one (class, channel) output; consecutive threads read consecutive channels of the transposed weights)
__global__ void __launch_bounds__(256)
sample_u_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                const float* __restrict__ eps, const float* __restrict__ z_in,
                const float* __restrict__ wsumT, float* __restrict__ z_out, float* __restrict__ u,
                int L, int C, int NCC /* n_class*C */, float* __restrict__ eu, int* __restrict__ ubig) {
  extern __shared__ float sz[];
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float z;
    if (z_in) z = z_in[(size_t)n * L + i];
    else z = mu[(size_t)n * L + i] + expf(0.5f * logvar[(size_t)n * L + i]) * eps[(size_t)n * L + i];
    sz[i] = z;
    z_out[(size_t)n * L + i] = z;
  }
  __syncthreads();
  int big = 0;
  for (int o = threadIdx.x; o < NCC; o += blockDim.x) {
    const int cls = o / C, co = o - cls * C;
    const float* wr = wsumT + (size_t)cls * L * C + co;
    float s0 = 0.f, s1 = 0.f;
    int i = 0;
    for (; i + 1 < L; i += 2) {
      s0 = fmaf(__ldg(wr + (size_t)i * C), sz[i], s0);
      s1 = fmaf(__ldg(wr + (size_t)(i + 1) * C), sz[i + 1], s1);
    }
    if (i < L) s0 = fmaf(__ldg(wr + (size_t)i * C), sz[i], s0);
    const float uv = s0 + s1;
    u[(size_t)n * NCC + o] = uv;
    // tc_layer1_kernel forms exp(u + p) as exp(u) * exp(p) (conv_tc.cu): safe while both factors stay far from the ends
    // of the fp32 range -- outside |u| <= IOD_L1_EXP_RANGE the slot is flagged and takes the evaluated exponential
    if (eu) {
      eu[(size_t)n * NCC + o] = expf(fminf(fmaxf(uv, -IOD_L1_EXP_RANGE), IOD_L1_EXP_RANGE));
      big |= !(fabsf(uv) <= IOD_L1_EXP_RANGE);
    }
  }
  if (ubig) {
    big = __syncthreads_or(big);
    if (threadIdx.x == 0) ubig[n] = big;
  }
}

// act0[n][y][x][co] = ELU(u[n][class(y,x)][co] + ptab[y][x][co])
template <typename OutT>
__global__ void __launch_bounds__(256)
layer1_kernel(const float* __restrict__ u, const float* __restrict__ ptab, OutT* __restrict__ act0,
              int H, int W, int C, int KS) {
  const int n = blockIdx.y;
  const int P = KS / 2;
  const int c4n = C / 4;
  const int total = H * W * c4n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c4 = i % c4n, pix = i / c4n;
    const int y = pix / W, x = pix % W;
    const int cls = border_class(y, H, P) * KS + border_class(x, W, P);
    const float4 uv = *reinterpret_cast<const float4*>(u + ((size_t)n * KS * KS + cls) * C + c4 * 4);
    const float4 pv = *reinterpret_cast<const float4*>(ptab + (size_t)pix * C + c4 * 4);
    float4 r;
    r.x = elu_f(uv.x + pv.x); r.y = elu_f(uv.y + pv.y);
    r.z = elu_f(uv.z + pv.z); r.w = elu_f(uv.w + pv.w);
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(act0) + ((size_t)n * H * W + pix) * C + c4 * 4) = r;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(act0) + ((size_t)n * H * W + pix) * C + c4 * 4) = pk;
    }
  }
}

int launch_sample_l1(Plan* p, const float* mu, const float* logvar, const float* eps,
                     const float* z_in, float* act0, cudaStream_t st) {
  const IodineShape& s = p->s;
  const int ncc = p->n_class * p->C;
  sample_u_kernel<<<p->BK, 256, s.L * sizeof(float), st>>>(mu, logvar, eps, z_in, p->wsumT, p->z, p->u,
                                                           s.L, p->C, ncc, tc_mode(p) ? p->eu : nullptr,
                                                           tc_mode(p) ? p->ubig : nullptr);
  IOD_LAUNCH_CHECK(p);
  const int per = p->HW * (p->C / 4);
  dim3 grid((per + 255) / 256 > 1024 ? 1024 : (per + 255) / 256, p->BK);
  if (tc_mode(p)) return tc_launch_layer1(p, act0, st);
  layer1_kernel<float><<<grid, 256, 0, st>>>(p->u, p->ptab, act0, s.H, s.W, p->C, s.dec_k);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// =====================================================================================
// posterior gradients, KL, latent vector
// =====================================================================================
// dz[n][ci] = sum_{cls,co} wsum[cls][co][ci] * G[n][cls][co]      (layer-1 dgrad, collapsed)
// dmu = dz - mu ; dlogvar = dz * 0.5 exp(logvar/2) eps - 0.5 (exp(logvar) - 1)   (row A4)
// xin[n][M + ...] = [mu, logvar, LN3(dmu), LN3(dlogvar)]   (iodine.py:253-275; unbiased std)
__global__ void __launch_bounds__(256)
post_grads_kernel(const float* __restrict__ G, const float* __restrict__ wsum,
                  const float* __restrict__ mu, const float* __restrict__ logvar,
                  const float* __restrict__ eps, float* __restrict__ dz_out,
                  float* __restrict__ xin, double* __restrict__ accum,
                  int L, int NCC, int M, int layernorm, float* __restrict__ raw_out /* [N][2L] or null */,
                  const double* __restrict__ stats, float* __restrict__ lnp, int HW) {
  extern __shared__ float sm[];
  float* sG = sm;            // [NCC]
  float* sa = sm + NCC;      // [L] dmu
  float* sb = sa + L;        // [L] dlogvar
  float* sp = sb + L;        // [parts][L] partial dz
  __shared__ float red[8];
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < NCC; i += blockDim.x) sG[i] = G[(size_t)n * NCC + i];
  __syncthreads();
  // dz: thread (part, ci) sums every parts-th (class, channel) row of wsum (coalesced over ci)
  const int parts = (int)blockDim.x / L > 0 ? (int)blockDim.x / L : 1;
  {
    const int part = threadIdx.x / L, ci0 = threadIdx.x % L;
    if (part < parts) {
      for (int ci = ci0; ci < L; ci += (parts > 1 ? L : (int)blockDim.x)) {
        float d0 = 0.f, d1 = 0.f;
        int o = part;
        for (; o + parts < NCC; o += 2 * parts) {
          d0 = fmaf(__ldg(wsum + (size_t)o * L + ci), sG[o], d0);
          d1 = fmaf(__ldg(wsum + (size_t)(o + parts) * L + ci), sG[o + parts], d1);
        }
        if (o < NCC) d0 = fmaf(__ldg(wsum + (size_t)o * L + ci), sG[o], d0);
        sp[part * L + ci] = d0 + d1;
      }
    }
  }
  __syncthreads();
  float klp = 0.f;
  for (int ci = threadIdx.x; ci < L; ci += blockDim.x) {
    float d = 0.f;
    for (int q = 0; q < parts; ++q) d += sp[q * L + ci];
    const float m = mu[(size_t)n * L + ci], lv = logvar[(size_t)n * L + ci], e = eps[(size_t)n * L + ci];
    dz_out[(size_t)n * L + ci] = d;
    sa[ci] = d - m;
    sb[ci] = d * 0.5f * expf(0.5f * lv) * e - 0.5f * (expf(lv) - 1.f);
    if (raw_out) {                                     // training tape: dJ/d(posterior) before the layer-norm
      raw_out[(size_t)n * 2 * L + ci] = sa[ci];
      raw_out[(size_t)n * 2 * L + L + ci] = sb[ci];
    }
    klp += 0.5f * (expf(lv) + m * m - 1.f - lv);                       // iodine.py:657-658
  }
  klp = warp_sum(klp);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = klp;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)red[w];
    atomicAdd(&accum[1], t);
  }
  // layer-norm over L (unbiased std), serial per block: L is at most a few hundred
  __shared__ float st[4];
  if (threadIdx.x < 2) {
    const float* v = threadIdx.x == 0 ? sa : sb;
    float m = 0.f;
    for (int i = 0; i < L; ++i) m += v[i];
    m /= (float)L;
    float q = 0.f;
    for (int i = 0; i < L; ++i) q += (v[i] - m) * (v[i] - m);
    const float sd = sqrtf(q / (float)(L > 1 ? L - 1 : 1));
    st[threadIdx.x * 2] = layernorm ? m : 0.f;
    st[threadIdx.x * 2 + 1] = layernorm ? 1.f / (sd + 1e-5f) : 1.f;
  }
  // (mean, 1 / (std + 1e-5)) of the four image-shaped layer-norms of get_input_encoding (iodine.py:376-395, biased std
  // over C,H,W) from the sums mixture_kernel accumulated: finalised once per slot here, so that the fused first
  // refinement layer (refine_tc.cu) reads eight floats instead of redoing the f64 arithmetic per work item
  if (lnp && threadIdx.x >= 32 && threadIdx.x < 36) {
    const int g = threadIdx.x - 32;
    float m_f = 0.f, is_f = 1.f;
    if (layernorm) {
      const double cnt = (g == 0) ? 3.0 * HW : (double)HW;
      const double m = stats[((size_t)n * 4 + g) * 2] / cnt;
      double var = stats[((size_t)n * 4 + g) * 2 + 1] / cnt - m * m;
      if (var < 0.0) var = 0.0;
      m_f = (float)m;
      is_f = 1.f / ((float)sqrt(var) + 1e-5f);
    }
    lnp[(size_t)n * 8 + g] = m_f;
    lnp[(size_t)n * 8 + 4 + g] = is_f;
  }
  __syncthreads();
  float* row = xin + (size_t)n * (M + 4 * L) + M;
  for (int ci = threadIdx.x; ci < L; ci += blockDim.x) {
    row[ci] = mu[(size_t)n * L + ci];
    row[L + ci] = logvar[(size_t)n * L + ci];
    row[2 * L + ci] = (sa[ci] - st[0]) * st[1];
    row[3 * L + ci] = (sb[ci] - st[2]) * st[3];
  }
}

// latent_out: optional [BK][2L] copy of the raw posterior gradients (dmu | dlogvar), kept by the training tape
int launch_post_grads(Plan* p, const float* mu, const float* logvar, const float* eps,
                      float* latent_out, cudaStream_t st) {
  const int ncc = p->n_class * p->C, L = p->s.L;
  const int parts = 256 / L > 0 ? 256 / L : 1;
  const size_t smem = (size_t)(ncc + 2 * L + parts * L) * sizeof(float);
  post_grads_kernel<<<p->BK, 256, smem, st>>>(p->G, p->wsum, mu, logvar, eps, p->dz, p->xin, p->accum,
                                              L, ncc, p->M, p->s.layernorm, latent_out, p->stats, p->lnp, p->HW);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// KL only (iodine_elbo / no-gradient passes)
__global__ void kl_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int n,
                          double* __restrict__ accum) {
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float m = mu[i], lv = logvar[i];
    s += 0.5f * (expf(lv) + m * m - 1.f - lv);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(&accum[1], (double)s);
}

int launch_kl(Plan* p, const float* mu, const float* logvar, cudaStream_t st) {
  kl_kernel<<<32, 256, 0, st>>>(mu, logvar, p->BK * p->s.L, p->accum);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// =====================================================================================
// refinement head
// =====================================================================================
// Y[N][O] (ldy) = act( X[N][I] (ldx) * W[O][I]^T + b  [+ X2[N][I2] * W2[O][I2]^T + b2] )
// ACT: 0 none, 1 ELU(ELU(.)) (MLP's own ELU at iodine.py:565 and the extra one at 485)
// RM = rows per thread (tile = 16*RM rows x 64 columns): small row tiles keep all SMs busy when N is
// only a few hundred slot-images.
template <int ACT, int RM>
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ X, int ldx, int I, const float* __restrict__ Wt,
              const float* __restrict__ b, const float* __restrict__ X2, int ldx2, int I2,
              const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ Y,
              int ldy, int N, int O, int kchunk, size_t zstride, float* __restrict__ Y_inner = nullptr) {
  constexpr int TM = 16 * RM, TN = 64, TK = 16;
  __shared__ float sx[TK][TM + 4];
  __shared__ float sw[TK][TN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int r0 = blockIdx.y * TM, c0 = blockIdx.x * TN;
  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // split-K: block z owns [z*kc, (z+1)*kc) of the concatenated reduction axis [X | X2] and writes its
  // partial sums to Y + z*zstride (summed by the consumer; deterministic, no atomics)
  const int kc = (int)kchunk, kz_lo = (int)blockIdx.z * kc, kz_hi = kz_lo + kc;
  Y += (size_t)blockIdx.z * zstride;
  for (int pass = 0; pass < 2; ++pass) {
    const float* Xp = pass ? X2 : X;
    const float* Wp = pass ? W2 : Wt;
    const int ld = pass ? ldx2 : ldx, In = pass ? I2 : I;
    if (!Xp) continue;
    const int base = pass ? I : 0;                       // offset of this operand on the concatenated axis
    const int lo = (kz_lo > base ? kz_lo : base) - base;
    const int hi = ((kz_hi < base + In) ? kz_hi : base + In) - base;
    for (int k0 = lo; k0 < hi; k0 += TK) {
      __syncthreads();
      for (int i = threadIdx.x; i < TM * TK; i += 256) {
        const int kk = i % TK, r = i / TK;
        sx[kk][r] = (r0 + r < N && k0 + kk < hi) ? Xp[(size_t)(r0 + r) * ld + k0 + kk] : 0.f;
      }
      for (int i = threadIdx.x; i < TN * TK; i += 256) {
        const int kk = i % TK, r = i / TK;
        sw[kk][r] = (c0 + r < O && k0 + kk < hi) ? Wp[(size_t)(c0 + r) * In + k0 + kk] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float xv[RM], wv[4];
#pragma unroll
        for (int i = 0; i < RM; ++i) xv[i] = sx[kk][ty * RM + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = sw[kk][tx * 4 + i];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int r = r0 + ty * RM + i;
    if (r >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= O) continue;
      float v = acc[i][j] + (blockIdx.z == 0 ? b[c] + (b2 ? b2[c] : 0.f) : 0.f);
      if (ACT == 1) {
        v = elu_f(v);
        if (Y_inner) Y_inner[(size_t)r * O + c] = v;     // training tape: the MLP's own ELU, before the extra one
        v = elu_f(v);
      }
      Y[(size_t)r * ldy + c] = v;
    }
  }
}

// LSTMCell pointwise (torch gate order i,f,g,o): c' = s(f) c + s(i) tanh(g), h' = s(o) tanh(c')
// gates arrive as `parts` split-K partial sums [parts][N][4M]
__global__ void lstm_pointwise_kernel(const float* __restrict__ gates, float* __restrict__ h,
                                      float* __restrict__ c, int N, int M, int parts,
                                      float* __restrict__ gsum_out /* [N][4M] summed gates or null */) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * M; i += gridDim.x * blockDim.x) {
    const int n = i / M, j = i % M;
    float gs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < parts; ++q) {
      const float* g = gates + ((size_t)q * N + n) * 4 * M;
      gs[0] += g[j]; gs[1] += g[M + j]; gs[2] += g[2 * M + j]; gs[3] += g[3 * M + j];
    }
    if (gsum_out) {
      float* go = gsum_out + (size_t)n * 4 * M;
      go[j] = gs[0]; go[M + j] = gs[1]; go[2 * M + j] = gs[2]; go[3 * M + j] = gs[3];
    }
    const float ig = sigmoid_f(gs[0]), fg = sigmoid_f(gs[1]), gg = tanhf(gs[2]), og = sigmoid_f(gs[3]);
    const float cn = fg * c[i] + ig * gg;
    c[i] = cn;
    h[i] = og * tanhf(cn);
  }
}

// Gaussian.update (iodine.py:642-643): mu += delta[:, :L], logvar += delta[:, L:]
// (delta arrives as `parts` split-K partial sums [parts][N][2L])
__global__ void update_kernel(const float* __restrict__ delta, float* __restrict__ mu,
                              float* __restrict__ logvar, int N, int L, int parts) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * L; i += gridDim.x * blockDim.x) {
    const int n = i / L, j = i % L;
    float dm = 0.f, dl = 0.f;
    for (int q = 0; q < parts; ++q) {
      const float* d = delta + ((size_t)q * N + n) * 2 * L;
      dm += d[j];
      dl += d[L + j];
    }
    mu[i] += dm;
    logvar[i] += dl;
  }
}

int launch_head(Plan* p, float* mu, float* logvar, float* h, float* c, cudaStream_t st) {
  const int N = p->BK, M = p->M, L = p->s.L, Cr = p->Cr, I = M + 4 * L;
  // MLP + double ELU, written into the first M columns of xin (torch.cat at iodine.py:487)
  {
    dim3 grid((M + 63) / 64, (N + 15) / 16);
    linear_kernel<1, 1><<<grid, 256, 0, st>>>(p->pool, Cr, Cr, p->mlp_w, p->mlp_b, nullptr, 0, 0, nullptr,
                                              nullptr, p->xin, I, N, M, Cr, 0, p->tape_u);
    IOD_LAUNCH_CHECK(p);
  }
  {
    const int ktot = I + M, kchunk = ((ktot + LSTM_KSPLIT - 1) / LSTM_KSPLIT + 15) / 16 * 16;
    dim3 grid((4 * M + 63) / 64, (N + 63) / 64, LSTM_KSPLIT);
    linear_kernel<0, 4><<<grid, 256, 0, st>>>(p->xin, I, I, p->w_ih, p->b_ih, h, M, M, p->w_hh, p->b_hh,
                                              p->gates, 4 * M, N, 4 * M, kchunk, (size_t)N * 4 * M);
    IOD_LAUNCH_CHECK(p);
  }
  lstm_pointwise_kernel<<<(N * M + 255) / 256, 256, 0, st>>>(p->gates, h, c, N, M, LSTM_KSPLIT, p->tape_gates);
  IOD_LAUNCH_CHECK(p);
  {  // both heads read the CELL state (iodine.py:488-492); delta reuses the gates buffer
    const int hk = ((M + LSTM_KSPLIT - 1) / LSTM_KSPLIT + 15) / 16 * 16;
    dim3 grid((2 * L + 63) / 64, (N + 15) / 16, LSTM_KSPLIT);
    linear_kernel<0, 1><<<grid, 256, 0, st>>>(c, M, M, p->head_w, p->head_b, nullptr, 0, 0, nullptr, nullptr,
                                              p->gates, 2 * L, N, 2 * L, hk, (size_t)N * 2 * L);
    IOD_LAUNCH_CHECK(p);
  }
  update_kernel<<<(N * L + 255) / 256, 256, 0, st>>>(p->gates, mu, logvar, N, L, LSTM_KSPLIT);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// Gaussian.init_unit (iodine.py:615-616) + lstm_hidden = None
__global__ void init_state_kernel(const float* __restrict__ im, const float* __restrict__ il,
                                  float* __restrict__ mu, float* __restrict__ lv, float* __restrict__ h,
                                  float* __restrict__ c, int N, int L, int M) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * L; i += gridDim.x * blockDim.x) {
    mu[i] = im[i % L];
    lv[i] = il[i % L];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * M; i += gridDim.x * blockDim.x) {
    h[i] = 0.f;
    c[i] = 0.f;
  }
}

int launch_init_state(Plan* p, float* mu, float* lv, float* h, float* c, cudaStream_t st) {
  init_state_kernel<<<64, 256, 0, st>>>(p->init_mean, p->init_logvar, mu, lv, h, c, p->BK, p->s.L, p->M);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

}  // namespace iod
