// mixture.cu -- the fused pixel-mixture kernel ("K4" of SURVEY.md) and its consumers.
//
// One pass over the decoder's 4-channel output computes, per pixel and for all K slots in
// registers: the mask softmax (reference lib/modeling/iodine.py:185), the per-channel
// Gaussian log-likelihood (661-666), the per-channel mixture logsumexp (213-216), the
// closed-form gradients that (B*elbo).backward() (iodine.py:90) leaves in mean.grad /
// mask.grad, their chain through sigmoid / softmax down to the decoder's pre-activation
// outputs ("seed4"), and every raw auxiliary channel of get_input_encoding() (243-343):
// mask_posterior (289-292), pixel likelihood (309-312) and leave-one-out likelihood
// (321-328), plus the sums the parameter-free layer-norm (376-395) needs.  No [B,K,C,H,W]
// intermediate of the reference (K_log_likelihood, log(mask), r, ...) touches HBM.
//
// mixture_kernel: libm arithmetic (exact fp32 mode).  mixture_fast_kernel: the same pass for the tensor-core modes
// (a third fewer instructions, MUFU exp / log).  Both write either raw fp32 aux channels (`auxs`, consumed by
// assemble_kernel / assemble16_kernel below: fp32 mode, training tape, shapes outside the fused first layer) or, FUSED,
// the refinement network's input in the form refine_l0f_kernel (refine_tc.cu) consumes -- then there is no assembly pass.
#include <stdlib.h>

#include "common.cuh"

namespace iod {

// Hardware exponential / logarithm for the tensor-core precision modes (FAST): MUFU.EX2 / MUFU.LG2 with the
// argument product split exactly (x * log2(e) = t + r, r from one FMA), so the result keeps ~2 ulp over the
// whole range, and WITHOUT flush-to-zero: the reference's un-stabilised exp(sum ll) (iodine.py:290) lives in the
// subnormals for pixels no slot explains.  The exact fp32 mode keeps libm.
template <bool FAST>
__device__ __forceinline__ float mix_exp(float x) {
  if constexpr (!FAST) return expf(x);
  const float t = x * 1.4426950408889634f;
  float r = fmaf(x, 1.4426950408889634f, -t);
  r = fmaf(x, 1.9259629911266175e-8f, r);
  float e;
  asm("ex2.approx.f32 %0, %1;" : "=f"(e) : "f"(t));
  return e * fmaf(r, 0.6931471805599453f, 1.f);
}
template <bool FAST>
__device__ __forceinline__ float mix_log(float x) {
  if constexpr (!FAST) return logf(x);
  float l;
  asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(x));
  return l * 0.6931471805599453f;
}
template <bool FAST>
__device__ __forceinline__ float mix_rcp(float x) {
  if constexpr (!FAST) return 1.f / x;
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <bool FAST>
__device__ __forceinline__ float mix_sigmoid(float v) { return mix_rcp<FAST>(1.f + mix_exp<FAST>(-v)); }

// auxs layout per slot-pixel (12 floats): 0-2 mean rgb | 3 mask | 4 logit | 5 mask_post |
// 6-8 dJ/dmean rgb | 9 dJ/dmask | 10 leave-one-out | 11 unused
// FUSED (tensor-core refinement encoder, refine_tc.cu: refine_l0f_kernel): the same 48 bytes per slot-pixel hold the
// refinement network's input stack in the form its first layer consumes -- no separate assembly pass:
//   pl0  uint4  [n][HW]  channels 0-7 of the stack, FINAL 16-bit values: image rgb | mean rgb | mask | logit
//   raw4 float4 [n][HW]  mask_posterior | dJ/dmean rgb         (fp32: the last three still lack their layer-norm,
//   raw2 float2 [n][HW]  dJ/dmask | leave-one-out               whose statistics only exist once this grid finishes)
// (pl0 at auxs, raw4 at auxs + 4 BK HW floats, raw2 at auxs + 8 BK HW floats)
template <int KMAX, bool FUSED, bool FAST>
__global__ void __launch_bounds__(128)
mixture_kernel(const float* __restrict__ out4, const float* __restrict__ x,
               float* __restrict__ seed4, float* __restrict__ auxs, float* __restrict__ lik,
               double* __restrict__ stats, double* __restrict__ accum,
               int K, int HW, float inv_2s2, float inv_s2, float ll_const, int want_grads,
               int seed_half /* 0: fp32 float4, 1: bf16 x8, 2: fp16 x8 */,
               int Kl, int k_off, int ll_on, int f16, size_t n_slot_pix) {
  // K-split: out4 holds all K slots of every image (out4_slot), this rank keeps the outputs of slots
  // [k_off, k_off + Kl) under LOCAL slot numbers; whole images: Kl == K, k_off == 0.  ll_on: this rank adds the
  // image log-likelihood to the step's sum (K-split: exactly one rank does).
  const int b = blockIdx.y, B = gridDim.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pix < HW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = 4;

  float lg[KMAX], mr[KMAX], mg[KMAX], mb[KMAX];   // logit -> mask ; pre-sigmoid -> mean
  float xr = 0.f, xg = 0.f, xb = 0.f;
  if (live) {
    xr = x[((size_t)b * 3 + 0) * HW + pix];
    xg = x[((size_t)b * 3 + 1) * HW + pix];
    xb = x[((size_t)b * 3 + 2) * HW + pix];
  }
  float lmax = -INFINITY;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    lg[k] = -INFINITY; mr[k] = mg[k] = mb[k] = 0.f;
    if (k < K && live) {
      const float4 v = reinterpret_cast<const float4*>(out4)[out4_slot(b, k, B, Kl, HW) + pix];
      mr[k] = mix_sigmoid<FAST>(v.x); mg[k] = mix_sigmoid<FAST>(v.y); mb[k] = mix_sigmoid<FAST>(v.z);
      lg[k] = v.w;
      lmax = fmaxf(lmax, v.w);
    }
  }
  // ---- mask = softmax_K(logits)                                      (iodine.py:185)
  float logit_raw[KMAX];
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    logit_raw[k] = lg[k];
    if (k < K && live) { lg[k] = mix_exp<FAST>(lg[k] - lmax); den += lg[k]; }
  }
  const float inv_den = live ? mix_rcp<FAST>(den) : 0.f;
  // ---- per-channel a_kc = log(mask+1e-12) + ll_kc ; s_c = logsumexp_k  (210-216)
  float amax_r = -INFINITY, amax_g = -INFINITY, amax_b = -INFINITY;
  float lm[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    lm[k] = 0.f;
    if (k < K && live) {
      lg[k] *= inv_den;                              // lg now holds the mask
      lm[k] = mix_log<FAST>(lg[k] + 1e-12f);
      const float dr = xr - mr[k], dg = xg - mg[k], db = xb - mb[k];
      amax_r = fmaxf(amax_r, lm[k] - dr * dr * inv_2s2 + ll_const);
      amax_g = fmaxf(amax_g, lm[k] - dg * dg * inv_2s2 + ll_const);
      amax_b = fmaxf(amax_b, lm[k] - db * db * inv_2s2 + ll_const);
    }
  }
  float sr = 0.f, sg = 0.f, sb = 0.f;
  float kl_tot = 0.f, mkl_tot = 0.f;                 // sum_k Lk_k ; sum_k mask_k Lk_k
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K && live) {
      const float dr = xr - mr[k], dg = xg - mg[k], db = xb - mb[k];
      const float llr = -dr * dr * inv_2s2 + ll_const, llg = -dg * dg * inv_2s2 + ll_const,
                  llb = -db * db * inv_2s2 + ll_const;
      sr += mix_exp<FAST>(lm[k] + llr - amax_r);
      sg += mix_exp<FAST>(lm[k] + llg - amax_g);
      sb += mix_exp<FAST>(lm[k] + llb - amax_b);
      const float Lk = mix_exp<FAST>(llr + llg + llb);        // un-stabilised, as the reference (290)
      kl_tot += Lk;
      mkl_tot += lg[k] * Lk;
    }
  }
  float s_r = 0.f, s_g = 0.f, s_b = 0.f;
  if (live) { s_r = amax_r + mix_log<FAST>(sr); s_g = amax_g + mix_log<FAST>(sg); s_b = amax_b + mix_log<FAST>(sb); }
  const float ll_pix = s_r + s_g + s_b;              // summed over channels (220)

  // ---- block reduction of the log-likelihood
  __shared__ double red[NW];
  {
    float v = warp_sum(live ? ll_pix : 0.f);
    if (lane == 0) red[warp] = (double)v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += red[w];
      if (ll_on) atomicAdd(&accum[0], t);
    }
  }
  if (!want_grads) return;

  // ---- gradients + aux channels
  const float likv = live ? mix_exp<FAST>(ll_pix) : 0.f;      // exp(sum_c s_c)           (309-310)
  if (live) lik[(size_t)b * HW + pix] = likv;
  float gm[KMAX];                                     // dJ/dmask_k
  float mgsum = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    gm[k] = 0.f;
    if (k < K && live) {
      const float dr = xr - mr[k], dg = xg - mg[k], db = xb - mb[k];
      // exp(ll_kc - s_c) = r_kc / (mask_k + 1e-12)
      const float er = mix_exp<FAST>(-dr * dr * inv_2s2 + ll_const - s_r);
      const float eg = mix_exp<FAST>(-dg * dg * inv_2s2 + ll_const - s_g);
      const float eb = mix_exp<FAST>(-db * db * inv_2s2 + ll_const - s_b);
      gm[k] = er + eg + eb;
      mgsum += lg[k] * gm[k];
    }
  }
  // Layer-norm sums.  Every slot contributes six sums over the block's pixels (grad_means: sum, sum^2 over the
  // three channels; grad_mask; leave-one-out) and the image two (pixel likelihood).  Slots are handled in batches
  // of MIX_KB = 5 (30 sums, + the 2 likelihood sums in the first batch = 32): a TRANSPOSED warp reduction brings
  // the 32 per-lane partials down to one total per lane in 31 shuffles (lane l ends up with the total of sum l)
  // instead of 5 shuffles per sum -- the kernel was bound by the shuffle unit (220 shuffles per pixel at K = 7).
  constexpr int MIX_KB = 5;
  constexpr int NBATCH = (KMAX + MIX_KB - 1) / MIX_KB;
  __shared__ double s_stat[NBATCH][32];
  for (int i = threadIdx.x; i < NBATCH * 32; i += blockDim.x) (&s_stat[0][0])[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int bt = 0; bt < NBATCH; ++bt) {
    if (bt * MIX_KB < K) {                          // block-uniform
      float vb[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) vb[i] = 0.f;
      if (bt == 0) { vb[30] = likv; vb[31] = likv * likv; }
#pragma unroll
      for (int kk = 0; kk < MIX_KB; ++kk) {
        const int k = bt * MIX_KB + kk;             // compile-time
        if (k < KMAX && k < K && live && k >= k_off && k < k_off + Kl) {
          const float mk = lg[k], m_r = mr[k], m_g = mg[k], m_b = mb[k], gmk = gm[k], lgr = logit_raw[k];
          const float dr = xr - m_r, dg = xg - m_g, db = xb - m_b;
          const float llr = -dr * dr * inv_2s2 + ll_const, llg = -dg * dg * inv_2s2 + ll_const,
                      llb = -db * db * inv_2s2 + ll_const;
          const float me = mk + 1e-12f;
          const float er = mix_exp<FAST>(llr - s_r), eg = mix_exp<FAST>(llg - s_g), eb = mix_exp<FAST>(llb - s_b);
          // dJ/dmean_kc = r_kc (x_c - mean_kc)/sigma^2 with r_kc = (mask+1e-12) * exp(ll - s)
          const float gr = me * er * dr * inv_s2, gg = me * eg * dg * inv_s2, gb = me * eb * db * inv_s2;
          const float Lk = mix_exp<FAST>(llr + llg + llb);
          const float mpost = Lk / kl_tot;                                   // (292) 0/0 -> NaN as ref
          const float loo = (mkl_tot - mk * Lk) / (1.f - mk + 1e-5f);        // (326-328)
          const size_t sp = ((size_t)(b * Kl + (k - k_off))) * HW + pix;
          if constexpr (FUSED) {
            reinterpret_cast<uint4*>(auxs)[sp] = make_uint4(pack_h2(xr, xg, f16), pack_h2(xb, m_r, f16),
                                                            pack_h2(m_g, m_b, f16), pack_h2(mk, lgr, f16));
            reinterpret_cast<float4*>(auxs)[n_slot_pix + sp] = make_float4(mpost, gr, gg, gb);
            reinterpret_cast<float2*>(auxs)[4 * n_slot_pix + sp] = make_float2(gmk, loo);
          } else {
            float4* ax = reinterpret_cast<float4*>(auxs) + sp * 3;
            ax[0] = make_float4(m_r, m_g, m_b, mk);
            ax[1] = make_float4(lgr, mpost, gr, gg);
            ax[2] = make_float4(gb, gmk, loo, 0.f);
          }
          // chain to the decoder's raw outputs: sigmoid' and softmax'
          const float4 sd = make_float4(gr * m_r * (1.f - m_r), gg * m_g * (1.f - m_g),
                                        gb * m_b * (1.f - m_b), mk * (gmk - mgsum));
          if (seed_half) {   // one 8-channel 16-bit plane for the tensor-core data-gradient (conv_tc.cu)
            reinterpret_cast<uint4*>(seed4)[sp] = make_uint4(pack_h2(sd.x, sd.y, seed_half == 2),
                                                             pack_h2(sd.z, sd.w, seed_half == 2), 0u, 0u);
          } else {
            reinterpret_cast<float4*>(seed4)[sp] = sd;
          }
          vb[6 * kk + 0] = gr + gg + gb;
          vb[6 * kk + 1] = gr * gr + gg * gg + gb * gb;
          vb[6 * kk + 2] = gmk;
          vb[6 * kk + 3] = gmk * gmk;
          vb[6 * kk + 4] = loo;
          vb[6 * kk + 5] = loo * loo;
        }
      }
      // transposed reduction: after the round with mask m, lanes with bit m clear hold the lower half of the
      // surviving sums and lanes with it set the upper half; lane l finishes with the total of sum l
#pragma unroll
      for (int m = 16, n = 32; m >= 1; m >>= 1, n >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
          const float send = up ? vb[i] : vb[i + n / 2];
          const float keep = up ? vb[i + n / 2] : vb[i];
          vb[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
      }
      atomicAdd(&s_stat[bt][lane], (double)vb[0]);
    }
  }
  __syncthreads();
  // block totals -> global f64 sums.  stats[n][group][2]: group 0 grad_means, 1 grad_mask, 2 likelihood, 3 loo
  for (int i = threadIdx.x; i < NBATCH * 32; i += blockDim.x) {
    const int bt = i >> 5, l = i & 31;
    const double v = s_stat[bt][l];
    if (l < 30) {
      const int k = bt * MIX_KB + l / 6, j = l % 6;
      if (k >= k_off && k < k_off + Kl) {
        const int grp = j >> 1, g = (grp == 2) ? 3 : grp;
        atomicAdd(&stats[((size_t)(b * Kl + (k - k_off)) * 4 + g) * 2 + (j & 1)], v);
      }
    } else if (bt == 0) {
      // likelihood statistics are per image; every slot of the image gets the same numbers
      for (int k = 0; k < Kl; ++k) atomicAdd(&stats[((size_t)(b * Kl + k) * 4 + 2) * 2 + (l - 30)], v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same pass for the tensor-core precision modes: identical outputs up to ~1e-6 relative, a third of the
// instructions (mixture_kernel is instruction-issue bound: ~490 per slot-pixel, most of them in 15 libm exponentials
// and their range handling).  What changes:
//   * every slot's p_kc = exp(a_kc - max_k a_kc) is computed ONCE and kept in registers: the responsibilities are
//     r_kc = p_kc / sum_k p_kc and exp(ll_kc - s_c) = r_kc / (mask_k + 1e-12), so the second and third evaluation of
//     the three per-channel exponentials (dJ/dmask, dJ/dmean) become two multiplications; L_k = exp(sum_c ll_kc) is
//     kept as well;
//   * exponentials / logarithms / reciprocals are MUFU.EX2 / LG2 / RCP.  Flush-to-zero forms where a flushed result is
//     indistinguishable (softmax terms next to the +1e-12, sigmoid, p_kc <= 1); the un-stabilised L_k and the pixel
//     likelihood (iodine.py:290, 309) keep the exact-product, subnormal-preserving form mix_exp<true>;
//   * mask_posterior = L_k / sum_k L_k with ONE reciprocal per pixel, the operands pre-scaled by 2^100 when the sum is
//     below 2^-100 (so that a sum in the subnormals still divides; 0 / 0 stays NaN as in the reference).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_ftz(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float lg2_fast(float x) { float l; asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(x)); return l; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_sub(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// F16: IEEE half (else bfloat16) for the 16-bit outputs (pl0; the seeds when seed_half != 0)
template <int KMAX, bool FUSED, bool F16>
__global__ void __launch_bounds__(128, KMAX <= 8 ? 5 : (KMAX <= 12 ? 3 : 2))
mixture_fast_kernel(const float* __restrict__ out4, const float* __restrict__ x,
                    float* __restrict__ seed4, float* __restrict__ auxs, float* __restrict__ lik,
                    double* __restrict__ stats, double* __restrict__ accum,
                    int K, int HW, float inv_2s2, float inv_s2, float ll_const, int want_grads,
                    int seed_half, int Kl, int k_off, int ll_on, int /*f16*/, size_t n_slot_pix) {
  constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const int b = blockIdx.y, B = gridDim.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;   // HW is a multiple of the block size (launch_mixture): every thread owns a pixel
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = 4;

  float mk[KMAX], mr[KMAX], mg[KMAX], mb[KMAX];     // mask, mean rgb
  float pr[KMAX], pg[KMAX], pb[KMAX];               // a_kc, then p_kc = exp(a_kc - max)
  float Lk[KMAX];                                   // exp(sum_c ll_kc), un-stabilised
  float lgr[KMAX];                                  // raw mask logits (an aux channel)
  const float xr = x[((size_t)b * 3 + 0) * HW + pix];
  const float xg = x[((size_t)b * 3 + 1) * HW + pix];
  const float xb = x[((size_t)b * 3 + 2) * HW + pix];
  // slot k of image b inside out4_all (out4_slot), stepped without divisions: + HW inside a rank's block of Kl slots,
  // + (B Kl - Kl + 1) HW across the rank boundary; 32-bit float4 indices (BK HW < 2^32)
  const float4* o4 = reinterpret_cast<const float4*>(out4) + pix;
  const uint32_t so_first = (uint32_t)b * (uint32_t)Kl * (uint32_t)HW;
  const uint32_t so_wrap = ((uint32_t)B * (uint32_t)Kl - (uint32_t)Kl + 1u) * (uint32_t)HW;
  float lmax = -INFINITY;
  {
    // all K loads first (one round trip to HBM instead of K), then the arithmetic
    float4 v[KMAX];
    uint32_t so = so_first; int kk = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      v[k] = make_float4(0.f, 0.f, 0.f, -INFINITY);
      if (k < K) {
        v[k] = o4[so];
        if (++kk == Kl) { kk = 0; so += so_wrap; } else so += (uint32_t)HW;
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      mk[k] = -INFINITY; lgr[k] = 0.f; mr[k] = mg[k] = mb[k] = 0.f; pr[k] = pg[k] = pb[k] = 0.f; Lk[k] = 0.f;
      if (k < K) {
        mr[k] = rcp_fast(1.f + ex2_ftz(-v[k].x * L2E));
        mg[k] = rcp_fast(1.f + ex2_ftz(-v[k].y * L2E));
        mb[k] = rcp_fast(1.f + ex2_ftz(-v[k].z * L2E));
        mk[k] = v[k].w; lgr[k] = v[k].w;
        lmax = fmaxf(lmax, v[k].w);
      }
    }
  }
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) { mk[k] = ex2_ftz((mk[k] - lmax) * L2E); den += mk[k]; }
  const float inv_den = rcp_fast(den);
  float amax_r = -INFINITY, amax_g = -INFINITY, amax_b = -INFINITY;
  float kl_tot = 0.f, mkl_tot = 0.f;                 // sum_k L_k ; sum_k mask_k L_k
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      mk[k] *= inv_den;                              // mask = softmax_K(logits)            (iodine.py:185)
      const float lm = lg2_fast(mk[k] + 1e-12f) * LN2;
      const float dr = xr - mr[k], dg = xg - mg[k], db = xb - mb[k];
      const float llr = fmaf(-dr * dr, inv_2s2, ll_const), llg = fmaf(-dg * dg, inv_2s2, ll_const),
                  llb = fmaf(-db * db, inv_2s2, ll_const);
      pr[k] = lm + llr; pg[k] = lm + llg; pb[k] = lm + llb;       // a_kc                     (210-216)
      amax_r = fmaxf(amax_r, pr[k]); amax_g = fmaxf(amax_g, pg[k]); amax_b = fmaxf(amax_b, pb[k]);
      Lk[k] = mix_exp<true>(llr + llg + llb);                     // un-stabilised, as the reference (290)
      kl_tot += Lk[k];
      mkl_tot = fmaf(mk[k], Lk[k], mkl_tot);
    }
  }
  float sr = 0.f, sg = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      pr[k] = ex2_ftz((pr[k] - amax_r) * L2E); sr += pr[k];
      pg[k] = ex2_ftz((pg[k] - amax_g) * L2E); sg += pg[k];
      pb[k] = ex2_ftz((pb[k] - amax_b) * L2E); sb += pb[k];
    }
  }
  const float ll_pix =                               // sum_c logsumexp_k a_kc            (213-220)
      (amax_r + lg2_fast(sr) * LN2) + (amax_g + lg2_fast(sg) * LN2) + (amax_b + lg2_fast(sb) * LN2);

  __shared__ double red[NW];
  {
    float v = warp_sum(ll_pix);
    if (lane == 0) red[warp] = (double)v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += red[w];
      if (ll_on) atomicAdd(&accum[0], t);
    }
  }
  if (!want_grads) return;

  const float likv = mix_exp<true>(ll_pix);       // exp(sum_c s_c)           (309-310)
  lik[(size_t)b * HW + pix] = likv;
  const float isr = rcp_fast(sr), isg = rcp_fast(sg), isb = rcp_fast(sb);
  // dJ/dmask_k = sum_c exp(ll_kc - s_c) = sum_c r_kc / (mask + 1e-12); recomputed per slot below rather than kept
  float mgsum = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      pr[k] *= isr; pg[k] *= isg; pb[k] *= isb;       // p now holds the responsibilities r_kc
      mgsum = fmaf(mk[k], (pr[k] + pg[k] + pb[k]) * rcp_fast(mk[k] + 1e-12f), mgsum);
    }
  }
  // mask_posterior denominators: one reciprocal, operands pre-scaled out of the subnormals
  const float kscale = (kl_tot < 7.8886090522101181e-31f) ? 1.2676506002282294e30f : 1.f;   // 2^-100, 2^100
  const float inv_kl = rcp_sub(kl_tot * kscale);

  constexpr int MIX_KB = 5;
  constexpr int NBATCH = (KMAX + MIX_KB - 1) / MIX_KB;
  __shared__ double s_stat[NBATCH][32];
  for (int i = threadIdx.x; i < NBATCH * 32; i += blockDim.x) (&s_stat[0][0])[i] = 0.0;
  __syncthreads();
  const uint32_t sp0 = (uint32_t)b * (uint32_t)Kl * (uint32_t)HW + (uint32_t)pix;   // output index of local slot 0
  uint4* const a_pl0 = reinterpret_cast<uint4*>(auxs);
  float4* const a_r4 = reinterpret_cast<float4*>(auxs) + n_slot_pix;
  float2* const a_r2 = reinterpret_cast<float2*>(auxs) + 4 * n_slot_pix;
#pragma unroll
  for (int bt = 0; bt < NBATCH; ++bt) {
    if (bt * MIX_KB < K) {                          // block-uniform
      float vb[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) vb[i] = 0.f;
      if (bt == 0) { vb[30] = likv; vb[31] = likv * likv; }
#pragma unroll
      for (int kk = 0; kk < MIX_KB; ++kk) {
        const int k = bt * MIX_KB + kk;             // compile-time
        if (k < KMAX && k < K && k >= k_off && k < k_off + Kl) {
          const float m_k = mk[k], m_r = mr[k], m_g = mg[k], m_b = mb[k];
          const float gmk = (pr[k] + pg[k] + pb[k]) * rcp_fast(m_k + 1e-12f);
          const float dr = xr - m_r, dg = xg - m_g, db = xb - m_b;
          // dJ/dmean_kc = r_kc (x_c - mean_kc) / sigma^2
          const float gr = pr[k] * dr * inv_s2, gg = pg[k] * dg * inv_s2, gb = pb[k] * db * inv_s2;
          const float mpost = (Lk[k] * kscale) * inv_kl;                             // (292) 0/0 -> NaN as ref
          const float loo = (mkl_tot - m_k * Lk[k]) * rcp_fast(1.f - m_k + 1e-5f);   // (326-328)
          const uint32_t sp = sp0 + (uint32_t)(k - k_off) * (uint32_t)HW;
          if constexpr (FUSED) {
            a_pl0[sp] = make_uint4(pack_h2(xr, xg, F16), pack_h2(xb, m_r, F16), pack_h2(m_g, m_b, F16), pack_h2(m_k, lgr[k], F16));
            a_r4[sp] = make_float4(mpost, gr, gg, gb);
            a_r2[sp] = make_float2(gmk, loo);
          } else {
            float4* ax = reinterpret_cast<float4*>(auxs) + (size_t)sp * 3;
            ax[0] = make_float4(m_r, m_g, m_b, m_k);
            ax[1] = make_float4(lgr[k], mpost, gr, gg);
            ax[2] = make_float4(gb, gmk, loo, 0.f);
          }
          const float4 sd = make_float4(gr * m_r * (1.f - m_r), gg * m_g * (1.f - m_g),
                                        gb * m_b * (1.f - m_b), m_k * (gmk - mgsum));
          if (seed_half) {
            reinterpret_cast<uint4*>(seed4)[sp] = make_uint4(pack_h2(sd.x, sd.y, F16), pack_h2(sd.z, sd.w, F16), 0u, 0u);
          } else {
            reinterpret_cast<float4*>(seed4)[sp] = sd;
          }
          vb[6 * kk + 0] = gr + gg + gb;
          vb[6 * kk + 1] = gr * gr + gg * gg + gb * gb;
          vb[6 * kk + 2] = gmk;
          vb[6 * kk + 3] = gmk * gmk;
          vb[6 * kk + 4] = loo;
          vb[6 * kk + 5] = loo * loo;
        }
      }
#pragma unroll
      for (int m = 16, n = 32; m >= 1; m >>= 1, n >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
          const float send = up ? vb[i] : vb[i + n / 2];
          const float keep = up ? vb[i + n / 2] : vb[i];
          vb[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
      }
      atomicAdd(&s_stat[bt][lane], (double)vb[0]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NBATCH * 32; i += blockDim.x) {
    const int bt = i >> 5, l = i & 31;
    const double v = s_stat[bt][l];
    if (l < 30) {
      const int k = bt * MIX_KB + l / 6, j = l % 6;
      if (k >= k_off && k < k_off + Kl) {
        const int grp = j >> 1, g = (grp == 2) ? 3 : grp;
        atomicAdd(&stats[((size_t)(b * Kl + (k - k_off)) * 4 + g) * 2 + (j & 1)], v);
      }
    } else if (bt == 0) {
      for (int k = 0; k < Kl; ++k) atomicAdd(&stats[((size_t)(b * Kl + k) * 4 + 2) * 2 + (l - 30)], v);
    }
  }
}

int launch_mixture(Plan* p, const float* x, bool want_grads, cudaStream_t st, bool fused_aux) {
  const IodineShape& s = p->s;
  const float sg = s.sigma;
  const float inv_2s2 = 1.f / (2.f * sg * sg), inv_s2 = 1.f / (sg * sg);
  const float ll_const = -logf(sg) - 0.5f * logf(2.f * 3.14159265358979323846f);
  IOD_CHECK_CUDA(cudaMemsetAsync(p->accum, 0, 2 * sizeof(double), st));
  if (want_grads)
    IOD_CHECK_CUDA(cudaMemsetAsync(p->stats, 0, (size_t)p->BK * 8 * sizeof(double), st));
  dim3 grid((p->HW + 127) / 128, s.B);
  const int seed_half = (tc_mode(p) && !tf_mode(p)) ? (s.precision == IODINE_FP16 ? 2 : 1) : 0;
  const int k_off = p->ks_rank * s.K, ll_on = p->ks_rank == 0;
  // hardware exp / log in the tensor-core modes (the exact fp32 mode keeps libm; IODINE_MIX_EXACT=1 keeps it everywhere)
  static const bool exact_env = getenv("IODINE_MIX_EXACT") != nullptr;
  const bool fast = tc_mode(p) && !exact_env && p->HW % 128 == 0;   // (mixture_fast_kernel: no partial blocks)
  const bool fused = fused_aux && want_grads;
  const size_t nsp = (size_t)p->BK * p->HW;
#define IOD_MIX_ARGS                                                                                               \
  p->out4_all, x, p->seed4, p->auxs, p->lik, p->stats, p->accum, p->K_total, p->HW, inv_2s2, inv_s2, ll_const,       \
      want_grads, seed_half, s.K, k_off, ll_on, half_is_f16(p), nsp
#define IOD_MIX(KM, FU)                                                                  \
  do {                                                                                   \
    if (fast && f16o) mixture_fast_kernel<KM, FU, true><<<grid, 128, 0, st>>>(IOD_MIX_ARGS);   \
    else if (fast) mixture_fast_kernel<KM, FU, false><<<grid, 128, 0, st>>>(IOD_MIX_ARGS);     \
    else mixture_kernel<KM, FU, false><<<grid, 128, 0, st>>>(IOD_MIX_ARGS);               \
  } while (0)
  const bool f16o = half_is_f16(p) != 0;           // (seed_half == 1, bf16 seeds, only occurs with bf16 outputs)
  if (want_grads) plan_prof_mark(p, st, 2);
  if (p->K_total <= 8) { if (fused) IOD_MIX(8, true); else IOD_MIX(8, false); }
  else if (p->K_total <= 12) { if (fused) IOD_MIX(12, true); else IOD_MIX(12, false); }
  else { if (fused) IOD_MIX(16, true); else IOD_MIX(16, false); }
#undef IOD_MIX
#undef IOD_MIX_ARGS
  if (want_grads) plan_prof_mark(p, st, 2);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// -------------------------------------------------------------------------------------
// assemble: the refinement network's 17-channel input (get_input_encoding, iodine.py:277-340)
// in NHWC with 3 zero pad channels, layer-norm (376-395, biased std over C,H,W per slot)
// applied from the sums the mixture kernel accumulated.  Channel order is the reference's
// code order: image 0-2 | means 3-5 | mask 6 | logits 7 | mask_post 8 | grad_means 9-11 |
// grad_mask 12 | likelihood 13 | leave-one-out 14 | x-coord 15 | y-coord 16.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
assemble_kernel(const float* __restrict__ auxs, const float* __restrict__ lik,
                const float* __restrict__ x, const double* __restrict__ stats,
                float* __restrict__ enc20, int K, int H, int W, int layernorm) {
  const int n = blockIdx.y, b = n / K;
  const int HW = H * W;
  __shared__ float s_mu[4], s_is[4];
  if (threadIdx.x < 4) {
    float mu = 0.f, is = 1.f;
    if (layernorm) {
      const double cnt = (threadIdx.x == 0) ? 3.0 * HW : (double)HW;
      const double sum = stats[((size_t)n * 4 + threadIdx.x) * 2 + 0];
      const double sq = stats[((size_t)n * 4 + threadIdx.x) * 2 + 1];
      const double m = sum / cnt;
      double var = sq / cnt - m * m;
      if (var < 0.0) var = 0.0;
      mu = (float)m;
      is = 1.f / ((float)sqrt(var) + 1e-5f);
    }
    s_mu[threadIdx.x] = mu; s_is[threadIdx.x] = is;
  }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const int y = pix / W, xx = pix % W;
  const float4* ax = reinterpret_cast<const float4*>(auxs) + ((size_t)n * HW + pix) * 3;
  const float4 a0 = ax[0], a1 = ax[1], a2 = ax[2];
  const float xr = x[((size_t)b * 3 + 0) * HW + pix], xg = x[((size_t)b * 3 + 1) * HW + pix],
              xb = x[((size_t)b * 3 + 2) * HW + pix];
  const float lk = lik[(size_t)b * HW + pix];
  const float cxv = (W > 1) ? -1.f + 2.f * (float)xx / (float)(W - 1) : -1.f;
  const float cyv = (H > 1) ? -1.f + 2.f * (float)y / (float)(H - 1) : -1.f;
  float4* o = reinterpret_cast<float4*>(enc20) + ((size_t)n * HW + pix) * 5;
  o[0] = make_float4(xr, xg, xb, a0.x);
  o[1] = make_float4(a0.y, a0.z, a0.w, a1.x);
  o[2] = make_float4(a1.y, (a1.z - s_mu[0]) * s_is[0], (a1.w - s_mu[0]) * s_is[0],
                     (a2.x - s_mu[0]) * s_is[0]);
  o[3] = make_float4((a2.y - s_mu[1]) * s_is[1], (lk - s_mu[2]) * s_is[2],
                     (a2.z - s_mu[3]) * s_is[3], cxv);
  o[4] = make_float4(cyv, 0.f, 0.f, 0.f);
}

int launch_assemble(Plan* p, const float* x, cudaStream_t st) {
  dim3 grid((p->HW + 255) / 256, p->BK);
  assemble_kernel<<<grid, 256, 0, st>>>(p->auxs, p->lik, x, p->stats, p->enc20, p->s.K, p->s.H,
                                        p->s.W, p->s.layernorm);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// 16-bit modes: the same stack for the tensor-core refinement encoder (refine_tc.cu) -- the 15 DATA channels
// in reference order, chunk-planar 16-bit [n][2][H][W][8] (plane 0 = channels 0-7, plane 1 = 8-14 + one zero).
// The two coordinate channels (15, 16) do not depend on the data; their convolution is folded into the
// per-position bias table of the encoder's first layer.
__global__ void __launch_bounds__(256)
assemble16_kernel(const float* __restrict__ auxs, const float* __restrict__ lik,
                  const float* __restrict__ x, const double* __restrict__ stats,
                  uint4* __restrict__ enc16, int K, int HW, int layernorm, int f16) {
  const int n = blockIdx.y, b = n / K;
  __shared__ float s_mu[4], s_is[4];
  if (threadIdx.x < 4) {
    float mu = 0.f, is = 1.f;
    if (layernorm) {
      const double cnt = (threadIdx.x == 0) ? 3.0 * HW : (double)HW;
      const double sum = stats[((size_t)n * 4 + threadIdx.x) * 2 + 0];
      const double sq = stats[((size_t)n * 4 + threadIdx.x) * 2 + 1];
      const double m = sum / cnt;
      double var = sq / cnt - m * m;
      if (var < 0.0) var = 0.0;
      mu = (float)m;
      is = 1.f / ((float)sqrt(var) + 1e-5f);
    }
    s_mu[threadIdx.x] = mu; s_is[threadIdx.x] = is;
  }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const float4* ax = reinterpret_cast<const float4*>(auxs) + ((size_t)n * HW + pix) * 3;
  const float4 a0 = ax[0], a1 = ax[1], a2 = ax[2];
  const float xr = x[((size_t)b * 3 + 0) * HW + pix], xg = x[((size_t)b * 3 + 1) * HW + pix],
              xb = x[((size_t)b * 3 + 2) * HW + pix];
  const float lk = lik[(size_t)b * HW + pix];
  uint4 o0, o1;
  o0.x = pack_h2(xr, xg, f16);
  o0.y = pack_h2(xb, a0.x, f16);
  o0.z = pack_h2(a0.y, a0.z, f16);
  o0.w = pack_h2(a0.w, a1.x, f16);
  o1.x = pack_h2(a1.y, (a1.z - s_mu[0]) * s_is[0], f16);
  o1.y = pack_h2((a1.w - s_mu[0]) * s_is[0], (a2.x - s_mu[0]) * s_is[0], f16);
  o1.z = pack_h2((a2.y - s_mu[1]) * s_is[1], (lk - s_mu[2]) * s_is[2], f16);
  o1.w = pack_h2((a2.z - s_mu[3]) * s_is[3], 0.f, f16);
  enc16[((size_t)n * 2 + 0) * HW + pix] = o0;
  enc16[((size_t)n * 2 + 1) * HW + pix] = o1;
}

int launch_assemble16(Plan* p, const float* x, cudaStream_t st) {
  dim3 grid((p->HW + 255) / 256, p->BK);
  assemble16_kernel<<<grid, 256, 0, st>>>(p->auxs, p->lik, x, p->stats, reinterpret_cast<uint4*>(p->enc16), p->s.K,
                                          p->HW, p->s.layernorm, half_is_f16(p));
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// tests only: reference layout [B,K,17,H,W] followed by latent [B,K,4L]
__global__ void export_aux_kernel(const float* __restrict__ enc20, const uint16_t* __restrict__ enc16,
                                  const float* __restrict__ xin, float* __restrict__ aux_out, int BK, int H, int W,
                                  int M, int L4, int f16) {
  const int HW = H * W;
  const size_t total = (size_t)BK * 17 * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int pix = i % HW;
    const int c = (i / HW) % 17;
    const int n = i / ((size_t)HW * 17);
    if (!enc16) {
      aux_out[i] = enc20[((size_t)n * HW + pix) * 20 + c];
    } else if (c < 15) {
      const uint16_t u = enc16[(((size_t)n * 2 + c / 8) * HW + pix) * 8 + c % 8];
      aux_out[i] = f16 ? __half2float(*reinterpret_cast<const __half*>(&u))
                       : __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&u));
    } else {                                       // coordinate channels are implicit in the 16-bit modes
      const int yy = pix / W, xx = pix % W;
      aux_out[i] = (c == 15) ? ((W > 1) ? -1.f + 2.f * (float)xx / (float)(W - 1) : -1.f)
                             : ((H > 1) ? -1.f + 2.f * (float)yy / (float)(H - 1) : -1.f);
    }
  }
  const size_t tl = (size_t)BK * L4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tl;
       i += (size_t)gridDim.x * blockDim.x) {
    const int n = i / L4, j = i % L4;
    aux_out[total + i] = xin[(size_t)n * (M + L4) + M + j];
  }
}

// tests only, fused aux path: the 17 channels as refine_l0f_kernel's producers build them (same arithmetic, same
// 16-bit rounding) from mixture_kernel<FUSED>'s output
__global__ void export_aux_fused_kernel(const uint4* __restrict__ pl0, const float4* __restrict__ raw4,
                                        const float2* __restrict__ raw2, const float* __restrict__ lik,
                                        const double* __restrict__ stats, const float* __restrict__ xin,
                                        float* __restrict__ aux_out, int BK, int K, int H, int W, int M, int L4,
                                        int layernorm, int f16) {
  const int HW = H * W;
  const size_t total = (size_t)BK * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = i % HW, n = i / HW;
    float mu[4], is[4];
    for (int g = 0; g < 4; ++g) {
      mu[g] = 0.f; is[g] = 1.f;
      if (layernorm) {
        const double cnt = (g == 0) ? 3.0 * HW : (double)HW;
        const double m = stats[((size_t)n * 4 + g) * 2] / cnt;
        double var = stats[((size_t)n * 4 + g) * 2 + 1] / cnt - m * m;
        if (var < 0.0) var = 0.0;
        mu[g] = (float)m;
        is[g] = 1.f / ((float)sqrt(var) + 1e-5f);
      }
    }
    const uint4 a = pl0[i];
    const float4 r4 = raw4[i];
    const float2 r2 = raw2[i];
    const float lk = lik[(size_t)(n / K) * HW + pix];
    uint4 o1;
    o1.x = pack_h2(r4.x, (r4.y - mu[0]) * is[0], f16);
    o1.y = pack_h2((r4.z - mu[0]) * is[0], (r4.w - mu[0]) * is[0], f16);
    o1.z = pack_h2((r2.x - mu[1]) * is[1], (lk - mu[2]) * is[2], f16);
    o1.w = pack_h2((r2.y - mu[3]) * is[3], 0.f, f16);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, o1.x, o1.y, o1.z, o1.w};
    float* o = aux_out + (size_t)n * 17 * HW + pix;
    for (int q = 0; q < 8; ++q) {
      const float2 v = unpack_h2(w[q], f16);
      o[(size_t)(2 * q) * HW] = v.x;
      if (2 * q + 1 < 15) o[(size_t)(2 * q + 1) * HW] = v.y;
    }
    const int yy = pix / W, xx = pix % W;
    o[(size_t)15 * HW] = (W > 1) ? -1.f + 2.f * (float)xx / (float)(W - 1) : -1.f;
    o[(size_t)16 * HW] = (H > 1) ? -1.f + 2.f * (float)yy / (float)(H - 1) : -1.f;
  }
  const size_t base = (size_t)BK * 17 * HW;
  const size_t tl = (size_t)BK * L4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tl; i += (size_t)gridDim.x * blockDim.x) {
    const int n = i / L4, j = i % L4;
    aux_out[base + i] = xin[(size_t)n * (M + L4) + M + j];
  }
}

int launch_export_aux(Plan* p, const float* x, float* aux_out, cudaStream_t st) {
  (void)x;
  const bool rtc = rtc_enabled(p);
  if (rtc_fused_aux(p)) {
    const size_t nsp = (size_t)p->BK * p->HW;
    export_aux_fused_kernel<<<p->num_sms * 4, 256, 0, st>>>(
        reinterpret_cast<const uint4*>(p->auxs), reinterpret_cast<const float4*>(p->auxs) + nsp,
        reinterpret_cast<const float2*>(p->auxs) + 4 * nsp, p->lik, p->stats, p->xin, aux_out, p->BK, p->s.K, p->s.H,
        p->s.W, p->M, 4 * p->s.L, p->s.layernorm, half_is_f16(p));
    IOD_LAUNCH_CHECK(p);
    return 0;
  }
  export_aux_kernel<<<p->num_sms * 4, 256, 0, st>>>(p->enc20, rtc ? reinterpret_cast<const uint16_t*>(p->enc16) : nullptr,
                                                    p->xin, aux_out, p->BK, p->s.H, p->s.W, p->M, 4 * p->s.L,
                                                    half_is_f16(p));
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// -------------------------------------------------------------------------------------
// recombine: IODINE.decode's tail (iodine.py:67-71): mask = softmax_K, pred = sum_k mask*mean;
// outputs in the reference's NCHW layouts.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
recombine_kernel(const float* __restrict__ out4, float* __restrict__ pred, float* __restrict__ mask,
                 float* __restrict__ mean, uint8_t* __restrict__ amax, int K, int HW, int B, int Kl) {
  // out4: all K slots of every image (out4_slot; K-split: gathered from the ranks); outputs are complete [B,K,...]
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  float lmax = -INFINITY;
  for (int k = 0; k < K; ++k)
    lmax = fmaxf(lmax, out4[(out4_slot(b, k, B, Kl, HW) + pix) * 4 + 3]);
  float den = 0.f;
  for (int k = 0; k < K; ++k) den += expf(out4[(out4_slot(b, k, B, Kl, HW) + pix) * 4 + 3] - lmax);
  const float inv = 1.f / den;
  float pr = 0.f, pg = 0.f, pb = 0.f;
  float mbest = -1.f;
  int kbest = 0;
  for (int k = 0; k < K; ++k) {
    const float4 v = reinterpret_cast<const float4*>(out4)[out4_slot(b, k, B, Kl, HW) + pix];
    const float m = expf(v.w - lmax) * inv;
    if (m > mbest) { mbest = m; kbest = k; }         // first maximum, as torch.argmax (lib/eval/ari_eval.py:32-39)
    const float r = sigmoid_f(v.x), g = sigmoid_f(v.y), bl = sigmoid_f(v.z);
    pr += m * r; pg += m * g; pb += m * bl;
    if (mask) mask[((size_t)(b * K + k)) * HW + pix] = m;
    if (mean) {
      mean[(((size_t)(b * K + k)) * 3 + 0) * HW + pix] = r;
      mean[(((size_t)(b * K + k)) * 3 + 1) * HW + pix] = g;
      mean[(((size_t)(b * K + k)) * 3 + 2) * HW + pix] = bl;
    }
  }
  if (amax) amax[(size_t)b * HW + pix] = (uint8_t)kbest;
  if (pred) {
    pred[((size_t)b * 3 + 0) * HW + pix] = pr;
    pred[((size_t)b * 3 + 1) * HW + pix] = pg;
    pred[((size_t)b * 3 + 2) * HW + pix] = pb;
  }
}

int launch_recombine(Plan* p, float* pred, float* mask, float* mean, int n_images, cudaStream_t st, uint8_t* amax) {
  dim3 grid((p->HW + 255) / 256, n_images);
  recombine_kernel<<<grid, 256, 0, st>>>(p->out4_all, pred, mask, mean, amax, p->K_total, p->HW, p->s.B, p->s.K);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

}  // namespace iod
