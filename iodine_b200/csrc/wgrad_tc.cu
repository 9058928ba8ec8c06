// wgrad_tc.cu -- tcgen05 weight gradients of the decoder's 3x3 convolutions (training step, csrc/train.cu).
//
//   dW[co][ci][dy][dx] += coef * sum_{n,y,x} g[n,co,y,x] * act[n,ci,y+dy-1,x+dx-1]
//
// (torch.nn.grad.conv2d_weight as autograd runs it for nn.Conv2d, reference lib/modeling/iodine.py:422, 583, under
// loss.backward(), lib/engine/train.py:63).  A GEMM whose contraction runs over PIXELS: 270 GFLOP per C -> C layer and
// ELBO evaluation at the CLEVR6 sizes, as much as the layer's forward pass.
//
// Both operands are read exactly as they lie in HBM, chunk-planar [slot-image][planes][H][W][16 bytes] (a plane = 8
// 16-bit channels; the IODINE_TF32 mode passes fp16 copies, see wgrad_tc_supported): eight consecutive pixels of one plane are 128 contiguous bytes = one core
// matrix of an MN-MAJOR (channel-contiguous), no-swizzle UMMA operand -- 8 K rows (pixels) of 16 bytes.  LBO = 128 B
// steps to the next eight pixels, SBO = the plane pitch steps to the next channel group.  No-swizzle descriptors only
// need 16-byte aligned starts, so the horizontal tap shift dx is a +16 B on the activation operand's start address:
// no im2col, no transpose.
//   * the ring keeps image rows as [row][plane][136 pixels]; the plane pitch is uniform ACROSS the row boundary
//     (the last plane of row j is followed by plane 0 of row j+1), so ONE descriptor spans two consecutive rows:
//     C -> C layers: A = gradient rows (y, y+1) x 64 channels (M = 128), B = activation rows (y, y+1) x 64 channels
//     (N = 128).  D[(r,co)][(s,ci)] then holds four tap rows at once: (r,s) = (0,0) dy=1, (0,1) dy=2, (1,0) dy=0
//     and (1,1) dy=1 again (discarded), so every (gradient row, dy) pair is produced exactly once by the sliding
//     row pairs.  decoder.conv (C -> 4): A = activation rows (M = 128), B = the two rows of the 4-channel seed plane
//     (N = 16), dy = s - r + 1 likewise;
//   * per row pair: 3 (dx) x (128 / K pixels) tcgen05.mma into three persistent TMEM accumulators that live for the
//     whole kernel; one epilogue per CTA adds them to the fp32 gradient;
//   * four producer warps stream the rows with 2 KB bulk copies (one plane row each), mbarrier pipelined.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace iod {

constexpr int WG_PS = 136;                 // pixels per plane row in the ring: 128 + zero halo (data at 1..128)
constexpr int WG_THREADS = 32 * 9;         // 4 producers, 1 issuer, 4 epilogue warps
constexpr int WG_MAX_R = 8;

// ring rows for (M-side planes + N-side planes) per image row; one more slot mirrors slot 0 (pair (R-1, 0) contiguous)
__host__ __device__ constexpr int wg_ring_rows(int planes) {
  return (212 * 1024 / (planes * WG_PS * 16) - 1) > WG_MAX_R ? WG_MAX_R : (212 * 1024 / (planes * WG_PS * 16) - 1);
}

struct WgtParams {
  const uint4* grad;             // dJ/d(pre-activation) (C -> C) or the seed plane (decoder.conv), chunk-planar
  const uint4* act;              // input activation of the layer, chunk-planar
  const uint4* zero_row;         // >= 2 KB of zeros
  float* dw;                     // PyTorch OIHW fp32, accumulated with atomics
  const int4* itab;              // work items {slot-image, first row, rows, -}
  const int32_t* coff;           // [ranges + 1] item range of work range c
  float coef;
  int32_t H;
  int32_t nsplit;                // CTAs per work range, each taking NPL planes of the N-side operand
  int32_t grad_planes, act_planes;   // planes per slot-image of the two tensors in HBM
  int32_t cin;                   // input channels of the layer (dW's second extent)
  uint32_t idesc;
};

struct WgtSmem {
  uint64_t full[WG_MAX_R];
  uint64_t empty[WG_MAX_R];
  uint64_t done;
  uint32_t tmem_base;
};

// MPL / NPL: planes per image row of the M-side / N-side operand; TF: tf32 planes (4 channels, K = 8 pixels per MMA);
// ACT_M: the activation is the M-side operand (decoder.conv), else the N-side one (C -> C layers)
template <int MPL, int NPL, bool TF, bool ACT_M>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgtParams p) {
  constexpr int PW = TF ? 4 : 8, KPX = TF ? 8 : 16, NKS = 128 / KPX;
  constexpr int M = 2 * MPL * PW, N = 2 * NPL * PW;
  static_assert(M == 128 && N % 16 == 0 && N >= 16 && N <= 256, "operand shape");
  constexpr int R = wg_ring_rows(MPL + NPL);
  static_assert(R >= 3, "ring too small");
  constexpr int MROW16 = MPL * WG_PS, NROW16 = NPL * WG_PS;              // 16-byte units per ring row
  constexpr int MRING = (R + 1) * MROW16 * 16, NRING = (R + 1) * NROW16 * 16;
  constexpr int TCOLS = (3 * N <= 32) ? 32 : (3 * N <= 64) ? 64 : (3 * N <= 128) ? 128 : (3 * N <= 256) ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_m = smem;
  uint8_t* s_n = smem + MRING;
  WgtSmem* sb = reinterpret_cast<WgtSmem*>(smem + MRING + NRING);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int range = (int)blockIdx.x / p.nsplit, csplit = (int)blockIdx.x - range * p.nsplit;
  constexpr int W = 128;

  {  // zero both rings once: halo columns stay zero for the whole kernel
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < (MRING + NRING) / 16; i += WG_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) {
      mbar_init(smem_u32(&sb->full[i]), 4);        // four producer warps
      mbar_init(smem_u32(&sb->empty[i]), 1);
    }
    mbar_init(smem_u32(&sb->done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)), "n"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sb->tmem_base;
  const int it0 = __ldg(p.coff + range), it1 = __ldg(p.coff + range + 1);

  if (warp < 4) {
    // =============================================================== producers
    // the MPL + NPL plane rows of one image row are dealt out to the four warps
    if (elect_one_sync()) {
      constexpr int TOT = MPL + NPL;
      const int q0 = warp * TOT / 4, q1 = (warp + 1) * TOT / 4;
      const size_t plane_px = (size_t)p.H * W;
      const uint32_t ring_m = smem_u32(s_m), ring_n = smem_u32(s_n);
      int jg = 0;                                   // ring rows since the kernel started
      for (int item = it0; item < it1; ++item) {
        const int4 d = __ldg(p.itab + item);
        const int n = d.x, y0 = d.y, th = d.z;
        for (int j = 0; j < th + 2; ++j, ++jg) {
          const int slot = jg % R;
          const uint32_t ph = (uint32_t)(jg / R) & 1u;
          mbar_wait(smem_u32(&sb->empty[slot]), ph ^ 1u, 1);
          const bool mirror = slot == 0;
          const uint32_t fb = smem_u32(&sb->full[slot]);
          mbar_expect_tx(fb, (uint32_t)(q1 - q0) * 2048u * (mirror ? 2u : 1u));     // (arrives even with no plane)
          const int y = y0 - 1 + j;
          for (int q = q0; q < q1; ++q) {
            const bool m_side = q < MPL;
            const bool is_act = m_side == ACT_M;
            // gradient rows outside [y0, y0+th) belong to another work item (or lie outside the image): zeros, so
            // that every (row, dy) product is counted once; activation rows are real wherever the image has them
            const bool real = is_act ? (y >= 0 && y < p.H) : (j >= 1 && j <= th);
            const int pl = m_side ? q : (q - MPL) + csplit * NPL;             // plane of the tensor in HBM
            const int planes = is_act ? p.act_planes : p.grad_planes;
            const uint4* src = real ? (is_act ? p.act : p.grad) + ((size_t)n * planes + pl) * plane_px + (size_t)y * W : p.zero_row;
            const uint32_t dst = m_side ? ring_m + (uint32_t)(slot * MROW16 + q * WG_PS + 1) * 16u
                                        : ring_n + (uint32_t)(slot * NROW16 + (q - MPL) * WG_PS + 1) * 16u;
            bulk_load_1d(dst, src, 2048u, fb);
            if (mirror) bulk_load_1d(dst + (uint32_t)(R * (m_side ? MROW16 : NROW16)) * 16u, src, 2048u, fb);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // =============================================================== MMA issuer
    const bool leader = elect_one_sync();
    constexpr uint64_t DESC_HI = (uint64_t)((uint32_t)WG_PS | (1u << 14)) << 32;   // SBO = plane pitch, version 1
    constexpr uint32_t LBO = 8u << 16;                                            // next eight pixels: 128 B
    const uint32_t m16 = smem_u32(s_m) >> 4, n16 = smem_u32(s_n) >> 4;
    int jg = 0;
    bool first = true;
    for (int item = it0; item < it1; ++item) {
      const int th = __ldg(p.itab + item).z;
      for (int q = 0; q <= th; ++q) {               // pair q = ring rows (q, q+1) of this item
        if (q == 0) mbar_wait(smem_u32(&sb->full[jg % R]), (uint32_t)(jg / R) & 1u, 2);
        mbar_wait(smem_u32(&sb->full[(jg + q + 1) % R]), (uint32_t)((jg + q + 1) / R) & 1u, 3);
        tc_fence_after();
        const int sa = (jg + q) % R;                // (slot sa + 1 == R is the mirror of slot 0)
        const uint32_t ab = (m16 + (uint32_t)(sa * MROW16)) | LBO;
        const uint32_t bb = (n16 + (uint32_t)(sa * NROW16)) | LBO;
        if (leader) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int ks = 0; ks < NKS; ++ks) {
              const uint32_t am = (uint32_t)((ACT_M ? dx : 1) + KPX * ks), bn = (uint32_t)((ACT_M ? 1 : dx) + KPX * ks);
              tc_mma<TF>(tmem_base + (uint32_t)(dx * N), DESC_HI | (uint64_t)(ab + am), DESC_HI | (uint64_t)(bb + bn), p.idesc,
                         (first && ks == 0) ? 0u : 1u);
            }
          }
          tc_commit(smem_u32(&sb->empty[sa]));                               // row q is not read again
          if (q == th) tc_commit(smem_u32(&sb->empty[(jg + q + 1) % R]));     // nor is the item's last row
        }
        first = false;
        __syncwarp();
      }
      jg += th + 2;
    }
    if (leader) tc_commit(smem_u32(&sb->done));
    __syncwarp();
  } else {
    // =============================================================== epilogue: one pass at the end
    const int quad = warp & 3;
    mbar_wait(smem_u32(&sb->done), 0, 5);
    tc_fence_after();
    if (it1 > it0) {
      const int m = quad * 32 + lane, rm = m >> 6, chm = m & 63;   // M row = (image row of the pair, channel)
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll 1
        for (int c = 0; c < N / 16; ++c) {
          uint32_t acc[16];
          IOD_TMEM_LD16(acc, tmem_base + (uint32_t)(dx * N + c * 16) + ((uint32_t)(quad * 32) << 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int nn = c * 16 + i, rn = nn / (NPL * PW), chn = nn % (NPL * PW) + csplit * NPL * PW;
            // rows of the pair: gradient row r, activation row s; dy = s - r + 1, the second dy = 1 quadrant is dropped
            const int r = ACT_M ? rn : rm, s = ACT_M ? rm : rn;
            if (r == 1 && s == 1) continue;
            const int dy = s - r + 1;
            const int co = ACT_M ? chn : chm, ci = ACT_M ? chm : chn;
            if (ACT_M && co >= 4) continue;                                   // the seed plane carries 4 real channels
            atomicAdd(p.dw + ((size_t)(co * p.cin + ci) * 9 + dy * 3 + dx), p.coef * __uint_as_float(acc[i]));
          }
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
  }
}

// 1 when the tensor-core weight gradient can run this plan's layers: tensor-core modes, C = 64, 3x3, W = 128.
// The kernel takes 16-bit chunk-planar operands.  kind::tf32 has no MN-major form on plain (no-swizzle) rows -- its
// only MN-major shared-memory layout wants 32 channels contiguous (SWIZZLE_128B_BASE32B), the planes hold 4 -- so the
// IODINE_TF32 mode hands this kernel fp16 copies of its operands (same 10-bit mantissa as tf32; wgrad_to_h16).
int wgrad_tc_supported(const Plan* p) {
  if (!tc_mode(p) || getenv("IODINE_WGRAD_FFMA")) return 0;
  return p->C == 64 && p->s.dec_k == 3 && p->s.W == 128 && tc_rs_worklist(p, 0, nullptr, nullptr, nullptr, nullptr);
}

// tf32 chunk-planar [n][C/4][HW][4 x fp32] -> fp16 chunk-planar [n][C/8][HW][8 x fp16]; nplanes8 = planes of the result.
// A 4-channel source (the seed, [n][HW][4]) gives one plane whose upper four channels are zero.
__global__ void __launch_bounds__(256)
wgrad_to_h16_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, size_t total, int HW, int nplanes8, int src_planes4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i % HW, nk = i / HW;
    const size_t k = nk % nplanes8, n = nk / nplanes8;
    const float4 a = __ldg(src + (n * src_planes4 + 2 * k) * HW + pix);
    const float4 b = (2 * k + 1 < (size_t)src_planes4) ? __ldg(src + (n * src_planes4 + 2 * k + 1) * HW + pix)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    dst[i] = make_uint4(pack_h2(a.x, a.y, 1), pack_h2(a.z, a.w, 1), pack_h2(b.x, b.y, 1), pack_h2(b.z, b.w, 1));
  }
}
int wgrad_to_h16(Plan* p, const void* src, void* dst, int channels, cudaStream_t st) {
  const int src_planes4 = channels / 4, nplanes8 = (channels + 7) / 8;
  const size_t total = (size_t)p->BK * nplanes8 * p->HW;
  wgrad_to_h16_kernel<<<p->num_sms * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint4*>(dst), total,
                                                      p->HW, nplanes8, src_planes4);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

template <int MPL, int NPL, bool TF, bool ACT_M>
static int wgrad_launch(Plan* p, const WgtParams& q, int ranges, cudaStream_t st) {
  constexpr int R = wg_ring_rows(MPL + NPL);
  const size_t smem = (size_t)(R + 1) * (MPL + NPL) * WG_PS * 16 + sizeof(WgtSmem) + 64;
  auto kern = wgrad_tc_kernel<MPL, NPL, TF, ACT_M>;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  kern<<<ranges * q.nsplit, WG_THREADS, smem, st>>>(q);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

static uint32_t wg_idesc(const Plan* p, int N) {
  const uint32_t fmt = (p->s.precision == IODINE_BF16) ? 1u : 0u;      // (IODINE_TF32 feeds fp16 copies)
  // cute::UMMA::InstrDescriptor: f32 accumulate, a/b format, A and B MN-major (bits 15, 16), N, M = 128
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// C -> C layer: dw[64][64][3][3] += coef * conv2d_weight(act_prev, g); both operands 16-bit chunk-planar
int launch_wgrad_tc(Plan* p, const void* act_prev, const void* g, float* dw, float coef, cudaStream_t st) {
  WgtParams q;
  const bool tf = false;     // (a kind::tf32 instantiation <16, 8, true, false> compiles but cannot be MN-major, see above)
  int ranges = 0;
  const void* zero = nullptr;
  q.nsplit = 1;
  IOD_REQUIRE(tc_rs_worklist(p, 0, &q.itab, &q.coff, &ranges, &zero), "wgrad_tc: no row work list for this plan");
  q.grad = reinterpret_cast<const uint4*>(g);
  q.act = reinterpret_cast<const uint4*>(act_prev);
  q.zero_row = reinterpret_cast<const uint4*>(zero);
  q.dw = dw; q.coef = coef; q.H = p->s.H; q.cin = p->C;
  q.grad_planes = q.act_planes = p->C / (tf ? 4 : 8);
  q.idesc = wg_idesc(p, 128);
  return wgrad_launch<8, 8, false, false>(p, q, ranges, st);
}

// decoder.conv: dw[4][64][3][3] += coef * conv2d_weight(act_last, seed), seed = one 8-channel 16-bit plane (4 real)
int launch_wgrad_tc_out4(Plan* p, const void* act_last, const void* seed8, float* dw, float coef, cudaStream_t st) {
  WgtParams q;
  int ranges = 0;
  const void* zero = nullptr;
  q.nsplit = 1;
  IOD_REQUIRE(tc_rs_worklist(p, 0, &q.itab, &q.coff, &ranges, &zero), "wgrad_tc: no row work list for this plan");
  q.grad = reinterpret_cast<const uint4*>(seed8);
  q.act = reinterpret_cast<const uint4*>(act_last);
  q.zero_row = reinterpret_cast<const uint4*>(zero);
  q.dw = dw; q.coef = coef; q.H = p->s.H; q.cin = p->C;
  q.grad_planes = 1;
  q.act_planes = p->C / 8;
  q.idesc = wg_idesc(p, 16);
  return wgrad_launch<8, 1, false, true>(p, q, ranges, st);
}

}  // namespace iod
