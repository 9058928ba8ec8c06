// refine_tc.cu -- tcgen05 convolutions of the refinement encoder (16-bit precision modes).
//
// Replaces RefinementNetwork.forward's conv stack (reference lib/modeling/iodine.py:480, MultiLayerConv
// with stride, 583/592) and the tail of get_input_encoding (iodine.py:277-340) that builds its input.
//
// The encoder's convolutions are STRIDED (ARCH.REF.STRIDE = 2), so the flat-ring trick of conv_tc.cu does
// not apply (a run of output pixels is not a run of input pixels).  They are also small (1.8 % of the
// path's FLOPs, outputs of 64x64 down to 8x8 per slot-image).  Formulation:
//   * implicit GEMM, M = 128 consecutive OUTPUT positions of the flattened (slot-image, y, x) space,
//     N = Cout, K = taps x Cin; accumulator in TMEM;
//   * four producer warps gather the A operand one TAP at a time with 16-byte cp.async copies (zero
//     fill outside the image = the zero padding) straight into the no-swizzle K-major core-matrix
//     layout (one 2 KB block of 128 positions x 8 channels per channel group), several taps in flight
//     through a ring of stages; cp.async completion -> fence.proxy.async -> mbarrier arrive hands a
//     stage to the tensor core;
//   * one thread issues Cin/16 tcgen05.mma per tap against the resident weight image, tcgen05.commit
//     frees the stage;
//   * four epilogue warps read TMEM, add the bias (layer 0: a per-position table that also carries the
//     two coordinate channels, whose convolution does not depend on the data), apply ELU and store the
//     16-bit chunk-planar activation of the next layer.
// Layer 0 therefore contracts over 15 data channels padded to 16 (one K step per tap) instead of 17.
//
// Two kernels: refine_tc_kernel (the gather formulation above: layers >= 1, and layer 0 of shapes the fused kernel does
// not cover) and refine_l0f_kernel (layer 0 of stride-2 3x3 encoders on even image sizes: reads the pixel-mixture
// kernel's output directly, applies the layer-norms and packs the aux stack while staging its operand, stride-2 taps as
// unit-stride UMMA operands over phase-split planes -- see the comment above it).
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace iod {

// A-operand ring: a stage holds TPS consecutive taps of one tile (all 9 for the 16-channel first layer, so a
// tile is ONE producer->issuer handshake: a barrier wait costs ~100 cycles even when it succeeds at once),
// ST stages; a producer thread keeps ST-2 cp.async groups in flight.
constexpr int RTC_MAX_STAGES = 16;
constexpr int RTC_ACC = 4;         // TMEM accumulator stages
// 4 producer warps, 1 issuer warp, then 4 or 8 epilogue warps (two per TMEM lane quadrant, each taking half of the
// accumulator columns, when N >= 32)
__host__ __device__ constexpr int rtc_epi_warps(int N) { return N >= 32 ? 8 : 4; }
__host__ __device__ constexpr int rtc_threads(int N) { return 128 + 32 + 32 * rtc_epi_warps(N); }

struct RtcParams {
  const uint4* in;       // chunk-planar [n][cin_planes][Hin][Win] (uint4 = 8 channels of one pixel)
  uint4* out;            // chunk-planar [n][N/8][Hout][Wout]
  const void* wimg;      // [tap][ks][k-half][n][8] 16-bit
  const float* bias;     // [N]                     (tab == nullptr)
  const float* tab;      // [N/8][Hout*Wout][8] bias + coordinate-channel convolution (layer 0) or nullptr
  uint32_t w_bytes;
  int32_t Hin, Win, Hout, Wout, S, pad, KS;
  int32_t cin_planes;
  int32_t total;         // BK * Hout * Wout output positions
  int32_t tiles;
  int32_t f16;
  uint32_t idesc;
};

struct RtcSmem {
  uint64_t full[RTC_MAX_STAGES];
  uint64_t empty[RTC_MAX_STAGES];
  uint64_t tfull[RTC_ACC];
  uint64_t tempty[RTC_ACC];
  uint64_t wbar;
  uint32_t tmem_base;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int N, int NKS, int TPS, int ST>
__global__ void __launch_bounds__(rtc_threads(N), 1) refine_tc_kernel(const __grid_constant__ RtcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int TMEM_COLS = (RTC_ACC * N < 32) ? 32 : RTC_ACC * N;
  constexpr int RTC_STAGES = ST, RTC_LAG = ST - 2;
  static_assert(ST >= 3 && ST <= RTC_MAX_STAGES, "stage count");
  constexpr uint32_t TAP_BYTES = (uint32_t)(2 * NKS) * 2048u;         // cin_planes x (128 positions x 16 B)
  constexpr uint32_t STAGE_BYTES = (uint32_t)TPS * TAP_BYTES;
  const uint32_t w_region = (p.w_bytes + 1023u) & ~1023u;
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + w_region;
  RtcSmem* sb = reinterpret_cast<RtcSmem*>(s_a + (size_t)RTC_STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntaps = p.KS * p.KS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RTC_STAGES; ++i) {
      mbar_init(smem_u32(&sb->full[i]), 128);      // every producer thread arrives
      mbar_init(smem_u32(&sb->empty[i]), 1);       // tcgen05.commit
    }
    for (int i = 0; i < RTC_ACC; ++i) {
      mbar_init(smem_u32(&sb->tfull[i]), 1);
      mbar_init(smem_u32(&sb->tempty[i]), rtc_epi_warps(N));   // one arrive per epilogue warp
    }
    mbar_init(smem_u32(&sb->wbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sb->tmem_base;
  const int HWo = p.Hout * p.Wout;
  // Tile order.  With a per-position bias table (layer 0) consecutive tiles of a CTA should share their table
  // slice so that it stays in L1: position-block major (all slot-images of positions [128 rb, 128 rb + 128) before
  // the next block).  Otherwise tiles are consecutive runs of the flattened (slot-image, y, x) space.
  const int n_img = p.total / HWo;
  const bool rb_major = p.tab != nullptr && (HWo % 128) == 0;
  auto tile_pos0 = [&](int tile) -> int {
    if (!rb_major) return tile * 128;
    const int rb = tile / n_img;
    return (tile - rb * n_img) * HWo + rb * 128;
  };

  if (warp < 4) {
    // =============================================================== producers (gather)
    const int tid = threadIdx.x;                   // = position inside the tile
    const uint32_t a_base = smem_u32(s_a) + (uint32_t)tid * 16u;
    const size_t plane_in = (size_t)p.Hin * p.Win;
    int stage = 0;
    uint32_t phase = 0;
    int issued = 0, arrived_stage = 0;             // groups committed; stage of the oldest un-arrived group
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int pos = tile_pos0(tile) + tid;
      const bool valid = pos < p.total;
      const int n = valid ? pos / HWo : 0;
      const int r = valid ? pos - n * HWo : 0;
      const int y = r / p.Wout, x = r - y * p.Wout;
      const uint4* in_n = p.in + (size_t)n * (2 * NKS) * plane_in;
      int dy = 0, dx = 0;
      for (int tap0 = 0; tap0 < ntaps; tap0 += TPS) {
        mbar_wait(smem_u32(&sb->empty[stage]), phase ^ 1u, 11);
#pragma unroll
        for (int t = 0; t < TPS; ++t) {
          const int iy = y * p.S + dy - p.pad, ix = x * p.S + dx - p.pad;
          const bool inb = valid && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
          const uint4* src = inb ? in_n + (size_t)iy * p.Win + ix : p.in;
          const uint32_t dst = a_base + (uint32_t)stage * STAGE_BYTES + (uint32_t)t * TAP_BYTES;
          const uint32_t nbytes = inb ? 16u : 0u;  // 0: cp.async writes 16 zero bytes (the zero padding)
#pragma unroll
          for (int g = 0; g < 2 * NKS; ++g) cp_async16(dst + (uint32_t)g * 2048u, src + (size_t)g * plane_in, nbytes);
          if (++dx == p.KS) { dx = 0; ++dy; }
        }
        cp_async_commit();
        ++issued;
        if (issued > RTC_LAG) {                    // the group committed RTC_LAG groups ago has landed
          cp_async_wait<RTC_LAG>();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(smem_u32(&sb->full[arrived_stage]));
          if (++arrived_stage == RTC_STAGES) arrived_stage = 0;
        }
        if (++stage == RTC_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    // drain
    cp_async_wait<0>();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int left = issued < RTC_LAG ? issued : RTC_LAG;
    for (int i = 0; i < left; ++i) {
      mbar_arrive(smem_u32(&sb->full[arrived_stage]));
      if (++arrived_stage == RTC_STAGES) arrived_stage = 0;
    }
  } else if (warp == 4) {
    // =============================================================== MMA issuer (+ weight load)
    if (lane == 0) {
      const uint32_t wbar = smem_u32(&sb->wbar);
      mbar_expect_tx(wbar, p.w_bytes);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg);
      for (uint32_t off = 0; off < p.w_bytes; off += 16384u) {
        const uint32_t nb = (p.w_bytes - off < 16384u) ? (p.w_bytes - off) : 16384u;
        bulk_load_1d(smem_u32(s_w + off), src + off, nb, wbar);
      }
      mbar_wait(wbar, 0, 12);
      constexpr uint64_t DESC_HI = (uint64_t)(8u | (1u << 14)) << 32;      // SBO = 128 B, version 1
      const uint32_t a16 = (smem_u32(s_a) >> 4) | (128u << 16);           // LBO = 2 KB between the K halves
      const uint32_t w16 = (smem_u32(s_w) >> 4) | ((uint32_t)N << 16);    // LBO = N (16-byte units)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&sb->tempty[acc]), aph ^ 1u, 13);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
        for (int tap0 = 0; tap0 < ntaps; tap0 += TPS) {
          mbar_wait(smem_u32(&sb->full[stage]), phase, 14);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < TPS; ++t) {
            const uint32_t a0 = a16 + (uint32_t)stage * (STAGE_BYTES >> 4) + (uint32_t)t * (TAP_BYTES >> 4);
            const uint32_t b0 = w16 + (uint32_t)((tap0 + t) * NKS) * (uint32_t)(2 * N);
#pragma unroll
            for (int ks = 0; ks < NKS; ++ks)
              tc_mma_bf16(d_tmem, DESC_HI | (a0 + (uint32_t)ks * 256u), DESC_HI | (b0 + (uint32_t)(ks * 2 * N)), p.idesc,
                          ((tap0 + t) | ks) ? 1u : 0u);
          }
          tc_commit(smem_u32(&sb->empty[stage]));
          if (++stage == RTC_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(smem_u32(&sb->tfull[acc]));
        if (++acc == RTC_ACC) { acc = 0; aph ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // =============================================================== epilogue warps
    constexpr int EW = rtc_epi_warps(N);
    constexpr int NC = (EW == 8) ? N / 2 : N;      // accumulator columns per warp
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    const int col0 = (EW == 8) ? ((warp - 5) / 4) * NC : 0;
    const int F16 = p.f16;
    int acc = 0;
    uint32_t aph = 0;
    // bias / table values are requested one tile AHEAD: issued after the accumulator wait (ptxas does not move
    // loads across the TMEM read) they cost one exposed L2 latency per 8 channels and bounded the layer
    float4 tb[NC / 4], tb_next[NC / 4];
    auto locate = [&](int tile, bool& valid, int& n, int& r) {
      const int pos = tile_pos0(tile) + quad * 32 + lane;
      valid = tile < p.tiles && pos < p.total;
      n = valid ? pos / HWo : 0;
      r = valid ? pos - n * HWo : 0;
    };
    auto load_bias = [&](int r, float4* dst) {
#pragma unroll
      for (int k = 0; k < NC / 8; ++k) {
        const int c = col0 + k * 8;
        const float4* bp = reinterpret_cast<const float4*>(p.tab ? p.tab + ((size_t)(c >> 3) * HWo + r) * 8 : p.bias + c);
        dst[2 * k] = __ldg(bp);
        dst[2 * k + 1] = __ldg(bp + 1);
      }
    };
    bool valid; int n, r;
    locate(blockIdx.x, valid, n, r);
    load_bias(r, tb);
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      bool valid_n; int n_n, r_n;
      locate(tile + gridDim.x, valid_n, n_n, r_n);
      load_bias(r_n, tb_next);
      mbar_wait(smem_u32(&sb->tfull[acc]), aph, 15);
      tc_fence_after();
      uint32_t v[NC];
      const uint32_t taddr = tmem_base + (uint32_t)(acc * N + col0) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int q = 0; q < NC / 16; ++q) IOD_TMEM_LD16((v + q * 16), taddr + q * 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sb->tempty[acc]));      // accumulator stage is free again
      if (valid) {
#pragma unroll
        for (int k = 0; k < NC / 8; ++k) {
          float f[8];
          const int c = col0 + k * 8;
          const float4 b0 = tb[2 * k], b1 = tb[2 * k + 1];
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = elu_fast(__uint_as_float(v[k * 8 + e]) + bb[e]);
          uint4 o;
          o.x = pack_h2(f[0], f[1], F16);
          o.y = pack_h2(f[2], f[3], F16);
          o.z = pack_h2(f[4], f[5], F16);
          o.w = pack_h2(f[6], f[7], F16);
          p.out[((size_t)n * (N / 8) + (c >> 3)) * HWo + r] = o;
        }
      }
      valid = valid_n; n = n_n; r = r_n;
#pragma unroll
      for (int k = 0; k < NC / 4; ++k) tb[k] = tb_next[k];
      if (++acc == RTC_ACC) { acc = 0; aph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Layer 0 with the aux-input assembly fused into its operand load (stride 2, 3x3, even image sizes).
//
// mixture_kernel<.., FUSED> leaves the refinement input as it can know it while its grid is still running: channels
// 0-7 final (16-bit plane `pl0`), mask-posterior and the five layer-normalised channels raw in fp32 (`raw4`, `raw2`,
// per-image `lik`) next to the f64 sums of their layer-norms (get_input_encoding, iodine.py:277-340; layernorm
// 376-395; post_grads_kernel turns the sums into (mean, 1 / (std + 1e-5)) per slot).  This kernel finishes the assembly on the way into shared memory -- there is no assembled tensor in HBM
// and no separate pass -- and replaces the 16-byte gather of refine_tc_kernel for this layer:
//   * a work item is TR output rows x Wseg output columns of one slot-image.  One producer warp per input row copies
//     the item's (2 TR + 1) x (2 Wseg + 1) input pixels ONCE, coalesced, with cp.async into a staging buffer (the
//     gather read every pixel 2.25 times), one whole item ahead of its use; the same threads then apply
//     (v - mean) / (std + 1e-5) from the slot's sums, pack to 16 bits and store each pixel into one of four PHASE
//     planes (row parity x column parity) of the item's operand buffer, pitch P = Wseg + 8 positions, one zero halo
//     position on the left;
//   * in phase space a stride-2 tap is a unit-stride run: tap (dy, dx) of output (r, c) is position (r + [dy == 2]) P
//     + c + (dx == 0 ? 7 : 8) of plane (dy != 1, dx != 1).  So, as in conv_tc.cu, 128 consecutive flat positions
//     r P + c are directly a K-major no-swizzle UMMA operand (K = 16: channels 0-7 | 8-14 + one zero) and a tile is
//     9 tcgen05.mma (M = 128, N = Cr); outputs at c >= Wseg are discarded; the last tile of an item starts at
//     th P - 128 (overlapping its predecessor) so that no tile reads past the item's rows;
//   * two operand buffers: the producers fill one while the tensor core reads the other; 8 TMEM accumulator stages;
//     the epilogue (bias + coordinate-channel table, ELU, 16-bit chunk-planar store) is refine_tc_kernel's.
// ------------------------------------------------------------------------------------------------
constexpr int L0F_PROD_WARPS = 9;
constexpr int L0F_ACC = 8;
__host__ __device__ constexpr int l0f_threads(int N) { return 32 * (L0F_PROD_WARPS + 1 + rtc_epi_warps(N)); }

struct L0fParams {
  const uint4* pl0;      // [n][H][W] channels 0-7, final 16-bit
  const float4* raw4;    // [n][H][W] mask_posterior | dJ/dmean rgb
  const float2* raw2;    // [n][H][W] dJ/dmask | leave-one-out likelihood
  const float* lik;      // [b][H][W] pixel likelihood
  const float* lnp;      // [n][8] layer-norm parameters of grad_means, grad_mask, likelihood, leave-one-out:
                         // 4 means, 4 x 1 / (std + 1e-5) (finalised per slot by post_grads_kernel, head.cu)
  uint4* out;            // chunk-planar [n][N/8][Ho][Wo]
  const void* wimg;      // [tap][k-half][n][8] 16-bit
  const float* rowtab;   // [Ho][2][N] bias + convolution of the y-coordinate channel (index 0: output column 0, whose
                         // left taps fall into the padding; 1: every other column)
  uint32_t w_bytes;
  int32_t H, W, Ho, Wo, K;
  int32_t TR, Wseg, P, nseg, strips, items;
  uint32_t magicP;       // 2^32 / P + 1: pos / P = umulhi(pos, magicP) for the flat positions of an item
  float xc_step;         // 2 / (W - 1): x-coordinate channel value of image column ix = -1 + ix * xc_step
  uint32_t lbo16;        // 16-byte units between the two channel planes of a phase
  uint32_t buf16;        // 8 * lbo16: 16-byte units per operand buffer (4 phases x 2 channel planes)
  uint32_t stage_bytes;  // one staging buffer: (2 TR + 1) rows x (2 Wseg + 1) pixels x 44 bytes, rounded up to 128
  uint32_t idesc;
};

struct L0fSmem {
  uint64_t full[2];
  uint64_t empty[2];
  uint64_t tfull[L0F_ACC];
  uint64_t tempty[L0F_ACC];
  uint64_t wbar;
  uint32_t tmem_base;
};

__device__ __forceinline__ void cp_async_ca(uint32_t dst, const void* src, int bytes /* 4 or 8 */) {
  if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <int N, bool F16>
__global__ void __launch_bounds__(l0f_threads(N), 1) refine_l0f_kernel(const __grid_constant__ L0fParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int TMEM_COLS = (L0F_ACC * N < 32) ? 32 : L0F_ACC * N;
  constexpr int EW = rtc_epi_warps(N);
  const uint32_t w_region = (p.w_bytes + 1023u) & ~1023u;
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + w_region;
  uint8_t* s_stage = s_a + (size_t)2 * p.buf16 * 16u;
  L0fSmem* sb = reinterpret_cast<L0fSmem*>(s_stage + (size_t)2 * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, TR = p.TR;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sb->full[i]), 32 * L0F_PROD_WARPS);   // every producer thread arrives
      mbar_init(smem_u32(&sb->empty[i]), 1);                     // tcgen05.commit
    }
    for (int i = 0; i < L0F_ACC; ++i) {
      mbar_init(smem_u32(&sb->tfull[i]), 1);
      mbar_init(smem_u32(&sb->tempty[i]), EW);
    }
    mbar_init(smem_u32(&sb->wbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == L0F_PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // both operand buffers start as zeros: positions no producer ever writes (halo padding left of position 7, rows of
  // a short last strip) feed discarded outputs only, but must not hold NaN patterns next to real operands
  {
    uint4* a4 = reinterpret_cast<uint4*>(s_a);
    const int n16 = 2 * (int)p.buf16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) a4[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sb->tmem_base;
  const int HWo = p.Ho * p.Wo, HW = p.H * p.W;
  const int per_img = p.strips * p.nseg;

  // item -> (slot-image n, first output row y0, rows th, first output column xoff, columns wv)
  auto decode = [&](int item, int& n, int& y0, int& th, int& xoff, int& wv) {
    n = item / per_img;
    const int r = item - n * per_img;
    const int strip = r / p.nseg, seg = r - strip * p.nseg;
    y0 = strip * TR;
    th = (p.Ho - y0 < TR) ? p.Ho - y0 : TR;
    xoff = seg * p.Wseg;
    wv = (p.Wo - xoff < p.Wseg) ? p.Wo - xoff : p.Wseg;
  };
  // first flat position of tile t of an item with th rows (nt tiles): the last one is pulled back inside the item
  auto tile_pos0 = [&](int t, int nt, int th) -> int {
    if (t + 1 < nt || nt == 1) return t * 128;
    return th * P - 128;
  };

  if (warp < L0F_PROD_WARPS) {
    // =============================================================== producers: copy ahead, normalise, pack, phase-split
    const uint32_t a_base = smem_u32(s_a), st_base = smem_u32(s_stage);
    const uint32_t lbo_b = p.lbo16 * 16u, buf_b = p.buf16 * 16u;
    const int NCP = 2 * p.Wseg + 1;               // staged pixels per row: index 0 = the halo column 2 xoff - 1
    const int rows_max = 2 * TR + 1;
    const uint32_t o_pl0 = 0u, o_r4 = (uint32_t)(rows_max * NCP) * 16u, o_r2 = 2u * o_r4, o_lk = o_r2 + (uint32_t)(rows_max * NCP) * 8u;
    // staged copy of one item: every thread copies exactly the pixels it converts later (row = warp, columns = lane + 32 u;
    // lane 0 also the halo column), so its own cp.async.wait_group is all the synchronisation the staging needs
    auto prefetch = [&](int item, int sbuf) {
      int n, y0, th, xoff, wv;
      decode(item, n, y0, th, xoff, wv);
      const uint4* pl0 = p.pl0 + (size_t)n * HW;
      const float4* raw4 = p.raw4 + (size_t)n * HW;
      const float2* raw2 = p.raw2 + (size_t)n * HW;
      const float* lik = p.lik + (size_t)(n / p.K) * HW;
      const int nrows = 2 * th + 1, ncols = 2 * wv;
      const uint32_t sbase = st_base + (uint32_t)sbuf * p.stage_bytes;
      for (int lr = warp; lr < nrows; lr += L0F_PROD_WARPS) {
        const int iy = 2 * y0 - 1 + lr;
        if (iy < 0 || iy >= p.H) continue;
        const size_t rofs = (size_t)iy * p.W + 2 * xoff;
        auto stage_px = [&](int c) {              // c = input column relative to 2 xoff (-1: the halo column)
          const uint32_t idx = (uint32_t)(lr * NCP + c + 1);
          cp_async16(sbase + o_pl0 + idx * 16u, pl0 + rofs + c, 16u);
          cp_async16(sbase + o_r4 + idx * 16u, raw4 + rofs + c, 16u);
          cp_async_ca(sbase + o_r2 + idx * 8u, raw2 + rofs + c, 8);
          cp_async_ca(sbase + o_lk + idx * 4u, lik + rofs + c, 4);
        };
        for (int lc = lane; lc < ncols; lc += 32) stage_px(lc);
        if (lane == 0 && xoff > 0) stage_px(-1);
      }
    };
    int buf = 0;
    uint32_t ph = 0;
    int sbuf = 0;
    if ((int)blockIdx.x < p.items) prefetch(blockIdx.x, 0);
    cp_async_commit();
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int nxt = item + (int)gridDim.x;
      if (nxt < p.items) prefetch(nxt, sbuf ^ 1);
      cp_async_commit();
      int n, y0, th, xoff, wv;
      decode(item, n, y0, th, xoff, wv);
      const float4 mu4 = __ldg(reinterpret_cast<const float4*>(p.lnp) + (size_t)n * 2);
      const float4 is4 = __ldg(reinterpret_cast<const float4*>(p.lnp) + (size_t)n * 2 + 1);
      const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, is[4] = {is4.x, is4.y, is4.z, is4.w};
      cp_async_wait<1>();                           // this item's copies (the group before the one just committed)
      mbar_wait(smem_u32(&sb->empty[buf]), ph ^ 1u, 21);
      const uint32_t bbase = a_base + (uint32_t)buf * buf_b;
      const uint32_t sbase = st_base + (uint32_t)sbuf * p.stage_bytes;
      const int nrows = 2 * th + 1, ncols = 2 * wv;
      for (int lr = warp; lr < nrows; lr += L0F_PROD_WARPS) {
        const int iy = 2 * y0 - 1 + lr;
        const bool row_in = iy >= 0 && iy < p.H;
        // local input row lr: even -> odd image row (phase py = 1, plane row lr / 2), odd -> even image row
        const int py = (lr & 1) ^ 1, prow = lr >> 1;
        const uint32_t rbase = bbase + (uint32_t)(py * 2) * 2u * lbo_b + (uint32_t)(prow * P) * 16u;
        auto emit = [&](int c) {                    // c = input column relative to 2 xoff (-1: the halo column)
          const bool inside = row_in && (c >= 0 || xoff > 0);
          uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
          if (inside) {
            const uint32_t idx = (uint32_t)(lr * NCP + c + 1);
            float4 r4; float2 r2; float lk;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o0.x), "=r"(o0.y), "=r"(o0.z), "=r"(o0.w) : "r"(sbase + o_pl0 + idx * 16u));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r4.x), "=f"(r4.y), "=f"(r4.z), "=f"(r4.w) : "r"(sbase + o_r4 + idx * 16u));
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r2.x), "=f"(r2.y) : "r"(sbase + o_r2 + idx * 8u));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lk) : "r"(sbase + o_lk + idx * 4u));
            o1.x = pack_h2(r4.x, (r4.y - mu[0]) * is[0], F16);
            o1.y = pack_h2((r4.z - mu[0]) * is[0], (r4.w - mu[0]) * is[0], F16);
            o1.z = pack_h2((r2.x - mu[1]) * is[1], (lk - mu[2]) * is[2], F16);
            // channel 15: the x-coordinate plane (iodine.py:334-339) rides in the operand's spare slot; the y-coordinate
            // plane's convolution only depends on the output row (and on whether the left taps are padding): rowtab
            o1.w = pack_h2((r2.y - mu[3]) * is[3], fmaf((float)(2 * xoff + c), p.xc_step, -1.f), F16);
          }
          // even columns -> phase px = 0, odd -> px = 1; column 2 j (+1) sits at position j + 8, the halo at 7
          const int px = c & 1, pos = ((c + 1) >> 1) + 7 + (px ^ 1);
          const uint32_t dst = rbase + (uint32_t)px * 2u * lbo_b + (uint32_t)pos * 16u;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o0.x), "r"(o0.y), "r"(o0.z), "r"(o0.w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + lbo_b), "r"(o1.x), "r"(o1.y), "r"(o1.z), "r"(o1.w) : "memory");
        };
        for (int lc = lane; lc < ncols; lc += 32) emit(lc);
        if (lane == 0) emit(-1);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(&sb->full[buf]));
      buf ^= 1;
      if (buf == 0) ph ^= 1u;
      sbuf ^= 1;
    }
    cp_async_wait<0>();
  } else if (warp == L0F_PROD_WARPS) {
    // =============================================================== MMA issuer (+ weight load)
    if (lane == 0) {
      const uint32_t wbar = smem_u32(&sb->wbar);
      mbar_expect_tx(wbar, p.w_bytes);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg);
      for (uint32_t off = 0; off < p.w_bytes; off += 16384u) {
        const uint32_t nb = (p.w_bytes - off < 16384u) ? (p.w_bytes - off) : 16384u;
        bulk_load_1d(smem_u32(s_w + off), src + off, nb, wbar);
      }
      mbar_wait(wbar, 0, 22);
      constexpr uint64_t DESC_HI = (uint64_t)(8u | (1u << 14)) << 32;      // SBO = 128 B, version 1
      const uint32_t a16 = (smem_u32(s_a) >> 4) | (p.lbo16 << 16);         // LBO = channel-plane stride
      const uint32_t w16 = (smem_u32(s_w) >> 4) | ((uint32_t)N << 16);     // LBO = N (16-byte units)
      int buf = 0; uint32_t ph = 0;
      int acc = 0; uint32_t aph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int n, y0, th, xoff, wv;
        decode(item, n, y0, th, xoff, wv);
        const int nt = (th * P + 127) >> 7;
        mbar_wait(smem_u32(&sb->full[buf]), ph, 23);
        tc_fence_after();
        const uint32_t ab = a16 + (uint32_t)buf * p.buf16;
        for (int t = 0; t < nt; ++t) {
          mbar_wait(smem_u32(&sb->tempty[acc]), aph ^ 1u, 24);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
          const uint32_t t0 = (uint32_t)tile_pos0(t, nt, th);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const uint32_t phase = (uint32_t)((dy != 1) * 2 + (dx != 1));
            const uint32_t off = phase * 2u * p.lbo16 + (uint32_t)((dy == 2) * P + (dx == 0 ? 7 : 8)) + t0;
            tc_mma_bf16(d_tmem, DESC_HI | (ab + off), DESC_HI | (w16 + (uint32_t)(tap * 2 * N)), p.idesc, tap ? 1u : 0u);
          }
          tc_commit(smem_u32(&sb->tfull[acc]));
          if (++acc == L0F_ACC) { acc = 0; aph ^= 1u; }
        }
        tc_commit(smem_u32(&sb->empty[buf]));
        buf ^= 1;
        if (buf == 0) ph ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // =============================================================== epilogue warps
    constexpr int NC = (EW == 8) ? N / 2 : N;      // accumulator columns per warp
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    const int col0 = (EW == 8) ? ((warp - (L0F_PROD_WARPS + 1)) / 4) * NC : 0;
    int acc = 0; uint32_t aph = 0;
    // (item, tile) walker
    int item = blockIdx.x, t = 0, nt = 0, n = 0, y0 = 0, th = 0, xoff = 0, wv = 0;
    auto load_item = [&]() {
      if (item < p.items) { decode(item, n, y0, th, xoff, wv); nt = (th * P + 127) >> 7; }
      t = 0;
    };
    // this thread's output of the current tile: the positions a pulled-back last tile shares with its predecessor
    // are left to the predecessor (whole warps of the last tile then have nothing to do)
    auto locate = [&](bool& valid, int& oy, int& ox) {
      valid = false; oy = 0; ox = 0;
      if (item >= p.items) return;
      const int pos = tile_pos0(t, nt, th) + quad * 32 + lane;
      const int r = (int)__umulhi((uint32_t)pos, p.magicP), c = pos - r * P;
      valid = r < th && c < wv && (t == 0 || pos >= t * 128);
      if (valid) { oy = y0 + r; ox = xoff + c; }     // (discarded positions read table row 0)
    };
    auto advance = [&]() { if (++t >= nt) { item += gridDim.x; load_item(); } };
    load_item();
    while (item < p.items) {
      bool valid; int oy, ox;
      locate(valid, oy, ox);
      const int on = n;
      // bias + y-coordinate table row of this output: requested before the accumulator wait (L1-resident, the lanes
      // of a warp share one or two rows)
      float4 tb[NC / 4];
      {
        const float4* bp = reinterpret_cast<const float4*>(p.rowtab + ((size_t)(oy * 2 + (ox > 0 ? 1 : 0)) * N + col0));
#pragma unroll
        for (int k = 0; k < NC / 4; ++k) tb[k] = __ldg(bp + k);
      }
      mbar_wait(smem_u32(&sb->tfull[acc]), aph, 25);
      tc_fence_after();
      uint32_t v[NC];
      const uint32_t taddr = tmem_base + (uint32_t)(acc * N + col0) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int q = 0; q < NC / 16; ++q) IOD_TMEM_LD16((v + q * 16), taddr + q * 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sb->tempty[acc]));
      if (valid) {
        uint4* op = p.out + ((size_t)on * (N / 8) + (col0 >> 3)) * HWo + (size_t)oy * p.Wo + ox;
#pragma unroll
        for (int k = 0; k < NC / 8; ++k) {
          float f[8];
          const float4 b0 = tb[2 * k], b1 = tb[2 * k + 1];
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = elu_fast(__uint_as_float(v[k * 8 + e]) + bb[e]);
          uint4 o;
          o.x = pack_h2(f[0], f[1], F16);
          o.y = pack_h2(f[2], f[3], F16);
          o.z = pack_h2(f[4], f[5], F16);
          o.w = pack_h2(f[6], f[7], F16);
          op[(size_t)k * HWo] = o;
        }
      }
      advance();
      if (++acc == L0F_ACC) { acc = 0; aph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == L0F_PROD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct RtcState {
  uint16_t* w[IODINE_MAX_LAYERS];     // packed weight images
  uint32_t w_bytes[IODINE_MAX_LAYERS];
  int cin[IODINE_MAX_LAYERS];         // contracted channels per layer (16 for layer 0, Cr after)
  float* tab0 = nullptr;              // [H1*W1][Cr] layer-0 bias + coordinate-channel table
  float* rowtab = nullptr;            // [H1][2][Cr] fused layer 0: bias + y-coordinate-channel table
  bool ok = false;
  bool fused = false;                 // layer 0 = refine_l0f_kernel (reads mixture_kernel<FUSED>'s output directly)
  int f_TR = 0, f_Wseg = 0, f_P = 0, f_nseg = 0, f_strips = 0;
  uint32_t f_lbo16 = 0, f_stage = 0;
  size_t f_smem = 0;
};

static RtcState* rtc_state(Plan* p) { return reinterpret_cast<RtcState*>(p->rtc); }

// (taps per stage, stages) of the instantiation that serves (Cout, kernel size, K-steps per tap); 0 = none
static bool rtc_cfg(int Cr, int KS, int nks, int* tps, int* st) {
  if (KS == 3 && nks == 1) { *tps = 9; *st = 4; return true; }
  if (KS == 3 && nks == 2) { *tps = 3; *st = 6; return Cr == 32; }
  if (KS == 3 && nks == 4) { *tps = 3; *st = 3; return Cr == 64; }
  if (KS == 5 && nks == 1) { *tps = 5; *st = 6; return true; }
  if (KS == 5 && nks == 2) { *tps = 1; *st = 12; return Cr == 32; }
  return false;
}

static size_t rtc_smem_bytes(uint32_t w_bytes, int nks, int tps, int st) {
  return (size_t)((w_bytes + 1023u) & ~1023u) + (size_t)st * tps * (2 * nks) * 2048 + sizeof(RtcSmem) + 64;
}

// 1 if every refine layer fits the tensor-core kernel (otherwise the FFMA path of conv_f32.cu runs)
int rtc_supported(const Plan* p) {
  const IodineShape& s = p->s;
  const int Cr = p->Cr, kk = s.ref_k * s.ref_k;
  if (Cr != 16 && Cr != 32 && Cr != 64) return 0;
  for (int l = 0; l < s.ref_layers; ++l) {
    const int cin = l == 0 ? 16 : Cr;
    const uint32_t wb = (uint32_t)kk * (cin / 16) * 2u * (uint32_t)Cr * 16u;
    int tps = 0, st = 0;
    if (!rtc_cfg(Cr, s.ref_k, cin / 16, &tps, &st)) return 0;
    if (rtc_smem_bytes(wb, cin / 16, tps, st) > (size_t)227 * 1024) return 0;
  }
  return 1;
}

// Geometry of the fused layer 0 (refine_l0f_kernel): stride 2, 3x3, even image sizes.  TR output rows per work item:
// as many as two buffers of (2 TR + 1) input rows allow, fewer when the launch would otherwise leave SMs without work.
static bool l0f_geometry(const Plan* p, RtcState* st) {
  const IodineShape& s = p->s;
  if (getenv("IODINE_NO_AUX_FUSE")) return false;
  if (s.ref_k != 3 || s.ref_stride != 2 || (s.H & 1) || (s.W & 1) || s.H < 4 || s.W < 4) return false;
  const int Ho = p->ref_h[1], Wo = p->ref_w[1];
  if (Ho != s.H / 2 || Wo != s.W / 2) return false;
  const int Wseg = Wo < 64 ? Wo : 64;
  const int nseg = (Wo + Wseg - 1) / Wseg, P = Wseg + 8;
  const uint32_t w_bytes = 9u * 2u * (uint32_t)p->Cr * 16u;
  for (int TR = 8; TR >= 1; TR >>= 1) {
    if (TR > Ho && TR > 1) continue;
    const int need = ((TR + 1) * P + 8 > P + 136) ? (TR + 1) * P + 8 : P + 136;   // see tile_pos0: no tile reads past it
    const uint32_t lbo16 = (uint32_t)((need + 7) / 8 * 8);
    const uint32_t stage = (uint32_t)(((2 * TR + 1) * (2 * Wseg + 1) * 44 + 127) / 128 * 128);
    const size_t smem = (size_t)((w_bytes + 1023u) & ~1023u) + (size_t)2 * 8 * lbo16 * 16 + (size_t)2 * stage +
                        sizeof(L0fSmem) + 64;
    const int strips = (Ho + TR - 1) / TR;
    const long long items = (long long)p->BK * strips * nseg;
    if (smem > (size_t)227 * 1024 || lbo16 >= 16384u) continue;
    if (items < 2LL * p->num_sms && TR > 1) continue;     // small problems: more, smaller items
    st->f_TR = TR; st->f_Wseg = Wseg; st->f_P = P; st->f_nseg = nseg; st->f_strips = strips;
    st->f_lbo16 = lbo16; st->f_stage = stage; st->f_smem = smem;
    return true;
  }
  return false;
}

int rtc_alloc(Plan* p) {
  RtcState* st = new RtcState();
  p->rtc = st;
  for (int l = 0; l < IODINE_MAX_LAYERS; ++l) { st->w[l] = nullptr; st->w_bytes[l] = 0; st->cin[l] = 0; }
  st->ok = rtc_supported(p) != 0;
  if (!st->ok) return 0;
  st->fused = l0f_geometry(p, st);
  const IodineShape& s = p->s;
  const int Cr = p->Cr, kk = s.ref_k * s.ref_k;
  for (int l = 0; l < s.ref_layers; ++l) {
    st->cin[l] = l == 0 ? 16 : Cr;
    st->w_bytes[l] = (uint32_t)kk * (st->cin[l] / 16) * 2u * (uint32_t)Cr * 16u;
    IOD_CHECK_CUDA(cudaMalloc((void**)&st->w[l], st->w_bytes[l]));
  }
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->tab0, (size_t)p->ref_h[1] * p->ref_w[1] * Cr * sizeof(float)));
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->rowtab, (size_t)p->ref_h[1] * 2 * Cr * sizeof(float)));
  return 0;
}

void rtc_free(Plan* p) {
  RtcState* st = rtc_state(p);
  if (!st) return;
  for (int l = 0; l < IODINE_MAX_LAYERS; ++l) cudaFree(st->w[l]);
  cudaFree(st->tab0);
  cudaFree(st->rowtab);
  delete st;
  p->rtc = nullptr;
}

bool rtc_enabled(const Plan* p) {
  const RtcState* st = reinterpret_cast<const RtcState*>(p->rtc);
  return st && st->ok;
}
bool rtc_fused_aux(const Plan* p) {
  const RtcState* st = reinterpret_cast<const RtcState*>(p->rtc);
  return st && st->ok && st->fused;
}

// weight image [tap][ks][k-half][n][8]; w is OIHW [Cr][CI][k][k], only input channels < cin_real are used
__global__ void rtc_pack_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int Cr, int CI, int cin_real,
                                int cin, int KS, int f16) {
  const int nks = cin / 16;
  const int total = KS * KS * nks * 2 * Cr * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i % 8, n = (i / 8) % Cr, kc = (i / (8 * Cr)) % 2, ks = (i / (16 * Cr)) % nks, tap = i / (16 * Cr * nks);
    const int k = (2 * ks + kc) * 8 + e;
    const float v = k < cin_real ? w[((size_t)n * CI + k) * KS * KS + tap] : 0.f;
    if (f16) { __half h = __float2half_rn(v); img[i] = *reinterpret_cast<uint16_t*>(&h); }
    else { __nv_bfloat16 h = __float2bfloat16(v); img[i] = *reinterpret_cast<uint16_t*>(&h); }
  }
}

// tab0[co/8][y][x][co%8] = bias[co] + zero-padded strided conv of the two coordinate planes (input channels 15, 16 of
// the refinement input: x = linspace(-1,1,W) along W, y along H; iodine.py:334-339)
__global__ void rtc_tab0_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ tab,
                                int Cr, int KS, int S, int H, int W, int Ho, int Wo) {
  const int P = KS / 2;
  const int total = Ho * Wo * Cr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Cr, x = (i / Cr) % Wo, y = i / (Cr * Wo);
    float s = b[co];
    for (int dy = 0; dy < KS; ++dy) {
      const int iy = y * S + dy - P;
      if (iy < 0 || iy >= H) continue;
      const float cyv = (H > 1) ? -1.f + 2.f * (float)iy / (float)(H - 1) : -1.f;
      for (int dx = 0; dx < KS; ++dx) {
        const int ix = x * S + dx - P;
        if (ix < 0 || ix >= W) continue;
        const float cxv = (W > 1) ? -1.f + 2.f * (float)ix / (float)(W - 1) : -1.f;
        s += w[(((size_t)co * 17 + 15) * KS + dy) * KS + dx] * cxv + w[(((size_t)co * 17 + 16) * KS + dy) * KS + dx] * cyv;
      }
    }
    tab[((size_t)(co >> 3) * (Ho * Wo) + (size_t)y * Wo + x) * 8 + (co & 7)] = s;
  }
}

// fused layer 0: rowtab[y][xc][co] = bias[co] + zero-padded stride-2 conv of the y-coordinate plane (input channel 16)
// at output row y; xc = 0: output column 0 (taps dx < pad fall into the padding), xc = 1: any other column (even image
// sizes: no tap leaves the image on the right / bottom).  The x-coordinate plane (channel 15) is a real operand channel.
__global__ void rtc_rowtab_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ tab,
                                  int Cr, int KS, int S, int H, int Ho) {
  const int P = KS / 2;
  const int total = Ho * 2 * Cr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Cr, xc = (i / Cr) % 2, y = i / (2 * Cr);
    float s = b[co];
    for (int dy = 0; dy < KS; ++dy) {
      const int iy = y * S + dy - P;
      if (iy < 0 || iy >= H) continue;
      const float cyv = (H > 1) ? -1.f + 2.f * (float)iy / (float)(H - 1) : -1.f;
      for (int dx = (xc == 0 ? P : 0); dx < KS; ++dx) s += w[(((size_t)co * 17 + 16) * KS + dy) * KS + dx] * cyv;
    }
    tab[i] = s;
  }
}

int rtc_setup_weights(Plan* p, const IodineWeights* w, cudaStream_t st_) {
  RtcState* st = rtc_state(p);
  if (!st || !st->ok) return 0;
  const IodineShape& s = p->s;
  const int Cr = p->Cr, f16 = half_is_f16(p);
  for (int l = 0; l < s.ref_layers; ++l) {
    // layer 0 contracts the 15 data channels; the fused kernel also takes the x-coordinate plane (channel 15) as data
    rtc_pack_kernel<<<32, 256, 0, st_>>>(w->ref_w[l], st->w[l], Cr, l == 0 ? 17 : Cr, l == 0 ? (st->fused ? 16 : 15) : Cr,
                                         st->cin[l], s.ref_k, f16);
    IOD_LAUNCH_CHECK(p);
  }
  if (st->fused) {
    rtc_rowtab_kernel<<<32, 256, 0, st_>>>(w->ref_w[0], w->ref_b[0], st->rowtab, Cr, s.ref_k, s.ref_stride, s.H, p->ref_h[1]);
    IOD_LAUNCH_CHECK(p);
  }
  rtc_tab0_kernel<<<128, 256, 0, st_>>>(w->ref_w[0], w->ref_b[0], st->tab0, Cr, s.ref_k, s.ref_stride, s.H, s.W, p->ref_h[1],
                                        p->ref_w[1]);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

template <int N, int NKS, int TPS, int ST>
static int rtc_launch_t(Plan* p, const RtcParams& q, cudaStream_t st_) {
  auto kern = refine_tc_kernel<N, NKS, TPS, ST>;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  const int grid = q.tiles < p->num_sms ? q.tiles : p->num_sms;
  kern<<<grid, rtc_threads(N), rtc_smem_bytes(q.w_bytes, NKS, TPS, ST), st_>>>(q);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// global average pool of the last layer (F.adaptive_avg_pool2d, iodine.py:481) from the chunk-planar layout
__global__ void rtc_pool_kernel(const uint4* __restrict__ in, float* __restrict__ pool, int HWo, int C, int f16) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x / 32; k < C / 8; k += blockDim.x / 32) {
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x % 32; i < HWo; i += 32) {
      const uint4 v = __ldg(in + ((size_t)n * (C / 8) + k) * HWo + i);
      const float2 a = unpack_h2(v.x, f16), b = unpack_h2(v.y, f16), c = unpack_h2(v.z, f16), d = unpack_h2(v.w, f16);
      s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = warp_sum(s[e]);
    if (threadIdx.x % 32 == 0)
#pragma unroll
      for (int e = 0; e < 8; ++e) pool[(size_t)n * C + k * 8 + e] = s[e] / (float)HWo;
  }
}

template <int N, bool F16>
static int l0f_launch_t(Plan* p, const L0fParams& q, size_t smem, cudaStream_t st_) {
  auto kern = refine_l0f_kernel<N, F16>;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  const int grid = q.items < p->num_sms ? q.items : p->num_sms;
  kern<<<grid, l0f_threads(N), smem, st_>>>(q);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// layer 0 straight from mixture_kernel<FUSED>'s output (p->auxs holds pl0 | raw4 | raw2) -> r16[0]
static int l0f_launch(Plan* p, cudaStream_t st_) {
  RtcState* st = rtc_state(p);
  const IodineShape& s = p->s;
  const size_t nsp = (size_t)p->BK * p->HW;
  L0fParams q;
  q.pl0 = reinterpret_cast<const uint4*>(p->auxs);
  q.raw4 = reinterpret_cast<const float4*>(p->auxs) + nsp;
  q.raw2 = reinterpret_cast<const float2*>(p->auxs) + 4 * nsp;
  q.lik = p->lik;
  q.lnp = p->lnp;
  q.out = reinterpret_cast<uint4*>(p->r16[0]);
  q.wimg = st->w[0];
  q.rowtab = st->rowtab;
  q.w_bytes = st->w_bytes[0];
  q.H = s.H; q.W = s.W; q.Ho = p->ref_h[1]; q.Wo = p->ref_w[1]; q.K = s.K;
  q.TR = st->f_TR; q.Wseg = st->f_Wseg; q.P = st->f_P; q.nseg = st->f_nseg; q.strips = st->f_strips;
  q.items = p->BK * st->f_strips * st->f_nseg;
  q.magicP = (uint32_t)(0x100000000ull / (uint64_t)st->f_P) + 1u;
  q.xc_step = s.W > 1 ? 2.f / (float)(s.W - 1) : 0.f;
  const int f16 = half_is_f16(p);
  q.lbo16 = st->f_lbo16; q.buf16 = 8u * st->f_lbo16; q.stage_bytes = st->f_stage;
  const uint32_t fmt = f16 ? 0u : 1u;
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p->Cr >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#define L0F_CASE(n) if (p->Cr == n) return f16 ? l0f_launch_t<n, true>(p, q, st->f_smem, st_) : l0f_launch_t<n, false>(p, q, st->f_smem, st_);
  L0F_CASE(64) L0F_CASE(32) L0F_CASE(16)
#undef L0F_CASE
  set_error("refine_l0f: unsupported ref_chan=%d", p->Cr);
  return 1;
}

// All refine conv layers on the tensor cores: the aux stack (fused: mixture_kernel's output; otherwise enc16 from
// assemble16, mixture.cu) -> ... -> pool[BK][Cr]
int rtc_launch_refine_convs(Plan* p, cudaStream_t st_) {
  RtcState* st = rtc_state(p);
  const IodineShape& s = p->s;
  const int Cr = p->Cr;
  const void* cur = p->enc16;
  int l_first = 0;
  if (st->fused) {
    if (l0f_launch(p, st_)) return 1;
    cur = p->r16[0];
    l_first = 1;
  }
  for (int l = l_first; l < s.ref_layers; ++l) {
    RtcParams q;
    q.in = reinterpret_cast<const uint4*>(cur);
    q.out = reinterpret_cast<uint4*>(p->r16[l & 1]);
    q.wimg = st->w[l];
    q.w_bytes = st->w_bytes[l];
    q.bias = p->ref_b[l];
    q.tab = l == 0 ? st->tab0 : nullptr;
    q.Hin = p->ref_h[l]; q.Win = p->ref_w[l]; q.Hout = p->ref_h[l + 1]; q.Wout = p->ref_w[l + 1];
    q.S = s.ref_stride; q.pad = s.ref_k / 2; q.KS = s.ref_k;
    q.cin_planes = st->cin[l] / 8;
    q.total = p->BK * q.Hout * q.Wout;
    q.tiles = (q.total + 127) / 128;
    q.f16 = half_is_f16(p);
    const uint32_t fmt = q.f16 ? 0u : 1u;
    q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(Cr >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int nks = st->cin[l] / 16;
    int rc = 1;
    set_error("refine_tc: unsupported ref_chan=%d / K-steps %d", Cr, nks);
#define RTC_CASE(n, ksz, k, tps, stg) if (Cr == n && s.ref_k == ksz && nks == k) rc = rtc_launch_t<n, k, tps, stg>(p, q, st_);
    RTC_CASE(64, 3, 1, 9, 4) RTC_CASE(32, 3, 1, 9, 4) RTC_CASE(16, 3, 1, 9, 4)
    RTC_CASE(64, 3, 4, 3, 3) RTC_CASE(32, 3, 2, 3, 6)
    RTC_CASE(64, 5, 1, 5, 6) RTC_CASE(32, 5, 1, 5, 6) RTC_CASE(16, 5, 1, 5, 6)
    RTC_CASE(32, 5, 2, 1, 12)
#undef RTC_CASE
    if (rc) return 1;
    cur = p->r16[l & 1];
  }
  const int HWo = p->ref_h[s.ref_layers] * p->ref_w[s.ref_layers];
  rtc_pool_kernel<<<p->BK, 256, 0, st_>>>(reinterpret_cast<const uint4*>(cur), p->pool, HWo, Cr,
                                          half_is_f16(p));
  IOD_LAUNCH_CHECK(p);
  return 0;
}

}  // namespace iod
