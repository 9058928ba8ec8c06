// wgrad_tc.cu -- tcgen05 weight gradient of the decoder's C -> C 3x3 layers (training step, csrc/train.cu).
//
//   dW[co][ci][dy][dx] += coef * sum_{n,y,x} g[n,co,y,x] * act[n,ci,y+dy-1,x+dx-1]
//
// (torch.nn.grad.conv2d_weight as autograd runs it for nn.Conv2d, reference lib/modeling/iodine.py:583, under
// loss.backward(), lib/engine/train.py:63).  A GEMM whose contraction runs over PIXELS: 270 GFLOP per layer and ELBO
// evaluation at the CLEVR6 sizes, as much as the layer's forward pass.
//
// Both operands are read exactly as they lie in HBM, chunk-planar [slot-image][C/8][H][W][8 x 16 bit]: eight
// consecutive pixels of one plane are 128 contiguous bytes = one core matrix of an MN-MAJOR (channel-contiguous),
// no-swizzle UMMA operand -- 8 K rows (pixels) of 16 bytes (8 channels).  LBO = 128 B steps to the next eight pixels,
// SBO = the plane pitch steps to the next eight channels.  No-swizzle descriptors only need 16-byte aligned starts,
// so the horizontal tap shift dx is a +16 B on the activation operand's start address: no im2col, no transpose.
//   * the ring keeps image rows as [row][plane][136 pixels]; the plane pitch is uniform ACROSS the row boundary
//     (plane 8 of row j = plane 0 of row j+1), so ONE descriptor spans two consecutive rows:
//     A = gradient rows (y, y+1) x 64 channels -> M = 128; B = activation rows (y, y+1) x 64 channels -> N = 128.
//     D[(r,co)][(s,ci)] then holds four tap rows at once: (r,s) = (0,0) dy=1, (0,1) dy=2, (1,0) dy=0 and (1,1)
//     dy=1 again (discarded), so every (gradient row, dy) pair is produced exactly once by the sliding row pairs;
//   * per row pair: 3 (dx) x 8 (K = 16 pixels) tcgen05.mma M=128 N=128 into three persistent TMEM accumulators
//     (384 columns) that live for the whole kernel; one epilogue per CTA adds them to the fp32 gradient;
//   * four producer warps stream the rows with 2 KB bulk copies (one plane row each), mbarrier pipelined.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace iod {

constexpr int WG_R = 5;                    // ring rows; slot WG_R mirrors slot 0 so that the pair (R-1, 0) is contiguous
constexpr int WG_PS = 136;                 // pixels per plane row in the ring: 128 + zero halo (data at 1..128)
constexpr int WG_PLANES = 8;               // C = 64 channels
constexpr int WG_ROW16 = WG_PLANES * WG_PS;            // 16-byte units per ring row
constexpr int WG_RING_BYTES = (WG_R + 1) * WG_ROW16 * 16;
constexpr int WG_THREADS = 32 * 9;         // 4 producers, 1 issuer, 4 epilogue warps

struct WgtParams {
  const uint4* g;                // dJ/d(pre-activation l), chunk-planar
  const uint4* act;              // activation l-1, chunk-planar
  const uint4* zero_row;         // >= 2 KB of zeros
  float* dw;                     // [64][64][3][3] fp32 (PyTorch OIHW), accumulated with atomics
  const int4* itab;              // work items {slot-image, first row, rows, -}
  const int32_t* coff;           // [grid + 1] item range of CTA c
  float coef;
  int32_t H;
  uint32_t idesc;
};

struct WgtSmem {
  uint64_t full[WG_R];
  uint64_t empty[WG_R];
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgtParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_g = smem;
  uint8_t* s_a = smem + WG_RING_BYTES;
  WgtSmem* sb = reinterpret_cast<WgtSmem*>(smem + 2 * WG_RING_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  constexpr int W = 128;

  {  // zero both rings once: halo columns stay zero for the whole kernel
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 2 * WG_RING_BYTES / 16; i += WG_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_R; ++i) {
      mbar_init(smem_u32(&sb->full[i]), 4);        // four producer warps
      mbar_init(smem_u32(&sb->empty[i]), 1);
    }
    mbar_init(smem_u32(&sb->done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sb->tmem_base;
  const int it0 = __ldg(p.coff + cta), it1 = __ldg(p.coff + cta + 1);

  if (warp < 4) {
    // =============================================================== producers
    // warp 0 / 1: gradient planes 0-3 / 4-7; warp 2 / 3: activation planes 0-3 / 4-7
    if (elect_one_sync()) {
      const bool is_act = warp >= 2;
      const int pl0 = (warp & 1) * 4;
      const uint4* src_t = is_act ? p.act : p.g;
      const uint32_t ring = smem_u32(is_act ? s_a : s_g);
      const size_t plane_px = (size_t)p.H * W;
      int jg = 0;                                   // ring rows since the kernel started
      for (int item = it0; item < it1; ++item) {
        const int4 d = __ldg(p.itab + item);
        const int n = d.x, y0 = d.y, th = d.z;
        for (int j = 0; j < th + 2; ++j, ++jg) {
          const int slot = jg % WG_R;
          const uint32_t ph = (uint32_t)(jg / WG_R) & 1u;
          mbar_wait(smem_u32(&sb->empty[slot]), ph ^ 1u, 1);
          const bool mirror = slot == 0;
          const uint32_t fb = smem_u32(&sb->full[slot]);
          mbar_expect_tx(fb, 4u * 2048u * (mirror ? 2u : 1u));
          const int y = y0 - 1 + j;
          // gradient rows outside [y0, y0+th) belong to another work item (or lie outside the image): zeros, so
          // that every (row, dy) product is counted once; activation rows are real wherever the image has them
          const bool real = is_act ? (y >= 0 && y < p.H) : (j >= 1 && j <= th);
          const uint4* src = real ? src_t + ((size_t)n * WG_PLANES + pl0) * plane_px + (size_t)y * W : p.zero_row;
          const size_t sstep = real ? plane_px : 0;
          uint32_t dst = ring + (uint32_t)(slot * WG_ROW16 + pl0 * WG_PS + 1) * 16u;
          for (int c = 0; c < 4; ++c) {
            bulk_load_1d(dst, src, 2048u, fb);
            if (mirror) bulk_load_1d(dst + (uint32_t)(WG_R * WG_ROW16) * 16u, src, 2048u, fb);
            src += sstep;
            dst += (uint32_t)WG_PS * 16u;
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
    const uint32_t g16 = smem_u32(s_g) >> 4, a16 = smem_u32(s_a) >> 4;
    int jg = 0;
    bool first = true;
    for (int item = it0; item < it1; ++item) {
      const int th = __ldg(p.itab + item).z;
      for (int q = 0; q <= th; ++q) {               // pair q = ring rows (q, q+1) of this item
        if (q == 0) mbar_wait(smem_u32(&sb->full[jg % WG_R]), (uint32_t)(jg / WG_R) & 1u, 2);
        mbar_wait(smem_u32(&sb->full[(jg + q + 1) % WG_R]), (uint32_t)((jg + q + 1) / WG_R) & 1u, 3);
        tc_fence_after();
        const int sa = (jg + q) % WG_R;             // (slot sa + 1 == WG_R is the mirror of slot 0)
        const uint32_t ab = (g16 + (uint32_t)(sa * WG_ROW16 + 1)) | LBO;
        const uint32_t bb = (a16 + (uint32_t)(sa * WG_ROW16)) | LBO;
        if (leader) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              tc_mma<false>(tmem_base + (uint32_t)(dx * 128), DESC_HI | (uint64_t)(ab + (uint32_t)(16 * ks)),
                            DESC_HI | (uint64_t)(bb + (uint32_t)(dx + 16 * ks)), p.idesc, (first && ks == 0) ? 0u : 1u);
            }
          }
          tc_commit(smem_u32(&sb->empty[sa]));                                   // row q is not read again
          if (q == th) tc_commit(smem_u32(&sb->empty[(jg + q + 1) % WG_R]));      // nor is the item's last row
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
      const int m = quad * 32 + lane, r = m >> 6, co = m & 63;
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          const int s = c >> 2;                      // activation row of the pair
          uint32_t acc[16];
          IOD_TMEM_LD16(acc, tmem_base + (uint32_t)(dx * 128 + c * 16) + ((uint32_t)(quad * 32) << 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (r == 1 && s == 1) continue;            // the duplicate dy = 1 quadrant
          const int dy = (r == 0) ? 1 + s : 0;
          const int ci0 = (c & 3) * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            atomicAdd(p.dw + ((size_t)(co * 64 + ci0 + i) * 9 + dy * 3 + dx), p.coef * __uint_as_float(acc[i]));
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// 1 when the tensor-core weight gradient can run this plan's C -> C layers: 16-bit operand modes, C = 64, 3x3, W = 128
int wgrad_tc_supported(const Plan* p) {
  if (!tc_mode(p) || tf_mode(p) || getenv("IODINE_WGRAD_FFMA")) return 0;
  return p->C == 64 && p->s.dec_k == 3 && p->s.W == 128 && tc_rs_worklist(p, nullptr, nullptr, nullptr, nullptr);
}

int launch_wgrad_tc(Plan* p, const void* act_prev, const void* g, float* dw, float coef, cudaStream_t st) {
  WgtParams q;
  int grid = 0;
  const void* zero = nullptr;
  IOD_REQUIRE(tc_rs_worklist(p, &q.itab, &q.coff, &grid, &zero), "wgrad_tc: no row work list for this plan");
  q.g = reinterpret_cast<const uint4*>(g);
  q.act = reinterpret_cast<const uint4*>(act_prev);
  q.zero_row = reinterpret_cast<const uint4*>(zero);
  q.dw = dw;
  q.coef = coef;
  q.H = p->s.H;
  const uint32_t fmt = p->s.precision == IODINE_FP16 ? 0u : 1u;
  // cute::UMMA::InstrDescriptor: f32 accumulate, a/b format, A and B MN-major (bits 15, 16), N = 128, M = 128
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t smem = 2 * (size_t)WG_RING_BYTES + sizeof(WgtSmem) + 64;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(q);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

}  // namespace iod
