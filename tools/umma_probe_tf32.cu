// umma_probe_tf32.cu -- how many SM cycles does one tcgen05.mma kind::tf32 (M=128, N, K=8) cost from no-swizzle K-major
// shared-memory operands, issued back to back by one thread (the row-streaming pattern of conv_tc.cu: 3 dx x 8 K-steps
// per ring row), next to kind::f16 (K=16) with the same bytes per operand?  Timing only: operand contents are zeros.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe_tf32 tools/umma_probe_tf32.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool TF>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (TF)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}

// rows x (3 x NKS) MMAs; A: ring rows of pitch 136 positions, planes 16 B per position (8 f16 or 4 tf32 channels),
// one K-step = two planes; B: [entry][k-half][N][16 B]
template <int N, bool TF, int NKS>
__global__ void __launch_bounds__(128, 1) rows_kernel(int rows, int commits, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (220 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t fmt = TF ? 2u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint64_t HI = (uint64_t)(8u | (1u << 14)) << 32;
  const uint32_t Ps = 136, R = 4, ps16 = R * Ps;                    // plane stride: R ring rows
  const uint32_t w_bytes = 3u * NKS * 2u * N * 16u;
  const uint32_t a_base16 = smem_u32(smem + ((w_bytes + 1023) & ~1023u)) >> 4;
  const uint32_t w_base16 = smem_u32(smem) >> 4;
  if (tid == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < rows; ++r) {
      const uint32_t rb = (a_base16 + (uint32_t)(r % R) * Ps) | (ps16 << 16);
      const uint32_t wb = w_base16 | ((uint32_t)N << 16);
      const uint32_t d = tmem_base + (uint32_t)((r & 1) * 256);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          const int e = dx * NKS + ks;
          mma<TF>(d, HI | (rb + (uint32_t)dx + (uint32_t)(2 * ks) * ps16), HI | (wb + (uint32_t)(e * 2 * N)), idesc, 1u);
        }
      for (int c = 0; c < commits; ++c)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[(r * 2 + c) & 15])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&done)) : "memory");
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int N, bool TF, int NKS>
static void run(long long* d_cyc, int grid) {
  CK(cudaFuncSetAttribute(rows_kernel<N, TF, NKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  std::vector<long long> cyc(148);
  const int rows = 1000;
  for (int commits = 0; commits <= 2; commits += 2) {
    rows_kernel<N, TF, NKS><<<grid, 128, 220 * 1024>>>(rows, commits, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    printf("%s N=%3d K-steps/tap=%d grid=%3d commits/row=%d : %6.1f cycles/MMA, %6.0f cycles/row of %d MMAs, %.0f FLOP/clk/SM\n", TF ? "tf32" : "f16 ", N, NKS,
           grid, commits, (double)mx / (rows * 3.0 * NKS), (double)mx / rows, 3 * NKS,
           2.0 * 128 * N * (TF ? 8 : 16) * 3 * NKS * rows / (double)mx);
  }
}

int main() {
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 148 * sizeof(long long)));
  for (int grid : {1, 148}) {
    run<192, false, 4>(d_cyc, grid);     // the fp16 production row: 12 MMAs N=192 K=16
    run<96, true, 8>(d_cyc, grid);       // the tf32 production row (per CTA of a pair): 24 MMAs N=96 K=8
    run<192, true, 4>(d_cyc, grid);      // tf32 with half of the input channels (K-split idea): 12 MMAs N=192 K=8
    run<192, true, 8>(d_cyc, grid);      // tf32 with all weights resident: 24 MMAs N=192 K=8
    run<256, true, 4>(d_cyc, grid);
    run<64, true, 8>(d_cyc, grid);
    run<96, false, 4>(d_cyc, grid);
  }
  return 0;
}
