// umma_probe.cu -- micro-benchmark + correctness probe for tcgen05.mma shared-memory operand layouts.
//
// Question it answers ("why the decoder ring is laid out the way it is"): how many SM
// cycles does one tcgen05.mma (M=128, N, K=16, f16 -> f32) cost when the A operand is
//   mode 0: the no-swizzle "core matrix" layout (8 rows x 16 B contiguous, LBO between the two K halves)
//           with an aligned start, or a start shifted by one row (16 B) -- the shifted-tap views of the
//           flat-ring implicit GEMM;
//   mode 1: the 128-byte-swizzled K-major layout (one 128 B row per pixel) with the start shifted by whole
//           rows (128 B) and the descriptor's base_offset field set accordingly.
// and whether the row-shifted swizzled view produces the right numbers.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
// run  : tools/umma_probe            (prints one line per configuration)
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int AROWS = 128 + 24;     // logical A rows available (so that row shifts stay in bounds)
constexpr int KTOT = 64;            // 4 K-steps of 16
constexpr int NMAX = 256;

struct Params {
  int mode;        // 0 no-swizzle planar, 1 SW128 K-major
  int N;
  int shift;       // A row shift
  int bo_mode;     // mode 1: 0 -> base_offset 0, 1 -> base_offset = (addr >> 7) & 7
  int reps;        // MMA groups (each group = 4 K-steps)
  int check;       // write D
  int aplane;      // mode 0: bytes between 8-channel planes of A (the descriptor's LBO)
  int commit_every; // timing runs: tcgen05.commit (to an unobserved mbarrier) after every n MMA groups (0 = never)
  int stages;      // timing runs: rotate the accumulator over this many 64-column TMEM stages per commit group
};

__device__ __forceinline__ float a_val(int r, int k) { return (float)(((r * 7 + k * 3) % 13) - 6); }
__device__ __forceinline__ float b_val(int n, int k) { return (float)(((n * 5 + k * 11) % 9) - 4); }

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p, long long* cycles, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                          // 64 KB region
  uint8_t* sB = smem + 163840;                 // after the A region (8 planes x <= 20 KB)
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- fill operands (generic proxy)
  const int APLANE = p.aplane;                 // mode 0: bytes per 8-channel plane
  for (int i = tid; i < AROWS * KTOT; i += 128) {
    const int r = i / KTOT, k = i % KTOT;
    size_t off;
    if (p.mode == 0) off = (size_t)(k / 8) * APLANE + (size_t)r * 16 + (k % 8) * 2;
    else off = (size_t)r * 128 + (size_t)(((k / 8) ^ (r & 7)) * 16) + (k % 8) * 2;
    *reinterpret_cast<__half*>(sA + off) = __float2half(a_val(r, k));
  }
  for (int i = tid; i < p.N * KTOT; i += 128) {
    const int n = i / KTOT, k = i % KTOT;
    size_t off;
    if (p.mode == 0) off = (size_t)(k / 8) * (p.N * 16) + (size_t)n * 16 + (k % 8) * 2;   // [k8][n][8]
    else off = (size_t)n * 128 + (size_t)(((k / 8) ^ (n & 7)) * 16) + (k % 8) * 2;
    *reinterpret_cast<__half*>(sB + off) = __float2half(b_val(n, k));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (tid == 0) {
    uint64_t adesc[4], bdesc[4];
    for (int ks = 0; ks < 4; ++ks) {
      if (p.mode == 0) {
        const uint32_t a = smem_u32(sA) + (uint32_t)(2 * ks) * APLANE + (uint32_t)p.shift * 16;
        const uint32_t b = smem_u32(sB) + (uint32_t)(2 * ks) * (p.N * 16);
        adesc[ks] = (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((APLANE >> 4) & 0x3FFF) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
        bdesc[ks] = (uint64_t)((b >> 4) & 0x3FFF) | ((uint64_t)(p.N & 0x3FFF) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
      } else {
        const uint32_t a = smem_u32(sA) + (uint32_t)p.shift * 128 + (uint32_t)ks * 32;
        const uint32_t b = smem_u32(sB) + (uint32_t)ks * 32;
        const uint64_t bo = p.bo_mode ? (uint64_t)((a >> 7) & 7) : 0ull;
        adesc[ks] = (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | (1ull << 46) | (bo << 49) | (2ull << 61);
        bdesc[ks] = (uint64_t)((b >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | (1ull << 46) | (2ull << 61);
      }
    }
    const long long t0 = clock64();
    for (int r = 0; r < p.reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t acc = (r > 0 || ks > 0) ? 1u : 0u;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
            ::"r"(tmem_base + (uint32_t)(p.stages > 1 ? ((r / (p.commit_every ? p.commit_every : 1)) % p.stages) * 64 : 0)), "l"(adesc[ks]), "l"(bdesc[ks]), "r"(idesc), "r"(acc));
      }
      if (p.commit_every && (r + 1) % p.commit_every == 0)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[(r / p.commit_every) & 3])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    // wait for completion
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    const long long t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (p.check && blockIdx.x == 0) {
    for (int c0 = 0; c0 < p.N; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_base + (uint32_t)c0 + ((uint32_t)(warp * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[(size_t)tid * NMAX + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------
// tile_kernel: the production issue pattern of conv_tc.cu (36 distinct descriptor pairs per tile:
// 3 row bases x 3 dx x 4 K-steps against 36 weight tiles, N = 64) with the things one can vary:
// commits per tile, accumulator-stage rotation, number of issuing warps.  Operand contents are
// whatever is in shared memory (timing only).
// ------------------------------------------------------------------------------------------------
struct TileParams { int tiles; int commits; int rotate; int issuers; int N; int rs; };

__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}

template <int N>
__global__ void __launch_bounds__(128, 1) tile_kernel(TileParams p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  __shared__ uint64_t done[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (220 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint64_t HI = (uint64_t)(8u | (1u << 14)) << 32;
  const uint32_t Ps = 136, ps16 = 8 * 136;          // ring pitch, plane stride (16-byte units): 8 rows
  const uint32_t a_base16 = smem_u32(smem + 76 * 1024) >> 4;
  const uint32_t w_base16 = smem_u32(smem) >> 4;
  long long t0 = 0;
  if (warp < p.issuers) {
    t0 = clock64();
    for (int t = warp; t < p.tiles; t += p.issuers) {
      const uint32_t stage = p.rotate ? (uint32_t)(t & 3) : (uint32_t)warp;
      const uint32_t d_tmem = __shfl_sync(0xffffffffu, tmem_base + stage * 64u, 0);
      const uint32_t wb = __shfl_sync(0xffffffffu, w_base16, 0) | ((uint32_t)N << 16);
      uint32_t rb[3];
      int pr = t % 5;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        rb[dy] = __shfl_sync(0xffffffffu, (a_base16 + (uint32_t)pr * Ps) | (ps16 << 16), 0);
        if (++pr == 5) pr = 0;
      }
      if (lane == 0 && p.rs) {                     // row-streaming pattern: 12 MMAs (dx, ks) on ONE ring row
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int e = dx * 4 + ks;
            mma_f16(tmem_base, HI | (rb[0] + (uint32_t)dx + (uint32_t)(2 * ks) * ps16), HI | (wb + (uint32_t)(e * 2 * N)), idesc, 1u);
          }
        for (int c = 0; c < p.commits; ++c)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[(t * 2 + c) & 15])) : "memory");
      } else if (lane == 0) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const int e = (dy * 3 + dx) * 4 + ks;
              mma_f16(d_tmem, HI | (rb[dy] + (uint32_t)dx + (uint32_t)(2 * ks) * ps16), HI | (wb + (uint32_t)(e * 2 * N)), idesc,
                      e > 0 ? 1u : 0u);
            }
        for (int c = 0; c < p.commits; ++c)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[(t * 2 + c) & 15])) : "memory");
      }
      __syncwarp();
    }
    if (lane == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done[warp])) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&done[warp])) : "memory");
      cycles[blockIdx.x * 4 + warp] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

template <int N>
static void run_rows(long long* d_cyc) {
  CK(cudaFuncSetAttribute(tile_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  std::vector<long long> cyc(148 * 4);
  for (int commits = 0; commits <= 2; commits += 2) {
    TileParams p{1200, commits, 0, 1, N, 1};
    tile_kernel<N><<<148, 128, 220 * 1024>>>(p, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = cyc[i * 4] > mx ? cyc[i * 4] : mx;
    printf("rows  N=%d commits/row=%d : %.1f cycles/MMA (%.0f cycles/row of 12 MMAs)\n", N, commits, (double)mx / (1200.0 * 12),
           (double)mx / 1200.0);
  }
}

template <int N>
static void run_tiles(long long* d_cyc) {
  CK(cudaFuncSetAttribute(tile_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  std::vector<long long> cyc(148 * 4);
  for (int issuers = 1; issuers <= 4; issuers *= 2)
    for (int rotate = 0; rotate <= 1; ++rotate)
      for (int commits = 0; commits <= 2; ++commits) {
        TileParams p{400, commits, rotate, issuers, N, 0};
        tile_kernel<N><<<148, 128, 220 * 1024>>>(p, d_cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < 148; ++i)
          for (int w = 0; w < issuers; ++w) mx = cyc[i * 4 + w] > mx ? cyc[i * 4 + w] : mx;
        printf("tiles N=%d issuers=%d rotate=%d commits/tile=%d : %.1f cycles/MMA (%.0f cycles/tile)\n", N, issuers, rotate, commits,
               (double)mx / (400.0 * 36), (double)mx / 400.0);
      }
}

static float ha(int r, int k) { return (float)(((r * 7 + k * 3) % 13) - 6); }
static float hb(int n, int k) { return (float)(((n * 5 + k * 11) % 9) - 4); }

int main() {
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840 + 32768 + 1024));
  long long* d_cyc; float* d_D;
  CK(cudaMalloc(&d_cyc, 148 * 4 * sizeof(long long)));
  run_rows<192>(d_cyc);
  run_rows<128>(d_cyc);
  run_rows<96>(d_cyc);
  run_rows<64>(d_cyc);
  run_rows<48>(d_cyc);
  run_rows<16>(d_cyc);
  if (getenv("PROBE_ROWS_ONLY")) return 0;
  run_tiles<64>(d_cyc);
  run_tiles<16>(d_cyc);
  if (getenv("PROBE_TILES_ONLY")) return 0;
  CK(cudaMalloc(&d_D, 128 * NMAX * sizeof(float)));
  std::vector<float> D(128 * NMAX);
  std::vector<long long> cyc(148);
  printf("mode N shift bo aplane | correct | cycles/MMA (1 CTA) | cycles/MMA (148 CTAs, max)\n");
  const int SM = 163840 + 32768 + 1024;
  const int Ns[] = {16, 64, 128};
  struct Cfg { int mode, shift, bo, aplane, ce, st; } cfgs[] = {
      {0, 1, 0, 17408, 0, 1}, {0, 1, 0, 17408, 9, 1}, {0, 1, 0, 17408, 9, 4}, {0, 1, 0, 17408, 3, 4}, {0, 1, 0, 17408, 1, 4},
      {0, 1, 0, 17408, 1, 1}};
  for (const Cfg& c : cfgs) {
    for (int N : Ns) {
      Params p{c.mode, N, c.shift, c.bo, 1, 1, c.aplane, 0, 1};
      CK(cudaMemset(d_D, 0, 128 * NMAX * sizeof(float)));
      probe_kernel<<<1, 128, SM>>>(p, d_cyc, d_D);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), d_D, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0.f;
          for (int k = 0; k < KTOT; ++k) ref += ha(r + c.shift, k) * hb(n, k);
          if (D[(size_t)r * NMAX + n] != ref) ++bad;
        }
      const int reps = 1800;
      Params q{c.mode, N, c.shift, c.bo, reps, 0, c.aplane, c.ce, c.st};
      probe_kernel<<<1, 128, SM>>>(q, d_cyc, d_D);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
      const double one = (double)cyc[0] / (reps * 4);
      probe_kernel<<<148, 128, SM>>>(q, d_cyc, d_D);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(cyc.data(), d_cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      printf("%d %3d %2d %d %5d ce=%d st=%d | %s (%d bad) | %7.1f | %7.1f\n", c.mode, N, c.shift, c.bo, c.aplane, c.ce, c.st, bad ? "WRONG" : "ok", bad, one,
             (double)mx / (reps * 4));
    }
  }
  return 0;
}
