// conv_tc.cu -- tcgen05 (5th-gen tensor core) decoder convolutions for sm_100a.
//
// Replaces, in the IODINE_BF16 precision mode, the same reference code as conv_f32.cu:
// MultiLayerConv.forward layers >= 1 (reference lib/modeling/iodine.py:586-594), Decoder.conv
// (iodine.py:422,435) and the data-gradient half of the autograd convolution_backward that
// (B*elbo).backward() (iodine.py:90) runs through the decoder.
//
// Formulation: im2col-free implicit GEMM on a FLATTENED, zero-padded image.
//   * activations live in HBM "chunk-planar": [slot-image][C/8][H][W][8] bf16, i.e. one plane
//     per group of 8 channels, 16 bytes per pixel per plane;
//   * a TMA producer streams halo rows (image row y, columns -pad .. W+pad-1, zero filled by
//     the TMA unit outside the image) of every plane into a shared-memory RING of rows with
//     pitch Ps (a multiple of 8 pixels).  In ring space pixel (row j, col c) sits at flat
//     position j*Ps + c, so the 3x3 (5x5) tap (dy,dx) of ANY run of 128 consecutive output
//     positions is the run of 128 consecutive ring positions shifted by dy*Ps + dx;
//   * that run is exactly a K-major, no-swizzle UMMA operand: 8 consecutive positions x 16 B
//     form one core matrix (128 contiguous bytes), SBO = 128 B between groups of 8 positions,
//     LBO = the plane stride between the two 8-channel halves of one K=16 step.  No-swizzle
//     descriptors only need a 16-byte aligned start address, which is what makes the shifted
//     views legal without any im2col copy;
//   * one elected thread issues tcgen05.mma (M=128 positions, N=Cout, K=16) for every
//     (tap, 16-channel step) into a TMEM accumulator; 4 accumulator stages let the epilogue
//     warps (tcgen05.ld -> bias/ELU or ELU' -> bf16 -> coalesced 16-byte stores, straight from
//     registers into the chunk-planar output) overlap the next tile's MMAs;
//   * ring slots wrap: the first `m` slots are mirrored behind the ring end so that a 128-run
//     that straddles the wrap point is still contiguous.
// The 4-channel ends of the decoder use the same kernel: decoder.conv (C -> 4) runs with
// N = 16 (the smallest N of an M=128 UMMA; 12 zero columns), its data-gradient (4 -> C) packs TWO taps into one K=16 step by
// pointing LBO at the flat distance between the taps (one 8-channel plane, 4 real channels).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include <type_traits>
#include "tc_ptx.cuh"

namespace iod {

// ------------------------------------------------------------------------------------------------
// parameters of one launch
// ------------------------------------------------------------------------------------------------
constexpr int TC_MAX_RING = 32;
// threads: warp 0 TMA, warps 1..TC_ISSUERS MMA issuers (warp 1 also owns TMEM), then 4 or 8 epilogue
// warps (two per TMEM lane quadrant, each taking half of the accumulator columns, when N >= 32).
// Several issuers because ONE thread cannot feed the tensor pipe on these shapes: a tile is 36
// distinct (A,B) descriptor pairs, ~7 issue cycles per uniform-datapath instruction and ~6
// instructions per MMA measured, against 48 cycles of tensor-pipe time per M=128,N=64,K=16 MMA
// (tools/umma_probe.cu).  Issuer w takes tiles w, w+TC_ISSUERS, ... of the CTA's tile sequence; a
// tile owns one accumulator stage, so issuers never share an accumulator.
constexpr int TC_ISSUERS = 2;
__host__ __device__ constexpr int tc_epi_warps(int N) { return N >= 32 ? 8 : 4; }
// row-streaming kernels carry more warps behind the epilogue warps: the scout and, where the register file allows, a
// second MMA issuer (see the issuer).  A 13th warp caps the block at 128 registers per thread (allocation is per four
// warps), which the forward epilogues fit and the data-gradient ones (140-168: prefetch queue of the saved
// activation, running class sums) do not: those keep the single issuer.
__host__ __device__ constexpr bool tc_iss2(int epi, bool rs, bool tf) { (void)tf; return rs && (epi == 0 || epi == 2); }
__host__ __device__ constexpr int tc_threads(int N, bool rs = false, bool iss2 = false) {
  return 32 * (1 + TC_ISSUERS) + 32 * tc_epi_warps(N) + (rs ? 32 : 0) + (iss2 ? 32 : 0);
}
constexpr int TC_ACC_STAGES = 4;        // accumulator stages of the generic kernels (8 measured no faster: profiles/r2_issuer_experiments.md)
constexpr int TC_RS_UNIT = 5;        // input rows per issue unit of the row-streaming variant (<= 16; TcParams::rs_unit < ring rows)
constexpr int TC_RS_SLOTS = 8;       // accumulator slots of the row-streaming variant (one per output row in flight) ...
constexpr int TC_RS_SLOTS_MAX = 16;  // ... 16 where they fit the 512 TMEM columns (N <= 32): the KS-slot window of a row then
                                     // straddles the end of the slot ring -- two MMAs per step instead of one -- half as often

// EPI_DSUM: data-gradient whose output is not stored but reduced over the border classes of the collapsed first
// layer straight from the accumulators (row-streaming kernels only; replaces EPI_DGRAD + tc_class_sum_kernel)
enum TcEpi { EPI_FWD = 0, EPI_DGRAD = 1, EPI_OUT4 = 2, EPI_DSUM = 3 };

struct TcMaps {                   // one source buffer: rows are fetched as runs of 128 pixels
  CUtensorMap full;               // 4-D {2W (8-byte elements), H, planes, BK}, box 256 elements
  CUtensorMap tail;               // same tensor, box = the remainder of the halo row
};

struct alignas(64) TcParams {
  TcMaps maps;
  int32_t f16;                    // 1: IEEE half operands, 0: bfloat16
  int32_t dbg;                    // IODINE_TC_DEBUG bit mask (timing experiments only; results are wrong when set):
                                  // 1 no TMEM reads, 2 no epilogue stores, 4 no TMA (generic producer), 8 no activation loads,
                                  // 16 no input rows (row-streaming producers)
  int32_t nch_in;                 // input planes
  int32_t nch_out;                // output planes of the whole layer (all channel splits)
  int32_t n_tot;                  // output channels of the whole layer
  int32_t nsplit;                 // CTAs sharing one work range, each computing N of the n_tot output channels
                                  // (grid = nsplit x ranges; CTA c: range c / nsplit, channel block c % nsplit).
                                  // tf32 C=64: the 147 KB weight image does not fit beside a ring, two halves do
  int32_t H, W;
  int32_t pad;
  int32_t Ps;                     // ring row pitch, positions
  int32_t R, m;                   // ring rows, mirrored rows
  int32_t segs;                   // W/128 when tiles are row aligned, 0 = flat tiling
  int32_t TH;                     // image rows per work item (divides H: every item is TH rows)
  int32_t NT;                     // tiles per work item
  int32_t strips;                 // ceil(H / TH)
  int32_t items;                  // BK * strips (row-streaming: * rs_segs)
  int32_t rs_segs;                // row-streaming: 128-column segments per image row (an item is one of them)
  int32_t rs_unit;                // row-streaming: input rows per issue unit (all waits first, then the MMAs), < R
  // row-streaming work list: every CTA owns ONE contiguous range of the flattened (slot-image, segment, row)
  // space, cut only at image boundaries -- items of up to H rows instead of TH-row strips: rows per CTA are
  // balanced to +-1 (strips: 25 vs 24 per CTA at B = 32) and a CTA restarts the row stream 2-3 times per launch
  // instead of 25 (each restart = KS-1 halo rows re-read and 2(KS-1) short, issue-bound MMA passes)
  const int4* itab;               // {slot-image n, first row y0, rows th, first column xoff}
  const int32_t* coff;            // [grid + 1] item range of CTA c
  int32_t rev;                    // walk the work items from the last to the first: consecutive layers alternate
                                  // direction so that a layer starts on what the previous one wrote last (still in L2)
  uint32_t idesc;
  uint32_t idesc_n[8];            // row-streaming: instruction descriptor for c accumulator blocks (N = c * Cout)
  uint32_t w_bytes;               // weight image bytes (multiple of 16)
  uint32_t box_bytes;             // bytes of one halo row of one plane ((W + 2 pad) * 16)
  int32_t n_full;                 // full 128-pixel boxes per halo row
  int32_t tail_px;                // pixels of the tail box (0 = none)
  uint32_t plane_stride16;        // ring plane stride, 16-byte units
  uint32_t a_off[40];             // A-descriptor offset of (dx, K-step): dx + 2*ks*plane_stride16 (read as constants)
  const void* wimg;               // packed bf16 weights (global), layout = smem image
  const uint4* in;                // row-streaming: chunk-planar source activation (1-D bulk row copies)
  const uint4* zero_row;          // row-streaming: W x 16 zero bytes (rows outside the image)
  const float* bias;              // [N] (EPI_FWD, EPI_OUT4)
  const uint4* actp;              // chunk-planar previous activation (EPI_DGRAD, EPI_DSUM)
  float* G;                       // [n][3*3][N] class sums (EPI_DSUM), accumulated with atomics
  int32_t apf;                    // row-streaming data-gradients: producers prefetch the saved activation into L2
  int32_t lead;                   // row-streaming producers: at most this many ring rows ahead of the consumer (0: ring depth)
  void* out;                      // chunk-planar bf16 (uint4 per position-plane) or fp32 out4
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct TcSmem {                    // tail of the dynamic shared memory block
  uint64_t full[TC_MAX_RING];
  uint64_t empty[TC_MAX_RING];
  uint64_t tfull[TC_RS_SLOTS_MAX];
  uint64_t tempty[TC_RS_SLOTS_MAX];
  uint64_t wbar;
  // row-streaming: the scout hands one issue unit of row plans at a time to the issuer(s); double-buffered, guarded by
  // named barriers (3 + pp / 7 + pp: plans of buffer pp ready for issuer 0 / 1; 5 + pp: buffer pp consumed)
  uint4 plan[2][16];               // per row of the unit: {A base | LBO, first accumulator, packed counts, rows of the unit | last-unit flag}
  uint32_t tmem_base;
};

__device__ __forceinline__ int tc_num_tiles(const TcParams& p, int th) {
  return p.segs ? th * p.segs : ((th - 1) * p.Ps + p.W - 1) / 128 + 1;
}

// One tile = KS*KS*NKS tcgen05.mma (M=128, N, K=16), fully unrolled.  rb[dy] is the ring position
// (16-byte units, LBO field already OR-ed in) of the tile's first pixel in halo row dy, with the ring
// wrap resolved per ROW by the caller: every descriptor is then one uniform add of a compile-time
// multiple of a run-time stride, so the issuing thread spends a few uniform-datapath instructions
// per MMA (measured: per-tap wrap selects made the issue loop, not the tensor pipe, the bound).
//   NKS >  0 : C-channel input, NKS = C/16 K-steps per tap; LBO = plane stride
//   NKS == 0 : one 8-channel plane, two taps per K=16 step; LBO = flat distance between the taps
// B image (weights) is [mma][k-half][n][8]: LBO = N (16-byte units), SBO = 128 B for A and B.
//   PST16 > 0: the plane stride is a compile-time constant (geometry-specialised instantiation): every A
//   offset becomes an immediate, which is what keeps ptxas from hoisting/spilling uniform registers.
template <int N, int KS, int NKS, int PST16, bool TF>
__device__ __forceinline__ void tc_issue_tile(uint32_t d_tmem, const uint32_t (&rb)[KS], uint32_t Ps,
                                              const uint32_t (&a_off)[40], uint32_t w_base16, uint32_t idesc) {
  constexpr uint64_t DESC_HI = (uint64_t)(8u | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
  // start addresses stay below 2^14, so the LBO field can be OR-ed in once and offsets added after
  const uint32_t wb = w_base16 | ((uint32_t)N << 16);
  if constexpr (NKS > 0) {
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          const int e = (dy * KS + dx) * NKS + ks;
          const uint32_t off = PST16 ? (uint32_t)(dx + 2 * ks * PST16) : a_off[dx * NKS + ks];
          tc_mma<TF>(d_tmem, DESC_HI | (rb[dy] + off),
                     DESC_HI | (wb + (uint32_t)(e * 2 * N)), idesc, e > 0 ? 1u : 0u);
        }
      }
    }
  } else {
    constexpr int KK = KS * KS, NPAIR = (KK + 1) / 2;
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) {
      const int ta = (2 * q + 1 < KK) ? 2 * q : KK - 2;          // the odd tail re-reads tap KK-2 against zeros
      const int tb = ta + 1;
      const uint32_t sa = (uint32_t)(ta / KS) * Ps + (uint32_t)(ta % KS);
      const uint32_t sb = (uint32_t)(tb / KS) * Ps + (uint32_t)(tb % KS);
      tc_mma<TF>(d_tmem, DESC_HI | ((rb[ta / KS] + (uint32_t)(ta % KS)) | ((sb - sa) << 16)),
                 DESC_HI | (wb + (uint32_t)(q * 2 * N)), idesc, q > 0 ? 1u : 0u);
    }
  }
}

// Row-streaming issue (W == 128: a tile is one output row, ring row j is the input row of the strip).
// Input row j contributes to the KS output rows q = j-(KS-1) .. j through tap rows dy = j-q, and all of
// them read the SAME A operand (the row, shifted by dx) -- so one tcgen05.mma per (dx, K-step) with
// N = KS*Cout updates KS accumulators at once: B is the weight blocks [dy=KS-1 | ... | dy=0] stacked along N,
// D is KS consecutive 64-column accumulator slots (slot = output row mod TC_RS_SLOTS).  M=128,N=192,K=16
// costs 96 tensor cycles against 3 x 48 for three N=64 instructions that would each re-read A from shared
// memory (tools/umma_probe.cu): the layer runs at the MMA floor instead of the operand-bandwidth bound.
//   c0 / c1 : accumulator blocks before / after the slot ring wraps (c0 >= 1), first block = weight block boff
//   has_new : the LAST block is a fresh accumulator: its first contribution (step e == 0) must overwrite.
// WRAP: the accumulator blocks straddle the end of the slot ring (second run from slot 0) -- a compile-time
// flag so that the common case carries no predicated-off instructions; id0/id1/id1n/idn are the instruction
// descriptors (N = blocks * Cout) loaded once per row, not once per MMA.
template <int N, int KS, int NKS, int PST16, bool WRAP, bool TF>
__device__ __forceinline__ void tc_issue_row(uint32_t tmem_base, uint32_t d0, uint32_t rb, const uint32_t (&a_off)[40],
                                             uint32_t w_base16, int c0, int c1, int boff, bool has_new, int n0, int n1,
                                             uint32_t id_c0, uint32_t id_c1, uint32_t id_n0, uint32_t id_n1, uint32_t id_1) {
  constexpr uint64_t DESC_HI = (uint64_t)(8u | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
  constexpr uint32_t NB = (uint32_t)(KS * N);                        // rows of one K half of a weight entry
  const uint32_t wb = (w_base16 + (uint32_t)boff * (uint32_t)N) | (NB << 16);   // LBO = NB (16-byte units)
  const uint32_t wb1 = wb + (uint32_t)c0 * (uint32_t)N;                          // first block after the wrap
  // step 0: old blocks accumulate (n0 / n1 of them per run), the new block (last of the last run) overwrites
  const uint32_t dn = WRAP ? tmem_base + (uint32_t)((c1 - 1) * N) : d0 + (uint32_t)((c0 - 1) * N);
  const uint32_t wn = WRAP ? wb1 + (uint32_t)((c1 - 1) * N) : wb + (uint32_t)((c0 - 1) * N);
#pragma unroll
  for (int dx = 0; dx < KS; ++dx) {
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int e = dx * NKS + ks;
      const uint32_t off = PST16 ? (uint32_t)(dx + 2 * ks * PST16) : a_off[dx * NKS + ks];
      const uint64_t ad = DESC_HI | (rb + off);
      const uint32_t eo = (uint32_t)e * 2u * NB;
      if (e == 0) {
        if (n0 > 0) tc_mma<TF>(d0, ad, DESC_HI | (wb + eo), id_n0, 1u);
        if (WRAP && n1 > 0) tc_mma<TF>(tmem_base, ad, DESC_HI | (wb1 + eo), id_n1, 1u);
        if (has_new) tc_mma<TF>(dn, ad, DESC_HI | (wn + eo), id_1, 0u);
      } else {
        tc_mma<TF>(d0, ad, DESC_HI | (wb + eo), id_c0, 1u);
        if (WRAP) tc_mma<TF>(tmem_base, ad, DESC_HI | (wb1 + eo), id_c1, 1u);
      }
    }
  }
}

// named barriers (ids 1..15; 0 is __syncthreads): whole warps only
__device__ __forceinline__ void tc_named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tc_named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Walks the (work item, tile) sequence of one CTA; identical in every role.  A tile starts at flat
// position row*Ps + rem of its item's halo block; all stepping is incremental (an integer division
// costs the single issuing thread more than an MMA).
struct TcTileIter {
  int item, t, ntiles, n, y0, th;
  int row, rem, seg;
  int xoff = 0;                                  // row-streaming: first image column of the item's 128-column segment
  int ps = 0;                                    // compile-time ring pitch of a specialised kernel (0: p.Ps)
  int it_end = 0;                                // table mode (row-streaming): item = index into p.itab
  int cta = 0, ncta = 1;                         // work range of this CTA / number of ranges (grid / nsplit)
  __device__ __forceinline__ void load_item(const TcParams& p) {
    if (p.itab) {
      if (item < it_end) {
        const int4 d = __ldg(p.itab + item);
        n = d.x; y0 = d.y; th = d.z; xoff = d.w;
        ntiles = th;                               // one 128-column tile per row
      }
      t = 0; row = 0; rem = 0; seg = 0;
      return;
    }
    if (item < p.items) {
      const int im = p.rev ? p.items - 1 - item : item;
      int rest = im;
      if (p.rs_segs > 1) { rest = im / p.rs_segs; xoff = (im - rest * p.rs_segs) * 128; }
      n = rest / p.strips;                       // one division per work item (>= 8 tiles)
      y0 = (rest - n * p.strips) * p.TH;
      th = (p.H - y0 < p.TH) ? (p.H - y0) : p.TH;
      ntiles = tc_num_tiles(p, th);
    }
    t = 0; row = 0; rem = 0; seg = 0;
  }
  __device__ __forceinline__ void init(const TcParams& p) {
    cta = (int)blockIdx.x / p.nsplit; ncta = (int)gridDim.x / p.nsplit;
    if (p.itab) { item = __ldg(p.coff + cta); it_end = __ldg(p.coff + cta + 1); }
    else item = cta;
    load_item(p);
  }
  __device__ __forceinline__ bool valid(const TcParams& p) const { return p.itab ? item < it_end : item < p.items; }
  __device__ __forceinline__ bool last_of_item() const { return t + 1 >= ntiles; }
  __device__ __forceinline__ void next(const TcParams& p) {
    if (++t >= ntiles) { item += p.itab ? 1 : ncta; load_item(p); return; }
    if (p.itab) { ++row; return; }
    const int Ps = ps ? ps : p.Ps;
    if (p.segs) {
      if (++seg == p.segs) { seg = 0; rem = 0; ++row; } else rem += 128;
    } else {
      rem += 128;
      while (rem >= Ps) { rem -= Ps; ++row; }
    }
  }
};

// Transposed warp reduction: every lane holds NV values (NV = 32 or 16); afterwards lane (j << log2(32/NV)) -- and
// the lanes that differ from it in the low log2(32/NV) bits -- holds in v[0] the sum over all 32 lanes of value j.
// In the round with lane mask M, lanes with that bit clear keep the lower half of the surviving values and send the
// upper half (and vice versa): 31 shuffles for NV = 32 instead of 5 per value.
template <int NV, int M>
__device__ __forceinline__ void tc_transposed_sum(float* v, int lane) {
  if constexpr (M >= 1) {
    if constexpr (NV > 1) {
      const bool up = (lane & M) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        const float send = up ? v[i] : v[i + NV / 2];
        const float keep = up ? v[i + NV / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
      }
      tc_transposed_sum<NV / 2, M / 2>(v, lane);
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], M);
      tc_transposed_sum<1, M / 2>(v, lane);
    }
  }
}

// PS / PST16: ring pitch and plane stride as compile-time constants (0 = read them from the parameters).
// RS: row-streaming variant (W == 128, one tile per output row): see tc_issue_row.
// TF: tf32 operands -- planes hold 4 fp32 channels (still 16 bytes per pixel), K = 8 per MMA = two planes, NKS = C/8.
template <int N, int EPI, int KS, int NKS, int PS, int PST16, bool RS, bool TF>
__global__ void __launch_bounds__(tc_threads(N, RS, tc_iss2(EPI, RS, TF)), 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int PW = TF ? 4 : 8;                       // channels per plane
  const int cta = (int)blockIdx.x / p.nsplit, ncta = (int)gridDim.x / p.nsplit;
  const int csplit = (int)blockIdx.x - cta * p.nsplit; // this CTA's block of N output channels
  constexpr bool ISS2 = tc_iss2(EPI, RS, TF);
  constexpr int NTHREADS = tc_threads(N, RS, ISS2);
  constexpr int EW = tc_epi_warps(N);
  constexpr int NC = (EW == 8) ? N / 2 : N;            // accumulator columns per epilogue warp
  constexpr int ACC = RS ? (N <= 32 ? TC_RS_SLOTS_MAX : TC_RS_SLOTS) : TC_ACC_STAGES;       // accumulator stages (TMEM)
  constexpr int TMEM_COLS = (ACC * N < 32) ? 32 : ACC * N;
  const uint32_t w_region = (p.w_bytes + 1023u) & ~1023u;
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + w_region;
  const uint32_t plane_stride16 = PST16 ? (uint32_t)PST16 : p.plane_stride16;
  const uint32_t plane_bytes = plane_stride16 * 16u;
  TcSmem* sb = reinterpret_cast<TcSmem*>(s_a + (size_t)p.nch_in * plane_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int pad = KS / 2;
  const int Ps = PS ? PS : p.Ps, R = p.R;
  const int F16 = p.f16;

  // ---- one-time setup: zero the ring (no stale NaN patterns under discarded positions),
  //      barriers, TMEM
  {
    uint4* a4 = reinterpret_cast<uint4*>(s_a);
    const int n16 = p.nch_in * (int)plane_stride16;
    for (int i = threadIdx.x; i < n16; i += NTHREADS) a4[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) {
      mbar_init(smem_u32(&sb->full[i]), RS ? 2 : 1);   // row-streaming: two producer warps
      mbar_init(smem_u32(&sb->empty[i]), RS ? 1 : TC_ISSUERS);   // every issuer hands every row back once
    }
    for (int i = 0; i < ACC; ++i) {
      mbar_init(smem_u32(&sb->tfull[i]), ISS2 ? 2 : 1);   // two issuers: the owners of the row's last two input rows commit
      mbar_init(smem_u32(&sb->tempty[i]), EW);     // one arrive per epilogue warp
    }
    mbar_init(smem_u32(&sb->wbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sb->tmem_base;

  if (RS && (warp == 0 || warp == 2)) {
    // =============================================================== producers, row-streaming
    // W == 128: the left/right halo columns of a ring row are always zero padding, so they are never
    // loaded (the ring is zeroed once) and a halo row is one contiguous 2 KB run per 8-channel plane:
    // plain 1-D bulk copies, half of the planes per producer warp (a tensor-map copy costs ~110-140 issue
    // cycles, and 16 of them per row -- 8 planes x (128-pixel box + 2-pixel tail) -- made a single
    // producer the bound of the layer).  Rows above / below the image are copied from a zero row.
    if constexpr (RS) {
      const int pw = warp >> 1;                    // producer 0 / 1
      if (warp == 0 && lane == 0) {                // weights: one shot, resident for the whole kernel
        const uint32_t wbar = smem_u32(&sb->wbar);
        mbar_expect_tx(wbar, p.w_bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) + (size_t)csplit * p.w_bytes;
        for (uint32_t off = 0; off < p.w_bytes; off += 16384u) {
          const uint32_t n = (p.w_bytes - off < 16384u) ? (p.w_bytes - off) : 16384u;
          bulk_load_1d(smem_u32(s_w + off), src + off, n, wbar);
        }
      }
      const bool prod_leader = elect_one_sync();
      const int c_lo = (p.nch_in * pw) / 2, c_hi = (p.nch_in * (pw + 1)) / 2;     // this producer's planes
      const uint32_t row_bytes = (uint32_t)Ps * 16u;
      const uint32_t a_base = smem_u32(s_a);
      const size_t plane_px = (size_t)p.H * p.W;
      const int segs = p.rs_segs > 1 ? p.rs_segs : 1;
      int slot = 0;
      uint32_t phase = 0;
      // run-ahead cap: before row g is requested, row g - lead must have been consumed.  Everything the ring lets
      // the producers request sits in the memory system's queues in front of the epilogue's own loads (saved
      // activation), so a full ring (9 rows x 16 KB per SM) costs those loads microseconds of queueing.
      const int lead = (p.lead > 0 && p.lead < R) ? p.lead : 0;
      int g_row = 0, slot2 = 0;
      uint32_t phase2 = 0;
      if (prod_leader) {
        for (int item = __ldg(p.coff + cta), it_end = __ldg(p.coff + cta + 1); item < it_end; ++item) {
          const int4 d = __ldg(p.itab + item);
          const int n = d.x, y0 = d.y, seg = d.w >> 7;
          const int nrows = d.z + 2 * pad;
          // columns copied per row: the segment plus the neighbouring pixel on every side that lies inside the
          // image; a side on the image border keeps ring halo zeros (W > 128: re-zeroed per row, the slot may
          // have held a row of another segment)
          const int left = (seg > 0) ? pad : 0, right = (seg < segs - 1) ? pad : 0;
          const uint32_t run_bytes = (uint32_t)(128 + left + right) * 16u;
          const uint32_t dst_off = (uint32_t)(pad - left) * 16u;
          const bool zl = segs > 1 && !left, zr = segs > 1 && !right;
          const uint32_t per_plane = run_bytes + (zl ? (uint32_t)pad * 16u : 0u) + (zr ? (uint32_t)pad * 16u : 0u);
          for (int j = 0; j < nrows; ++j) {
            const uint32_t fb = smem_u32(&sb->full[slot]);
            mbar_wait(smem_u32(&sb->empty[slot]), phase ^ 1u, 1);
            if (lead && g_row >= lead) {
              mbar_wait(smem_u32(&sb->empty[slot2]), phase2, 1);
              if (++slot2 == R) { slot2 = 0; phase2 ^= 1u; }
            }
            ++g_row;
            if (p.dbg & 16) {                      // timing experiment: no input traffic at all (the ring keeps its zeros)
              mbar_arrive(fb);
              if (++slot == R) { slot = 0; phase ^= 1u; }
              continue;
            }
            mbar_expect_tx(fb, per_plane * (uint32_t)(c_hi - c_lo));              // (arrives even with no plane)
            const int y = y0 - pad + j;
            const bool inside = y >= 0 && y < p.H;
            const uint4* src = inside ? p.in + ((size_t)n * p.nch_in + c_lo) * plane_px + (size_t)y * p.W + (seg * 128 - left)
                                      : p.zero_row;
            const size_t sstep = inside ? plane_px : 0;
            uint32_t dst = a_base + (uint32_t)slot * row_bytes + (uint32_t)c_lo * plane_bytes;
            for (int c = c_lo; c < c_hi; ++c) {
              bulk_load_1d(dst + dst_off, src, run_bytes, fb);
              if (zl) bulk_load_1d(dst, p.zero_row, (uint32_t)pad * 16u, fb);
              if (zr) bulk_load_1d(dst + (uint32_t)(pad + 128) * 16u, p.zero_row, (uint32_t)pad * 16u, fb);
              src += sstep;
              dst += plane_bytes;
            }
            // The epilogue multiplies output row y by ELU' of the saved activation: it loads that row with plain
            // loads about a ring depth + a TMEM queue later.  Those loads queue behind everything the producers
            // have requested (microseconds under load), which three tiles of register prefetch do not cover --
            // pull the row into L2 now.
            if (p.apf && inside && j >= pad && j < nrows - pad) {
              // (this CTA's N/PW planes of the saved activation, half of them per producer warp)
              const int np = N / PW, q_lo = (np * pw) / 2, q_hi = (np * (pw + 1)) / 2;
              const uint4* ap = p.actp + ((size_t)n * p.nch_out + csplit * np + q_lo) * plane_px + (size_t)y * p.W + seg * 128;
              for (int c = q_lo; c < q_hi; ++c) {
                bulk_prefetch_l2(ap, 128u * 16u);
                ap += plane_px;
              }
            }
            if (++slot == R) { slot = 0; phase ^= 1u; }
          }
        }
      }
      __syncwarp();
    }
  } else if (!RS && warp == 0) {
    // =============================================================== TMA producer (generic kernels only)
    // lane 0 owns the barriers; the copies of one halo row (planes x mirror copies x boxes) are
    // issued by as many lanes in parallel
    if constexpr (!RS) {
    if (lane == 0) {  // weights: one shot, resident for the whole kernel
      const uint32_t wbar = smem_u32(&sb->wbar);
      mbar_expect_tx(wbar, p.w_bytes);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) + (size_t)csplit * p.w_bytes;
      for (uint32_t off = 0; off < p.w_bytes; off += 16384u) {
        const uint32_t n = (p.w_bytes - off < 16384u) ? (p.w_bytes - off) : 16384u;
        bulk_load_1d(smem_u32(s_w + off), src + off, n, wbar);
      }
    }
    const uint32_t a_base = smem_u32(s_a);
    const uint32_t row_bytes = (uint32_t)Ps * 16u;
    const int nbox = p.n_full + (p.tail_px ? 1 : 0);
    // One elected lane issues every copy of a halo row (planes x boxes, plus the mirror copies) with
    // uniform-register arithmetic: spreading the copies over lanes serialises them anyway (each UTMALDG is
    // issued per active lane) and cost ~110 cycles per copy, which made the producer the bound of the layer.
    const bool prod_leader = elect_one_sync();
    int slot = 0;
    uint32_t phase = 0;
    for (int item = cta; item < p.items; item += ncta) {
      const int im = p.rev ? p.items - 1 - item : item;
      const int n = im / p.strips, y0 = (im - n * p.strips) * p.TH;
      const int th = (p.H - y0 < p.TH) ? (p.H - y0) : p.TH;
      const int nrows = th + 2 * pad;
      for (int j = 0; j < nrows; ++j) {
        const uint32_t fb = smem_u32(&sb->full[slot]);
        const bool mirrored = slot < p.m;
        if (prod_leader) {
          mbar_wait(smem_u32(&sb->empty[slot]), phase ^ 1u, 1);
          if (p.dbg & 4) {
            mbar_arrive(fb);
          } else {
            mbar_expect_tx(fb, p.box_bytes * (uint32_t)p.nch_in * (mirrored ? 2u : 1u));
            const int y = y0 - pad + j;
            const uint32_t dst_row = a_base + (uint32_t)slot * row_bytes;
            const uint32_t mir = (uint32_t)R * row_bytes;
            for (int c = 0; c < p.nch_in; ++c) {
              const uint32_t d = dst_row + (uint32_t)c * plane_bytes;
              for (int k = 0; k < nbox; ++k) {
                const CUtensorMap* mp = (k >= p.n_full) ? &p.maps.tail : &p.maps.full;
                const int x8 = 256 * k - 2 * pad;    // 8-byte elements (2 per pixel-plane): box k starts at pixel 128k - pad
                tma_load_4d(d + (uint32_t)k * 2048u, mp, fb, x8, y, c, n);
                if (mirrored) tma_load_4d(d + mir + (uint32_t)k * 2048u, mp, fb, x8, y, c, n);
              }
            }
          }
        }
        if (++slot == R) { slot = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
    }
  } else if (RS && (warp == 1 || (ISS2 && warp == 4 + EW))) {
    // =============================================================== MMA issuer(s), row-streaming
    // Input rows are consumed strictly in order, each exactly once.  The per-row bookkeeping (ring slot, accumulator
    // blocks, weight block, which barriers to signal) and every wait on the ring / accumulator barriers are the SCOUT
    // warp's (below): the tensor pipe's instruction queue is short, and serial arithmetic + barrier polls in the issuing
    // thread left it idle.  What remains per row is the ~100 uniform-datapath instructions that build its descriptors
    // (ptxas schedules them ahead of the row's MMA burst) -- with ISS2 two warps take alternate rows, so that one builds
    // its descriptors while the other's burst runs; a token (named barriers 1 / 2) keeps the bursts in row order, which
    // the overwrite of a fresh accumulator needs.  Completion: a thread's tcgen05.commit covers its own MMAs, so an
    // accumulator is complete when the owners of its last TWO input rows have both committed to it.
    if constexpr (RS) {
      {
        const int w = (warp == 1) ? 0 : 1;         // issuer index: rows with (running row number & 1) == w
        const bool leader = elect_one_sync();
        mbar_wait(smem_u32(&sb->wbar), 0, 2);
        const uint32_t w_base16 = smem_u32(s_w) >> 4;
        const uint32_t id_step = (uint32_t)(N >> 3) << 17, id_0 = p.idesc_n[1] - id_step;   // idesc of c blocks = id_0 + c id_step
        int pp = 0;
        int g = 0;                                 // running row number of the CTA
        for (int item = __ldg(p.coff + cta), it_end = __ldg(p.coff + cta + 1); item < it_end; ++item) {
          const int TH = __ldg(p.itab + item).z;   // rows of this item (1 .. H)
          const int nrows = TH + 2 * pad;
          const int unit = p.rs_unit;
          for (int j0 = 0; j0 < nrows; j0 += unit) {
            const int ju = (nrows - j0 < unit) ? nrows - j0 : unit;
            tc_named_sync((w ? 7 : 3) + pp, 64);   // the scout has published this unit's plans
            tc_fence_after();
            const uint4 mine = sb->plan[pp][lane & 15];
            const bool last_unit = item + 1 == it_end && j0 + unit >= nrows;
            for (int u = 0; u < ju; ++u, ++g) {
              if (ISS2 && (g & 1) != w) continue;
              const uint32_t rb = __shfl_sync(0xffffffffu, mine.x, u);
              const uint32_t d0 = __shfl_sync(0xffffffffu, mine.y, u);
              const uint32_t pk = __shfl_sync(0xffffffffu, mine.z, u);
              const uint32_t wb_t = __shfl_sync(0xffffffffu, w_base16, 0);
              const int c0 = (int)(pk & 7u), c1 = (int)((pk >> 3) & 7u), boff = (int)((pk >> 6) & 3u);
              const bool has_new = ((pk >> 8) & 1u) != 0;
              const int n0 = (int)((pk >> 9) & 7u), n1 = (int)((pk >> 12) & 7u);
              const uint32_t tf = (pk >> 15) & 31u, sl = (pk >> 20) & 31u, tf1 = (pk >> 25) & 31u;
              // instruction descriptor for c accumulator blocks: N = c * Cout sits in bits 17.. (make_idesc)
              const uint32_t id_c0 = id_0 + (uint32_t)c0 * id_step, id_c1 = id_0 + (uint32_t)c1 * id_step,
                             id_n0 = id_0 + (uint32_t)n0 * id_step, id_n1 = id_0 + (uint32_t)n1 * id_step,
                             id_1 = id_0 + id_step;
              if (ISS2 && g > 0) tc_named_sync(w == 0 ? 2 : 1, 64);   // the previous row's burst has been issued
              if (leader) {
                if (c1 > 0)
                  tc_issue_row<N, KS, NKS, PST16, true, TF>(tmem_base, d0, rb, p.a_off, wb_t, c0, c1, boff, has_new, n0, n1, id_c0,
                                                              id_c1, id_n0, id_n1, id_1);
                else
                  tc_issue_row<N, KS, NKS, PST16, false, TF>(tmem_base, d0, rb, p.a_off, wb_t, c0, c1, boff, has_new, n0, n1, id_c0,
                                                               id_c1, id_n0, id_n1, id_1);
                if (tf != 31u) tc_commit(smem_u32(&sb->tfull[tf]));              // this row was the accumulator's last input row
                if (ISS2 && tf1 != 31u) tc_commit(smem_u32(&sb->tfull[tf1]));    // ... its last but one
                tc_commit(smem_u32(&sb->empty[sl]));                             // ring row consumed
              }
              __syncwarp();
              if (ISS2 && !(last_unit && u == ju - 1)) tc_named_arrive(w == 0 ? 1 : 2, 64);   // hand the token on (nobody waits for the CTA's last row)
            }
            __syncwarp();
            tc_named_arrive(5 + pp, ISS2 ? 96 : 64);   // plans consumed
            pp ^= 1;
          }
        }
      }
    }
  } else if (RS && warp == 3 + EW) {
    // =============================================================== scout, row-streaming
    // Walks the same (item, unit, row) sequence one unit ahead of the issuer: lane u works out what row j0 + u of the
    // unit needs, the lanes wait -- in parallel -- for the ring rows to land and for the accumulator slots the unit
    // opens to be drained, then the plans go to shared memory and one arrive releases the unit to the issuer.
    if constexpr (RS) {
      const uint32_t a_base16 = smem_u32(s_a) >> 4;
      int slot = 0; uint32_t rph = 0;            // ring slot / phase of the next input row
      int qg = 0;                                // output rows (tiles) of all previous items of this CTA
      int pp = 0, nunit = 0;
      for (int item = __ldg(p.coff + cta), it_end = __ldg(p.coff + cta + 1); item < it_end; ++item) {
        const int TH = __ldg(p.itab + item).z;   // rows of this item (1 .. H)
        const int nrows = TH + 2 * pad;
        const int unit = p.rs_unit;
        for (int j0 = 0; j0 < nrows; j0 += unit, ++nunit) {
          const int ju = (nrows - j0 < unit) ? nrows - j0 : unit;
          uint4 plan = make_uint4(0u, 0u, 0u, 0u);
          {
            const int u = (lane < ju) ? lane : 0;
            const int j = j0 + u;
            int sl = slot + u;
            if (sl >= R) sl -= R;
            // output rows of this strip fed by input row j: q in [j-(KS-1), j] clipped to the strip
            const int qa = (j - (KS - 1) > 0) ? j - (KS - 1) : 0;
            const int qb = (j < TH - 1) ? j : TH - 1;
            // row q = j opens a fresh accumulator: the epilogue left the slot zeroed when it drained its previous row, so only
            // a slot's FIRST row of the kernel (TMEM is not initialised) overwrites on its first step
            const int has_new = (j <= TH - 1) && (qg + j < ACC);
            const int cnt = qb - qa + 1;                       // accumulator blocks written by this row
            const int sa = (qg + qa) & (ACC - 1);              // slot of the first one
            const int c0 = (cnt < ACC - sa) ? cnt : ACC - sa;  // blocks before the slot ring wraps
            const int c1 = cnt - c0;
            const int boff = qa - (j - (KS - 1));              // first weight block (block b <-> dy = KS-1-b)
            // blocks that already hold partial sums (step 0 accumulates into them) per run
            const int n0 = has_new ? (c1 ? c0 : c0 - 1) : c0;
            const int n1 = has_new ? (c1 ? c1 - 1 : 0) : c1;
            const int tf = (j >= KS - 1) ? ((qg + j - (KS - 1)) & (ACC - 1)) : 31;   // accumulator this row completes
            // accumulator whose last but one input row this is (two issuers: its owner commits to it as well)
            const int tf1 = (j >= KS - 2 && j - (KS - 2) <= TH - 1) ? ((qg + j - (KS - 2)) & (ACC - 1)) : 31;
            plan.x = (a_base16 + (uint32_t)(sl * Ps)) | (plane_stride16 << 16);
            plan.y = tmem_base + (uint32_t)(sa * N);
            plan.z = (uint32_t)c0 | ((uint32_t)c1 << 3) | ((uint32_t)boff << 6) | ((uint32_t)has_new << 8) | ((uint32_t)n0 << 9) |
                     ((uint32_t)n1 << 12) | ((uint32_t)tf << 15) | ((uint32_t)sl << 20) | ((uint32_t)tf1 << 25);
          }
          plan.w = (uint32_t)ju | ((item + 1 == it_end && j0 + unit >= nrows) ? 256u : 0u);   // rows of the unit | the CTA's last unit
          if (nunit >= 2) tc_named_sync(5 + pp, ISS2 ? 96 : 64);   // the unit that used this plan buffer has been issued
          // lane u waits on the ring row of unit row u, lane 16+u on the accumulator slot that row opens
          if (lane < ju) {
            int sl = slot + lane; uint32_t ph = rph;
            if (sl >= R) { sl -= R; ph ^= 1u; }
            mbar_wait(smem_u32(&sb->full[sl]), ph, 3);
          } else if (lane >= 16 && lane - 16 < ju && j0 + lane - 16 <= TH - 1) {
            const int qn = qg + j0 + lane - 16;  // output row q = j receives its first contribution: fresh slot
            mbar_wait(smem_u32(&sb->tempty[qn & (ACC - 1)]), (((uint32_t)qn / ACC) & 1u) ^ 1u, 4);
          }
          __syncwarp();
          if (lane < ju) sb->plan[pp][lane] = plan;
          __syncwarp();
          tc_named_arrive(3 + pp, 64);
          if (ISS2) tc_named_arrive(7 + pp, 64);
          pp ^= 1;
          slot += ju;
          if (slot >= R) { slot -= R; rph ^= 1u; }
        }
        qg += TH;
      }
    }
  } else if (!RS && warp <= TC_ISSUERS) {
    // =============================================================== MMA issuers (generic kernels only: the row-
    // streaming instantiations must not carry these 2*KS*KS*NKS MMAs and their per-MMA issue loops as dead code)
    // All lanes walk the tile sequence (so the per-tile bases are warp-uniform values); one elected
    // lane issues the MMAs and commits.  Issuer w = warp-1 owns tiles w, w+TC_ISSUERS, ...
    if constexpr (!RS) {
    const int w = warp - 1;
    // lane 0 rather than elect.sync on purpose: behind a plain lane test ptxas wraps every UTCHMMA in its
    // own small issue block, which keeps descriptor arithmetic next to its MMA; with elect.sync it hoists
    // all 2*KS*KS*NKS descriptors above the first MMA and spills uniform registers (measured slower)
    const bool leader = (lane == 0);
    mbar_wait(smem_u32(&sb->wbar), 0, 2);
    const uint32_t a_base16 = smem_u32(s_a) >> 4;
    const uint32_t w_base16 = smem_u32(s_w) >> 4;
    // Lean tile iterator: every work item is TH rows (TH divides H), so a tile is (item ordinal k, tile t)
    // with constant tiles / halo rows per item, stepped without divisions.
    const int NT = p.NT, nrows = p.TH + 2 * pad, segs = p.segs;
    const int my_items = ((int)p.items - cta + ncta - 1) / ncta;
    int k = 0, t = 0, row = 0, rem = 0, seg = 0;
    // ring bookkeeping in (slot, phase) form; g_* count halo rows since the kernel started
    int g0 = 0, slot0 = 0;                         // first halo row of the iterator's current item
    int g_ready = 0, rs = 0; uint32_t rph = 0;     // rows known to have landed
    int g_freed = 0, fs = 0;                       // rows this issuer has handed back to the producer
    int tcnt = 0;                                  // index of the iterator's tile in the CTA's sequence
    auto advance = [&]() {
      ++tcnt;
      if (++t == NT) {
        t = 0; row = 0; rem = 0; seg = 0; ++k;
        g0 += nrows;
        slot0 += nrows;
        while (slot0 >= R) slot0 -= R;
      } else if (segs) {
        if (++seg == segs) { seg = 0; rem = 0; ++row; } else rem += 128;
      } else {
        rem += 128;
        while (rem >= Ps) { rem -= Ps; ++row; }
      }
    };
    for (int i = 0; i < w && k < my_items; ++i) advance();
    while (k < my_items) {
      // last halo row this tile reads: row of (start + 127 + max tap shift)
      int need = row + 2 * pad;
      for (int v = rem + 127 + 2 * pad; v >= Ps; v -= Ps) ++need;
      if (need > nrows - 1) need = nrows - 1;
      while (g_ready <= g0 + need) {
        mbar_wait(smem_u32(&sb->full[rs]), rph, 3);
        ++g_ready;
        if (++rs == R) { rs = 0; rph ^= 1u; }
      }
      const int stage = tcnt % TC_ACC_STAGES;
      const uint32_t sph = (uint32_t)(tcnt / TC_ACC_STAGES) & 1u;
      mbar_wait(smem_u32(&sb->tempty[stage]), sph ^ 1u, 4);
      tc_fence_after();
      // ring position of the tile's first pixel in each of its KS halo rows (wrap resolved per row; the
      // mirrored rows behind the ring end keep every 128-run contiguous)
      uint32_t rb[KS];
      {
        int pr = slot0 + row;
        while (pr >= R) pr -= R;
        const uint32_t lbo_or = (NKS > 0) ? (plane_stride16 << 16) : 0u;
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          rb[dy] = __shfl_sync(0xffffffffu, (a_base16 + (uint32_t)(pr * Ps + rem)) | lbo_or, 0);
          if (++pr == R) pr = 0;
        }
      }
      const uint32_t d_tmem = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)(stage * N), 0);
      // re-derived per tile on purpose: a loop-invariant weight base makes ptxas hoist all KS*KS*NKS B
      // descriptors into vector registers and pay an R2UR (plus uniform-register spills) per MMA
      const uint32_t wb_t = __shfl_sync(0xffffffffu, w_base16, 0);
      // step to this issuer's next tile: rows below its first row are not read by this issuer again
      for (int i = 0; i < TC_ISSUERS && k < my_items; ++i) advance();
      const int free_upto = (k < my_items) ? g0 + row : g0;
      if (leader) {
        tc_issue_tile<N, KS, NKS, PST16, TF>(d_tmem, rb, (uint32_t)Ps, p.a_off, wb_t, p.idesc);
        tc_commit(smem_u32(&sb->tfull[stage]));
        int f = fs;
        for (int rf = g_freed; rf < free_upto; ++rf) {
          tc_commit(smem_u32(&sb->empty[f]));
          if (++f == R) f = 0;
        }
      }
      while (g_freed < free_upto) { ++g_freed; if (++fs == R) fs = 0; }
      __syncwarp();
    }
    // an issuer that ran out of tiles early still owes its arrival on the remaining rows
    if (leader) {
      int f = fs;
      for (int rf = g_freed; rf < g0; ++rf) {
        tc_commit(smem_u32(&sb->empty[f]));
        if (++f == R) f = 0;
      }
    }
    __syncwarp();
    }
  } else {
    // =============================================================== epilogue warps
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    const int half = (EW == 8) ? (warp - 1 - TC_ISSUERS) / 4 : 0;
    const int col0 = half * NC;                    // first accumulator column / output channel
    const int k0 = csplit * (N / PW) + col0 / PW;   // first output plane of this warp in the layer's plane list
    const int H = p.H, W = p.W;
    const size_t plane_sz = (size_t)H * W;
    constexpr bool DG = (EPI == EPI_DGRAD || EPI == EPI_DSUM);   // epilogues that read the saved activation
    static_assert(EPI != EPI_DSUM || (RS && KS == 3), "EPI_DSUM is a row-streaming 3x3 epilogue");
    constexpr int NB = (EPI == EPI_FWD) ? NC : 4;
    float bias_r[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) bias_r[i] = DG ? 0.f : __ldg(p.bias + csplit * N + col0 + i);

    constexpr int NAV = DG ? NC / PW : 1;
    // position of this thread's output pixel in tile `it`, or -1
    auto pix_of = [&](const TcTileIter& it) -> long long {
      int r = it.row, c = it.rem + quad * 32 + lane;
      while (c >= Ps) { c -= Ps; ++r; }
      c += it.xoff;
      if (c >= W || r >= it.th) return -1;
      return (long long)(it.y0 + r) * W + c;
    };
    auto load_prev = [&](const TcTileIter& it, long long pix, uint4* av) {
#pragma unroll
      for (int k = 0; k < NAV; ++k)
        av[k] = (pix >= 0 && !(p.dbg & 8)) ? __ldg(p.actp + ((size_t)it.n * p.nch_out + k0 + k) * plane_sz + (size_t)pix)
                           : make_uint4(0u, 0u, 0u, 0u);
    };

    // Tiles are walked by ONE iterator that runs three tiles ahead of the tile being written: the loads of
    // the previous activation (data-gradient epilogue) need ~3 tiles of lead under load -- with one tile of
    // lead the 16 KB per tile in flight per SM capped that stream below 1 TB/s and the epilogue, not the
    // tensor pipe, bounded the layer.
    TcTileIter far;
    far.ps = PS;
    far.init(p);
    int ntotal;                                    // tiles this CTA writes
    if (p.itab) {
      ntotal = 0;
      for (int i = __ldg(p.coff + cta); i < __ldg(p.coff + cta + 1); ++i) ntotal += __ldg(p.itab + i).z;
    } else {
      const int my_items_e = ((int)p.items - cta + ncta - 1) / ncta;
      ntotal = (my_items_e > 0 ? my_items_e : 0) * p.NT;
    }
    // EPI_DSUM keeps NC running sums per thread; it pays for those registers with the prefetch queue (one tile of
    // lead instead of three -- the producers pull its saved-activation rows into L2 long before, see `apf`), so
    // that nothing spills: the register cap of this 11-warp block is 168.
    constexpr bool PF3 = (EPI != EPI_DSUM);
    long long pix = -1, pix1 = -1, pix2 = -1, pix3 = -1;
    int cn = 0, cn1 = 0, cn2 = 0, cn3 = 0;
    uint4 av[NAV], av1[NAV], av2[PF3 ? NAV : 1], av3[PF3 ? NAV : 1];
    auto fetch = [&](long long& pix_o, int& n_o, uint4* av_o) {
      if (far.valid(p)) {
        pix_o = pix_of(far);
        n_o = far.n;
        if constexpr (DG) load_prev(far, pix_o, av_o);
        far.next(p);
      } else {
        pix_o = -1;
      }
    };
    fetch(pix, cn, av);
    if constexpr (PF3) {
      fetch(pix1, cn1, av1);
      fetch(pix2, cn2, av2);
    }
    int stage = 0;
    uint32_t sph = 0;
    // EPI_DSUM state: every thread keeps running sums of its own pixel column over the rows of the current
    // (slot-image, row class, segment) -- NC registers, one FFMA per value and tile -- and the warp reduces them
    // only when that triple changes (a few times per image): the thread of image column 0 / W-1 then holds exactly
    // the border-column class, the transposed warp sum the whole row class.
    constexpr int DS_SH = (NC == 32) ? 0 : (NC == 16) ? 1 : 2;
    constexpr int DS_NV = (EPI == EPI_DSUM) ? NC : 1;
    float ds_sum[DS_NV];
#pragma unroll
    for (int j = 0; j < DS_NV; ++j) ds_sum[j] = 0.f;
    int ds_n = -1, ds_rc = 0, ds_pix = 0;
    auto ds_flush = [&]() {
      if constexpr (EPI == EPI_DSUM) {
        if (ds_n >= 0) {                                            // warp-uniform
          const int xcol = ds_pix % W;
          const uint32_t bl = __ballot_sync(0xffffffffu, xcol == 0 || xcol == W - 1);
          float col = 0.f;
          int side = 0;
          if (bl) {
            const int src = __ffs(bl) - 1;
            side = (__shfl_sync(0xffffffffu, xcol, src) == 0) ? 0 : 2;
#pragma unroll
            for (int j = 0; j < DS_NV; ++j) {
              const float tv = __shfl_sync(0xffffffffu, ds_sum[j], src);
              if (lane == (j << DS_SH)) col = tv;
            }
          }
          tc_transposed_sum<DS_NV, 16>(ds_sum, lane);               // lane (j << DS_SH): row-class sum of channel j
          if ((lane & ((1 << DS_SH) - 1)) == 0) {
            float* g = p.G + ((size_t)ds_n * 9 + ds_rc * 3) * p.n_tot + csplit * N + col0 + (lane >> DS_SH);
            atomicAdd(g + p.n_tot, ds_sum[0] - col);                // interior columns
            if (bl) atomicAdd(g + side * p.n_tot, col);
          }
        }
#pragma unroll
        for (int j = 0; j < DS_NV; ++j) ds_sum[j] = 0.f;
      }
    };
    for (int tno = 0; tno < ntotal; ++tno) {
      if constexpr (PF3) fetch(pix3, cn3, av3);
      else fetch(pix1, cn1, av1);

      mbar_wait(smem_u32(&sb->tfull[stage]), sph, 5);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(stage * N + col0) + ((uint32_t)(quad * 32) << 16);
      uint32_t acc[NC];
      if (p.dbg & 1) {
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = 0u;
      } else {
#pragma unroll
        for (int q = 0; q < NC / 16; ++q) IOD_TMEM_LD16((acc + q * 16), taddr + q * 16);
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if constexpr (RS) {
        // row-streaming: leave the slot ZEROED for its next output row, so that the row's first MMA accumulates like
        // every other one (an overwriting first step would have to be a separate N = Cout instruction next to the
        // accumulating N = 2 Cout one: one more pass over the A operand per row)
        if (!(p.dbg & 1)) {
#pragma unroll
          for (int q = 0; q < NC / 16; ++q) tmem_zero16(taddr + q * 16);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sb->tempty[stage]));   // accumulator stage is free again

      if constexpr (EPI == EPI_DSUM) {
        // Row-streaming tiles are one full 128-pixel run of one image row: all 32 lanes hold a pixel of that row
        // (this thread's column is the same for every row of an item), the warp channels col0 .. col0+NC-1.
        const int pi = (int)pix;
        const int rc = (pi < W) ? 0 : (pi >= (H - 1) * W) ? 2 : 1;    // border class of the row (3x3: pad 1)
        if (cn != ds_n || rc != ds_rc || pi != ds_pix + W) {          // warp-uniform
          ds_flush();
          ds_n = cn; ds_rc = rc;
        }
        ds_pix = pi;
        auto accumulate = [&](auto h16) {
          constexpr bool H16 = decltype(h16)::value;
          if constexpr (TF) {
#pragma unroll
            for (int k = 0; k < NC / 4; ++k) {
              const uint32_t aw[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                ds_sum[k * 4 + e] = fmaf(__uint_as_float(acc[k * 4 + e]), fminf(__uint_as_float(aw[e]), 0.f) + 1.f, ds_sum[k * 4 + e]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < NC / 8; ++k) {
              const uint32_t aw[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 a;
                if constexpr (H16) a = __half22float2(*reinterpret_cast<const __half2*>(&aw[e]));
                else a = make_float2(__uint_as_float(aw[e] << 16), __uint_as_float(aw[e] & 0xffff0000u));
                // ELU'(x) from a = ELU(x): 1 for a > 0, a + 1 otherwise = min(a, 0) + 1
                ds_sum[k * 8 + 2 * e] = fmaf(__uint_as_float(acc[k * 8 + 2 * e]), fminf(a.x, 0.f) + 1.f, ds_sum[k * 8 + 2 * e]);
                ds_sum[k * 8 + 2 * e + 1] = fmaf(__uint_as_float(acc[k * 8 + 2 * e + 1]), fminf(a.y, 0.f) + 1.f, ds_sum[k * 8 + 2 * e + 1]);
              }
            }
          }
          (void)H16;
        };
        if (F16) accumulate(std::true_type{});
        else accumulate(std::false_type{});
      } else
      if (pix >= 0 && !(p.dbg & 2)) {
        if constexpr (EPI == EPI_OUT4) {
          float4 o;
          o.x = __uint_as_float(acc[0]) + bias_r[0];
          o.y = __uint_as_float(acc[1]) + bias_r[1];
          o.z = __uint_as_float(acc[2]) + bias_r[2];
          o.w = __uint_as_float(acc[3]) + bias_r[3];
          reinterpret_cast<float4*>(p.out)[(size_t)cn * plane_sz + (size_t)pix] = o;
        } else {
          uint4* outp = reinterpret_cast<uint4*>(p.out);
          if constexpr (TF) {
#pragma unroll
            for (int k = 0; k < NC / 4; ++k) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(acc[k * 4 + e]);
              if constexpr (EPI == EPI_FWD) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = elu_fast(v[e] + bias_r[k * 4 + e]);
              } else {
                const uint32_t aw[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] *= elu_grad_from_act(__uint_as_float(aw[e]));
              }
              // stored ROUNDED to tf32: the next layer's tensor core would otherwise truncate the low 13 bits
              uint4 o;
              o.x = __float_as_uint(to_tf32(v[0])); o.y = __float_as_uint(to_tf32(v[1]));
              o.z = __float_as_uint(to_tf32(v[2])); o.w = __float_as_uint(to_tf32(v[3]));
              outp[((size_t)cn * p.nch_out + k0 + k) * plane_sz + (size_t)pix] = o;
            }
          } else {
#pragma unroll
          for (int k = 0; k < NC / 8; ++k) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc[k * 8 + e]);
            if constexpr (EPI == EPI_FWD) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = elu_fast(v[e] + bias_r[k * 8 + e]);
            } else {
              const uint32_t aw[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 a = unpack_h2(aw[e], F16);
                v[2 * e] *= elu_grad_from_act(a.x);
                v[2 * e + 1] *= elu_grad_from_act(a.y);
              }
            }
            uint4 o;
            o.x = pack_h2(v[0], v[1], F16);
            o.y = pack_h2(v[2], v[3], F16);
            o.z = pack_h2(v[4], v[5], F16);
            o.w = pack_h2(v[6], v[7], F16);
            outp[((size_t)cn * p.nch_out + k0 + k) * plane_sz + (size_t)pix] = o;
          }
          }
        }
      }
      pix = pix1;
      cn = cn1;
      if constexpr (PF3) { pix1 = pix2; pix2 = pix3; cn1 = cn2; cn2 = cn3; }
      if constexpr (DG) {
#pragma unroll
        for (int k = 0; k < NAV; ++k) {
          av[k] = av1[k];
          if constexpr (PF3) { av1[k] = av2[k]; av2[k] = av3[k]; }
        }
      }
      if (++stage == ACC) { stage = 0; sph ^= 1u; }
    }
    ds_flush();
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcGeom {             // shared-memory geometry of one kernel flavour
  int Ps, R, m, segs, TH;
  uint32_t w_bytes, plane_stride16, box_bytes;
  size_t smem;
};

struct TcState {
  TcMaps map_act[IODINE_MAX_LAYERS];
  TcMaps map_g[2];
  TcMaps map_seed;
  uint16_t* w_fwd[IODINE_MAX_LAYERS];    // layers 1..n-1
  uint16_t* w_bwd[IODINE_MAX_LAYERS];
  uint16_t* w_out = nullptr;             // decoder.conv forward, N = 16 (4 real rows)
  uint16_t* w_in4 = nullptr;             // decoder.conv data-gradient, tap pairs
  float* ptab_c = nullptr;                    // chunk-planar fp32 copy of ptab
  float* eptab_c = nullptr;                   // exp() of it (clamped to +-IOD_L1_EXP_RANGE), same layout
  int* pbig = nullptr;                        // 1 if any |ptab| exceeds that range (then tc_layer1_kernel evaluates exp)
  int n_ent_cc = 0, n_ent_in4 = 0;
  TcGeom g_cc, g_out, g_in4;
  // row-streaming variant (W == 128): weight images with the KS tap rows stacked along N, own geometry
  bool rs = false;
  uint16_t* w_fwd_rs[IODINE_MAX_LAYERS];
  uint16_t* w_bwd_rs[IODINE_MAX_LAYERS];
  uint16_t* w_out_rs = nullptr;
  void* zero_row = nullptr;              // W x 16 zero bytes
  // row-streaming work lists (device): [0] one CTA per range (grid = SMs), [1] nsplit_cc CTAs per range
  int4* itab[2] = {nullptr, nullptr};
  int32_t* coff[2] = {nullptr, nullptr};
  int rs_grid[2] = {0, 0};
  TcGeom g_cc_rs, g_out_rs;
  bool attr_done = false;
  bool tf = false;                       // tf32 operands (planes of 4 fp32 channels)
  int pw = 8;                            // channels per plane
  int nsplit_cc = 1;                     // CTAs per work range of the C->C layers (output-channel blocks)
  int nsplit_in4 = 1;                    // the same for the 4->C data-gradient (tf32 C=64: the epilogue's prefetch
                                         // queue of the saved activation would not fit the register file at N=64)
};

static TcState* tc_state(Plan* p) { return reinterpret_cast<TcState*>(p->tc); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// The chunk-planar tensor [n][plane][y][x][8 x bf16] is described to the TMA unit in 8-byte
// elements (two per pixel-plane): a 16-byte innermost box would cost one L2 request per pixel,
// 256 elements of 8 bytes fetch 128 pixels (2 KB) per request stream.
static int make_map1(CUtensorMap* map, void* base, int W, int H, int planes, int BK, int box_px) {
  EncodeTiledFn fn = get_encode_fn();
  IOD_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)planes, (cuuint64_t)BK};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)box_px * 2, 1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IOD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

static int make_map(TcMaps* maps, void* base, int W, int H, int planes, int BK, int pad) {
  const int pbox = W + 2 * pad;
  const int n_full = pbox / 128, tail = pbox % 128;
  if (make_map1(&maps->full, base, W, H, planes, BK, n_full ? 128 : tail)) return 1;
  return make_map1(&maps->tail, base, W, H, planes, BK, tail ? tail : 128);
}

static int round_up(int v, int a) { return (v + a - 1) / a * a; }

// geometry for (input planes, N, entries, max extra reach of an entry's second K half)
static bool tc_geometry(const Plan* p, int nch_in, int N, int n_ent, int max_lbo_pos, TcGeom* g, bool rs = false) {
  const IodineShape& s = p->s;
  const int pad = s.dec_k / 2;
  g->Ps = round_up((rs ? 128 : s.W) + 2 * pad, 8);            // row-streaming: the ring holds one 128-column segment
  g->segs = rs ? 1 : (s.W % 128 == 0) ? s.W / 128 : 0;
  g->w_bytes = (uint32_t)n_ent * 2u * (uint32_t)N * 16u;
  g->box_bytes = (uint32_t)(s.W + 2 * pad) * 16u;
  g->m = (127 + 2 * pad + max_lbo_pos + g->Ps - 1) / g->Ps;   // rows mirrored behind the ring end
  if (rs) g->m = 0;                                            // row-streaming tiles never leave their ring row
  const size_t row_bytes = (size_t)nch_in * g->Ps * 16;
  // (exactly what g->smem adds up below: the fp16 C = 64 ring keeps its 9th row by ~600 bytes)
  const size_t budget = (size_t)227 * 1024 - round_up((int)g->w_bytes, 1024) - sizeof(TcSmem) - 64;
  int rows = (int)(budget / row_bytes);
  int R = rows - g->m;
  if (R > TC_MAX_RING) R = TC_MAX_RING;
  const int span = g->segs ? 2 * pad + 1 : (127 + 2 * pad * g->Ps + 2 * pad) / g->Ps + 2;
  if (R < span + 1 || R <= g->m) return false;
  g->R = R;
  g->plane_stride16 = (uint32_t)(R + g->m) * g->Ps;
  if (g->plane_stride16 >= 16384u) return false;              // LBO field is 14 bits
  // rows per work item: balance the grid (items >> SMs) against halo re-reads
  int TH = 8;                                                  // largest divisor of H that is <= 8 (generic kernels; row-streaming uses the work list)
  while (s.H % TH != 0) --TH;
  g->TH = TH;
  g->smem = (size_t)round_up((int)g->w_bytes, 1024) + (size_t)nch_in * g->plane_stride16 * 16 + sizeof(TcSmem) + 64;
  return g->smem <= (size_t)227 * 1024;
}

// C->C layers: geometry for `nsplit` CTAs per work range, each computing C / nsplit output channels
static bool tc_geometry_cc(const Plan* p, int nsplit, bool rs, TcGeom* g) {
  const IodineShape& s = p->s;
  const int pw = tf_mode(p) ? 4 : 8, planes = p->C / pw, kk = s.dec_k * s.dec_k;
  const int N = p->C / nsplit;
  if (N < 16 || N % 16) return false;
  return tc_geometry(p, planes, N, kk * (planes / 2), 0, g, rs);
}
static bool tc_rs_shape(const Plan* p) {
  return p->s.W % 128 == 0 && p->s.dec_k == 3 && !getenv("IODINE_TC_NO_RS");
}
// smallest channel split whose weight image leaves room for the ring (row-streaming: >= 4 ring rows)
static int tc_pick_split(const Plan* p, bool rs, TcGeom* g) {
  for (int ns = 1; ns <= 2; ns *= 2)
    if (tc_geometry_cc(p, ns, rs, g) && (!rs || g->R >= 4)) return ns;
  return 0;
}

int tc_supported(const Plan* p) {
  const IodineShape& s = p->s;
  const int C = s.dec_chan;
  TcGeom g;
  if (C % 16 != 0) { set_error("tensor-core modes: DEC.CONV_CHAN=%d must be a multiple of 16", C); return 0; }
  if ((tc_rs_shape(p) && tc_pick_split(p, true, &g)) || tc_pick_split(p, false, &g)) return 1;
  set_error("tensor-core modes: decoder shape (C=%d, k=%d, W=%d) does not fit the tensor-core kernel's shared memory",
            C, s.dec_k, s.W);
  return 0;
}

// work list: CTA range c owns units [c U / G, (c+1) U / G) of the flattened (slot-image, segment, row) space
static int tc_build_worklist(Plan* p, TcState* st, int which, int G_max) {
  const IodineShape& s = p->s;
  const int segs = s.W / 128;
  const long long U = (long long)p->BK * segs * s.H;
  const int G = (int)(U < G_max ? U : G_max);
  std::vector<int4> items;
  std::vector<int32_t> off(G + 1);
  for (int c = 0; c < G; ++c) {
    off[c] = (int32_t)items.size();
    long long u0 = U * c / G;
    const long long u1 = U * (c + 1) / G;
    while (u0 < u1) {
      const long long col = u0 / s.H;
      const int y0 = (int)(u0 - col * s.H);
      const int th = (int)((s.H - y0 < u1 - u0) ? (s.H - y0) : (u1 - u0));
      items.push_back(make_int4((int)(col / segs), y0, th, (int)(col % segs) * 128));
      u0 += th;
    }
  }
  off[G] = (int32_t)items.size();
  st->rs_grid[which] = G;
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->itab[which], items.size() * sizeof(int4)));
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->coff[which], off.size() * sizeof(int32_t)));
  IOD_CHECK_CUDA(cudaMemcpy(st->itab[which], items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
  IOD_CHECK_CUDA(cudaMemcpy(st->coff[which], off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return 0;
}

int tc_alloc(Plan* p) {
  TcState* st = new TcState();
  p->tc = st;
  const IodineShape& s = p->s;
  const int C = p->C, kk = s.dec_k * s.dec_k, pad = s.dec_k / 2;
  st->tf = tf_mode(p);
  st->pw = st->tf ? 4 : 8;
  const int planes = C / st->pw;
  for (int l = 0; l < IODINE_MAX_LAYERS; ++l) { st->w_fwd[l] = nullptr; st->w_bwd[l] = nullptr; }
  st->n_ent_cc = kk * (planes / 2);
  st->n_ent_in4 = (kk + 1) / 2;
  const int Ps = round_up(s.W + 2 * pad, 8);
  for (int l = 0; l < IODINE_MAX_LAYERS; ++l) { st->w_fwd_rs[l] = nullptr; st->w_bwd_rs[l] = nullptr; }
  st->rs = false;
  if (tc_rs_shape(p)) {
    const int ns = tc_pick_split(p, true, &st->g_cc_rs);
    if (ns && tc_geometry(p, planes, 16, st->n_ent_cc, 0, &st->g_out_rs, true)) { st->rs = true; st->nsplit_cc = ns; }
  }
  if (!st->rs) {
    st->nsplit_cc = tc_pick_split(p, false, &st->g_cc);
    IOD_REQUIRE(st->nsplit_cc > 0, "tc geometry (C->C) failed");
    IOD_REQUIRE(tc_geometry(p, planes, 16, st->n_ent_cc, 0, &st->g_out), "tc geometry (C->4) failed");
  } else {
    st->g_cc = st->g_cc_rs;                    // (sizes the weight images; the generic C->C kernels are not launched)
    st->g_out = st->g_out_rs;
  }
  st->nsplit_in4 = (st->tf && C == 64) ? 2 : 1;
  IOD_REQUIRE(tc_geometry(p, 1, C / st->nsplit_in4, st->n_ent_in4, Ps, &st->g_in4), "tc geometry (4->C) failed");
  const size_t cc_bytes = (size_t)st->g_cc.w_bytes * st->nsplit_cc;   // one image per output-channel block
  if (st->rs) {
    for (int l = 1; l < s.dec_layers; ++l) {
      IOD_CHECK_CUDA(cudaMalloc((void**)&st->w_fwd_rs[l], cc_bytes));
      IOD_CHECK_CUDA(cudaMalloc((void**)&st->w_bwd_rs[l], cc_bytes));
    }
    IOD_CHECK_CUDA(cudaMalloc((void**)&st->w_out_rs, st->g_out_rs.w_bytes));
    IOD_CHECK_CUDA(cudaMalloc(&st->zero_row, 4096));
    IOD_CHECK_CUDA(cudaMemset(st->zero_row, 0, 4096));
    if (tc_build_worklist(p, st, 0, p->num_sms)) return 1;
    if (st->nsplit_cc > 1 && tc_build_worklist(p, st, 1, p->num_sms / st->nsplit_cc)) return 1;
  }
  for (int l = 1; l < s.dec_layers; ++l) {
    IOD_CHECK_CUDA(cudaMalloc((void**)&st->w_fwd[l], cc_bytes));
    IOD_CHECK_CUDA(cudaMalloc((void**)&st->w_bwd[l], cc_bytes));
  }
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->w_out, st->g_out.w_bytes));
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->w_in4, (size_t)st->g_in4.w_bytes * st->nsplit_in4));
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->ptab_c, (size_t)p->HW * C * sizeof(float)));
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->eptab_c, (size_t)p->HW * C * sizeof(float)));
  IOD_CHECK_CUDA(cudaMalloc((void**)&st->pbig, sizeof(int)));
  return 0;
}

void tc_free(Plan* p) {
  TcState* st = tc_state(p);
  if (!st) return;
  for (int l = 0; l < IODINE_MAX_LAYERS; ++l) {
    cudaFree(st->w_fwd[l]); cudaFree(st->w_bwd[l]); cudaFree(st->w_fwd_rs[l]); cudaFree(st->w_bwd_rs[l]);
  }
  cudaFree(st->w_out_rs);
  cudaFree(st->zero_row);
  for (int i = 0; i < 2; ++i) { cudaFree(st->itab[i]); cudaFree(st->coff[i]); }
  cudaFree(st->w_out); cudaFree(st->w_in4); cudaFree(st->ptab_c); cudaFree(st->eptab_c); cudaFree(st->pbig);
  delete st;
  p->tc = nullptr;
}

int tc_on_workspace(Plan* p) {
  TcState* st = tc_state(p);
  const IodineShape& s = p->s;
  const int pad = s.dec_k / 2, planes = p->C / st->pw;
  for (int l = 0; l < s.dec_layers; ++l)
    if (make_map(&st->map_act[l], p->act[l], s.W, s.H, planes, p->BK, pad)) return 1;
  if (s.dec_layers > 1)
    for (int i = 0; i < 2; ++i)
      if (make_map(&st->map_g[i], p->gbuf[i], s.W, s.H, planes, p->BK, pad)) return 1;
  if (make_map(&st->map_seed, p->seed4, s.W, s.H, 1, p->BK, pad)) return 1;
  return 0;
}

// ---- weight packing -----------------------------------------------------------------------------
// image layout (= the shared-memory image): [entry][k-half (2)][n (N)][8 k-elements] bf16.
// C->C: entry = tap*(C/16) + ks, k-half kc covers input channels (2ks+kc)*8 .. +7.
//   fwd : B[n=co][k=ci]  = W[co][ci][dy][dx]
//   bwd : B[n=ci][k=co]  = W[co][ci][k-1-dy][k-1-dx]     (data-gradient = conv with flipped taps)
__device__ __forceinline__ uint16_t to_h(float v, int f16) {
  if (f16) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float from_h(uint16_t u, int f16) {
  if (f16) return __half2float(*reinterpret_cast<__half*>(&u));
  return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
}

// One kernel for every C -> X weight image.  rs = 0: entry = tap*nks + ks (image [entry][k-half][n][PW]);
// rs = 1: entry = dx*nks + ks and the KS tap rows stacked along n (image [entry][k-half][b*N + n][PW], block b <-> tap
// row dy = KS-1-b: the oldest of the KS output rows an input row feeds comes first).  PW = 8 16-bit elements or 4 tf32
// elements per (n, k-half); nks = C / (2 PW) K-steps per tap.  `nsplit` images follow each other, image h holding the
// output channels h*N .. h*N+N-1 (forward: co, data-gradient: ci).
__global__ void tc_pack_kernel(const float* __restrict__ w, void* __restrict__ fwd, void* __restrict__ bwd, int C, int NO,
                               int N, int nsplit, int KS, int rs, int tf, int f16) {
  const int PW = tf ? 4 : 8, nks = C / (2 * PW);
  const int per = KS * KS * nks * 2 * N * PW;
  const int total = per * nsplit;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
    const int h = j / per, i = j - h * per;
    const int e = i % PW, n = (i / PW) % N;
    int kc, ks, dy, dx;
    if (rs) {
      const int b = (i / (PW * N)) % KS;
      kc = (i / (PW * N * KS)) % 2; ks = (i / (2 * PW * N * KS)) % nks; dx = i / (2 * PW * N * KS * nks);
      dy = KS - 1 - b;
    } else {
      kc = (i / (PW * N)) % 2; ks = (i / (2 * PW * N)) % nks;
      const int tap = i / (2 * PW * N * nks);
      dy = tap / KS; dx = tap % KS;
    }
    const int k = (2 * ks + kc) * PW + e, ng = h * N + n;
    if (fwd) {
      const float v = ng < NO ? w[(((size_t)ng * C + k) * KS + dy) * KS + dx] : 0.f;
      if (tf) reinterpret_cast<float*>(fwd)[j] = to_tf32(v);
      else reinterpret_cast<uint16_t*>(fwd)[j] = to_h(v, f16);
    }
    if (bwd) {
      const float v = w[(((size_t)k * C + ng) * KS + (KS - 1 - dy)) * KS + (KS - 1 - dx)];
      if (tf) reinterpret_cast<float*>(bwd)[j] = to_tf32(v);
      else reinterpret_cast<uint16_t*>(bwd)[j] = to_h(v, f16);
    }
  }
}
// 4->C data-gradient of decoder.conv [4][C][k][k]; entry = tap pair (2q, 2q+1), one tap per K half; the odd tail
// reuses the previous tap with zero weights in its first half.  B[n=ci][k=o] (o < 4 real; 16-bit planes carry 4 zeros).
__global__ void tc_pack_in4_kernel(const float* __restrict__ w, void* __restrict__ img, int C, int N, int KS, int tf, int f16) {
  const int PW = tf ? 4 : 8;
  const int kk = KS * KS, npair = (kk + 1) / 2;
  const int per = npair * 2 * N * PW;                 // one image per block of N output channels
  const int total = per * (C / N);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
    const int h = j / per, i = j - h * per;
    const int e = i % PW, n = h * N + (i / PW) % N, kc = (i / (PW * N)) % 2, q = i / (2 * PW * N);
    int tap = 2 * q + kc;
    bool zero = false;
    if (2 * q + 1 >= kk) {            // odd tail: (kk-2, kk-1) with the first half zeroed
      tap = kk - 2 + kc;
      zero = (kc == 0);
    }
    const int dy = tap / KS, dx = tap % KS;
    float v = 0.f;
    if (!zero && e < 4) v = w[(((size_t)e * C + n) * KS + (KS - 1 - dy)) * KS + (KS - 1 - dx)];
    if (tf) reinterpret_cast<float*>(img)[j] = to_tf32(v);
    else reinterpret_cast<uint16_t*>(img)[j] = to_h(v, f16);
  }
}
// eptab_c = exp(clamp(ptab_c)); *pbig = 1 if the clamp was ever active
__global__ void tc_ptab_exp_kernel(const float* __restrict__ ptab_c, float* __restrict__ eptab_c, int* __restrict__ pbig, int n) {
  int big = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = ptab_c[i];
    eptab_c[i] = expf(fminf(fmaxf(v, -IOD_L1_EXP_RANGE), IOD_L1_EXP_RANGE));
    big |= !(fabsf(v) <= IOD_L1_EXP_RANGE);
  }
  if (__syncthreads_or(big) && threadIdx.x == 0) atomicOr(pbig, 1);
}

__global__ void tc_ptab_planar_kernel(const float* __restrict__ ptab, float* __restrict__ ptab_c, int HW, int C, int PW) {
  const int total = HW * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i % PW, pix = (i / PW) % HW, k = i / (PW * HW);
    ptab_c[i] = ptab[(size_t)pix * C + k * PW + e];
  }
}

int tc_setup_weights(Plan* p, const IodineWeights* w, cudaStream_t st_) {
  TcState* st = tc_state(p);
  const IodineShape& s = p->s;
  const int C = p->C, KS = s.dec_k, f16 = half_is_f16(p), tf = st->tf ? 1 : 0, ns = st->nsplit_cc, Ncc = C / ns;
  if (!st->rs)
    for (int l = 1; l < s.dec_layers; ++l) {
      tc_pack_kernel<<<64, 256, 0, st_>>>(w->dec_w[l], st->w_fwd[l], st->w_bwd[l], C, C, Ncc, ns, KS, 0, tf, f16);
      IOD_LAUNCH_CHECK(p);
    }
  tc_pack_kernel<<<16, 256, 0, st_>>>(w->dec_out_w, st->w_out, nullptr, C, 4, 16, 1, KS, 0, tf, f16);
  IOD_LAUNCH_CHECK(p);
  tc_pack_in4_kernel<<<16, 256, 0, st_>>>(w->dec_out_w, st->w_in4, C, C / st->nsplit_in4, KS, tf, f16);
  IOD_LAUNCH_CHECK(p);
  if (st->rs) {
    for (int l = 1; l < s.dec_layers; ++l) {
      tc_pack_kernel<<<64, 256, 0, st_>>>(w->dec_w[l], st->w_fwd_rs[l], st->w_bwd_rs[l], C, C, Ncc, ns, KS, 1, tf, f16);
      IOD_LAUNCH_CHECK(p);
    }
    tc_pack_kernel<<<16, 256, 0, st_>>>(w->dec_out_w, st->w_out_rs, nullptr, C, 4, 16, 1, KS, 1, tf, f16);
    IOD_LAUNCH_CHECK(p);
  }
  tc_ptab_planar_kernel<<<256, 256, 0, st_>>>(p->ptab, st->ptab_c, p->HW, C, st->pw);   // after pack_ptab (same stream)
  IOD_LAUNCH_CHECK(p);
  IOD_CHECK_CUDA(cudaMemsetAsync(st->pbig, 0, sizeof(int), st_));
  tc_ptab_exp_kernel<<<256, 256, 0, st_>>>(st->ptab_c, st->eptab_c, st->pbig, p->HW * C);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// ---- launch -------------------------------------------------------------------------------------
static uint32_t make_idesc(int N, int fmt) {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format @7/@10 (F16F32Format: 0 = F16, 1 = BF16, 2 = TF32),
  // K-major A and B, n_dim = N>>3 @17, m_dim = 128>>4 @24
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
static int operand_fmt(const Plan* p) { return tf_mode(p) ? 2 : (p->s.precision == IODINE_FP16 ? 0 : 1); }

// N = output channels computed by one CTA, n_tot = output channels of the layer (nsplit = n_tot / N CTAs per range)
static void fill_common(const Plan* p, const TcGeom& g, int nch_in, int N, int n_tot, TcParams* q) {
  const IodineShape& s = p->s;
  const int pw = tf_mode(p) ? 4 : 8;
  q->nch_in = nch_in;
  q->nch_out = n_tot / pw;
  q->n_tot = n_tot;
  q->nsplit = n_tot / N;
  q->H = s.H; q->W = s.W; q->pad = s.dec_k / 2;
  q->Ps = g.Ps; q->R = g.R; q->m = g.m; q->segs = g.segs; q->TH = g.TH;
  q->strips = (s.H + g.TH - 1) / g.TH;
  q->rs_segs = 0;
  // rows per issue unit.  Short rings (tf32: 4 rows of 34.8 KB) hand rows over one at a time now that the scout warp
  // takes the waits: a ring row is requested again as soon as its own MMAs have completed (52.7k against 50.8k steps/s
  // with units of two rows, profiles/r2_issuer_experiments.md)
  q->rs_unit = g.R >= 7 ? TC_RS_UNIT : 1;
  if (getenv("IODINE_TC_RS_UNIT")) {                       // experiment: rows per issue unit
    const int u = atoi(getenv("IODINE_TC_RS_UNIT"));
    // below the ring depth, and a unit's fresh accumulators + the KS-1 still open ones must fit the slot ring (the scout
    // waits for all of a unit's slots before releasing it: a larger unit would wait for rows it has not released yet)
    const int acc = N <= 32 ? TC_RS_SLOTS_MAX : TC_RS_SLOTS;
    if (u >= 1 && u < g.R && u <= 16 && u + s.dec_k - 1 <= acc - 1) q->rs_unit = u;
  }
  q->rev = 0;
  q->itab = nullptr;
  q->coff = nullptr;
  q->G = nullptr;
  q->apf = 0;
  q->lead = getenv("IODINE_TC_LEAD") ? atoi(getenv("IODINE_TC_LEAD")) : 0;
  q->items = p->BK * q->strips;
  const int fmt = operand_fmt(p);
  q->idesc = make_idesc(N, fmt);
  for (int c = 0; c < 8; ++c) q->idesc_n[c] = (c >= 1 && c * N <= 256) ? make_idesc(c * N, fmt) : 0u;
  q->f16 = s.precision == IODINE_FP16;
  {
    static const int dbg = getenv("IODINE_TC_DEBUG") ? atoi(getenv("IODINE_TC_DEBUG")) : 0;
    q->dbg = dbg;
  }
  q->w_bytes = g.w_bytes;
  q->box_bytes = g.box_bytes;
  q->n_full = (s.W + 2 * q->pad) / 128;
  q->tail_px = (s.W + 2 * q->pad) % 128;
  q->plane_stride16 = g.plane_stride16;
  q->NT = g.segs ? g.TH * g.segs : ((g.TH - 1) * g.Ps + s.W - 1) / 128 + 1;
  const int nks = nch_in / 2, ks_dim = s.dec_k;
  for (int i = 0; i < 40; ++i) q->a_off[i] = 0;
  if (nks > 0)
    for (int dx = 0; dx < ks_dim; ++dx)
      for (int ks = 0; ks < nks; ++ks)
        if (dx * nks + ks < 40) q->a_off[dx * nks + ks] = (uint32_t)dx + (uint32_t)(2 * ks) * g.plane_stride16;
}

// dispatch on the compile-time shape <N, EPI, KS, NKS>
template <int N, int EPI, int KS, int NKS, int PS, int PST16, bool RS = false, bool TF = false>
static int tc_launch_g(Plan* p, const TcParams& q, size_t smem, cudaStream_t st_) {
  auto kern = conv_tc_kernel<N, EPI, KS, NKS, PS, PST16, RS, TF>;
  static bool attr_done = false;
  if (!attr_done) {
    IOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  const int ranges_max = p->num_sms / q.nsplit;
  const int ranges = q.items < ranges_max ? q.items : ranges_max;
  kern<<<ranges * q.nsplit, tc_threads(N, RS, tc_iss2(EPI, RS, TF)), smem, st_>>>(q);
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// Geometry-specialised instantiations for the shapes the reference's configs use (ring pitch and plane
// stride baked in as immediates); any other shape runs the generic kernel.  The constants are what
// tc_geometry() yields for: W=128 k=3 C=64 (configs/clevr6*.yaml) and W=64 k=3 C=32 (configs/dsprites*.yaml).
template <int N, int EPI, int KS, int NKS>
static int tc_launch_k(Plan* p, const TcParams& q, size_t smem, cudaStream_t st_) {
  const int ps = q.Ps, pst = (int)q.plane_stride16;
  if constexpr (KS == 3 && NKS == 4 && N == 64) { if (ps == 136 && pst == 1224) return tc_launch_g<N, EPI, KS, NKS, 136, 1224>(p, q, smem, st_); }
  if constexpr (KS == 3 && NKS == 4 && N == 16) { if (ps == 136 && pst == 1632) return tc_launch_g<N, EPI, KS, NKS, 136, 1632>(p, q, smem, st_); }
  if constexpr (KS == 3 && NKS == 0 && N == 64) { if (ps == 136 && pst == 4624) return tc_launch_g<N, EPI, KS, NKS, 136, 4624>(p, q, smem, st_); }
  if constexpr (KS == 3 && NKS == 2 && (N == 32 || N == 16)) { if (ps == 72 && pst == 2448) return tc_launch_g<N, EPI, KS, NKS, 72, 2448>(p, q, smem, st_); }
  if constexpr (KS == 3 && NKS == 0 && N == 32) { if (ps == 72 && pst == 2520) return tc_launch_g<N, EPI, KS, NKS, 72, 2520>(p, q, smem, st_); }
  if (getenv("IODINE_TC_VERBOSE")) fprintf(stderr, "conv_tc: generic kernel for N=%d EPI=%d KS=%d NKS=%d (Ps=%d, plane stride %d)\n", N, EPI, KS, NKS, ps, pst);
  return tc_launch_g<N, EPI, KS, NKS, 0, 0>(p, q, smem, st_);
}
// tf32 instantiations (run-time geometry; NKS = C/8 K-steps of two 4-channel planes)
template <int N, int EPI, int KS, int NKS>
static int tc_launch_tf(Plan* p, const TcParams& q, size_t smem, cudaStream_t st_) {
  if (getenv("IODINE_TC_VERBOSE")) fprintf(stderr, "conv_tc: tf32 kernel N=%d EPI=%d KS=%d NKS=%d nsplit=%d (Ps=%d, plane stride %d, R=%d)\n", N, EPI, KS, NKS, q.nsplit, q.Ps, (int)q.plane_stride16, q.R);
  return tc_launch_g<N, EPI, KS, NKS, 0, 0, false, true>(p, q, smem, st_);
}

// row-streaming variant (W == 128, 3x3): C -> C layers and decoder.conv forward
template <int EPI>
static int tc_launch_rs(Plan* p, const TcParams& q, size_t smem, bool out4, cudaStream_t st_) {
  const int C = p->C, pst = (int)q.plane_stride16;
  if (tf_mode(p)) {
    if (!out4) {
      if (C == 64 && q.nsplit == 2) return pst == 544 ? tc_launch_g<32, EPI, 3, 8, 136, 544, true, true>(p, q, smem, st_)
                                                      : tc_launch_g<32, EPI, 3, 8, 136, 0, true, true>(p, q, smem, st_);
      if (C == 32 && q.nsplit == 1) return tc_launch_g<32, EPI, 3, 4, 136, 0, true, true>(p, q, smem, st_);
      if (C == 16 && q.nsplit == 1) return tc_launch_g<16, EPI, 3, 2, 136, 0, true, true>(p, q, smem, st_);
    } else {
      if (C == 64) return tc_launch_g<16, EPI_OUT4, 3, 8, 136, 0, true, true>(p, q, smem, st_);
      if (C == 32) return tc_launch_g<16, EPI_OUT4, 3, 4, 136, 0, true, true>(p, q, smem, st_);
      if (C == 16) return tc_launch_g<16, EPI_OUT4, 3, 2, 136, 0, true, true>(p, q, smem, st_);
    }
    set_error("conv_tc (row-streaming, tf32): unsupported C=%d / split %d", C, q.nsplit);
    return 1;
  }
  if (!out4) {
    if (C == 64) return pst == 1224 ? tc_launch_g<64, EPI, 3, 4, 136, 1224, true>(p, q, smem, st_)
                                    : tc_launch_g<64, EPI, 3, 4, 136, 0, true>(p, q, smem, st_);
    if (C == 32) return tc_launch_g<32, EPI, 3, 2, 136, 0, true>(p, q, smem, st_);
    if (C == 16) return tc_launch_g<16, EPI, 3, 1, 136, 0, true>(p, q, smem, st_);
  } else {
    if (C == 64) return pst == 1632 ? tc_launch_g<16, EPI_OUT4, 3, 4, 136, 1632, true>(p, q, smem, st_)
                                    : tc_launch_g<16, EPI_OUT4, 3, 4, 136, 0, true>(p, q, smem, st_);
    if (C == 32) return tc_launch_g<16, EPI_OUT4, 3, 2, 136, 0, true>(p, q, smem, st_);
    if (C == 16) return tc_launch_g<16, EPI_OUT4, 3, 1, 136, 0, true>(p, q, smem, st_);
  }
  set_error("conv_tc (row-streaming): unsupported C=%d", C);
  return 1;
}

// C -> C layers (forward / data-gradient): N = C / nsplit, NKS = K-steps per tap
template <int EPI>
static int tc_launch_cc(Plan* p, const TcParams& q, size_t smem, cudaStream_t st_) {
  const int C = p->C, KS = p->s.dec_k;
  if (tf_mode(p)) {
    if (KS == 3 && C == 64 && q.nsplit == 2) return tc_launch_tf<32, EPI, 3, 8>(p, q, smem, st_);
    if (KS == 3 && C == 32 && q.nsplit == 1) return tc_launch_tf<32, EPI, 3, 4>(p, q, smem, st_);
    if (KS == 3 && C == 16 && q.nsplit == 1) return tc_launch_tf<16, EPI, 3, 2>(p, q, smem, st_);
    if (KS == 5 && C == 32 && q.nsplit == 1) return tc_launch_tf<32, EPI, 5, 4>(p, q, smem, st_);
    if (KS == 5 && C == 16 && q.nsplit == 1) return tc_launch_tf<16, EPI, 5, 2>(p, q, smem, st_);
    set_error("conv_tc (tf32): unsupported C=%d k=%d split %d", C, KS, q.nsplit);
    return 1;
  }
  if (q.nsplit != 1) { set_error("conv_tc: channel split %d is a tf32-only geometry", q.nsplit); return 1; }
  if (KS == 3 && C == 64) return tc_launch_k<64, EPI, 3, 4>(p, q, smem, st_);
  if (KS == 3 && C == 32) return tc_launch_k<32, EPI, 3, 2>(p, q, smem, st_);
  if (KS == 3 && C == 16) return tc_launch_k<16, EPI, 3, 1>(p, q, smem, st_);
  if (KS == 5 && C == 32) return tc_launch_k<32, EPI, 5, 2>(p, q, smem, st_);
  if (KS == 5 && C == 16) return tc_launch_k<16, EPI, 5, 1>(p, q, smem, st_);
  set_error("conv_tc: unsupported C=%d k=%d", C, KS);
  return 1;
}
// decoder.conv forward: N = 16 (4 real outputs)
static int tc_launch_o4(Plan* p, const TcParams& q, size_t smem, cudaStream_t st_) {
  const int C = p->C, KS = p->s.dec_k;
  if (tf_mode(p)) {
    if (KS == 3 && C == 64) return tc_launch_tf<16, EPI_OUT4, 3, 8>(p, q, smem, st_);
    if (KS == 3 && C == 32) return tc_launch_tf<16, EPI_OUT4, 3, 4>(p, q, smem, st_);
    if (KS == 3 && C == 16) return tc_launch_tf<16, EPI_OUT4, 3, 2>(p, q, smem, st_);
    if (KS == 5 && C == 32) return tc_launch_tf<16, EPI_OUT4, 5, 4>(p, q, smem, st_);
    if (KS == 5 && C == 16) return tc_launch_tf<16, EPI_OUT4, 5, 2>(p, q, smem, st_);
    set_error("conv_tc (tf32): unsupported C=%d k=%d", C, KS);
    return 1;
  }
  if (KS == 3 && C == 64) return tc_launch_k<16, EPI_OUT4, 3, 4>(p, q, smem, st_);
  if (KS == 3 && C == 32) return tc_launch_k<16, EPI_OUT4, 3, 2>(p, q, smem, st_);
  if (KS == 3 && C == 16) return tc_launch_k<16, EPI_OUT4, 3, 1>(p, q, smem, st_);
  if (KS == 5 && C == 32) return tc_launch_k<16, EPI_OUT4, 5, 2>(p, q, smem, st_);
  if (KS == 5 && C == 16) return tc_launch_k<16, EPI_OUT4, 5, 1>(p, q, smem, st_);
  set_error("conv_tc: unsupported C=%d k=%d", C, KS);
  return 1;
}
// decoder.conv data-gradient: one input plane (the 4 gradient channels), tap pairs (NKS = 0), N = C
static int tc_launch_i4(Plan* p, const TcParams& q, size_t smem, cudaStream_t st_) {
  const int C = p->C, KS = p->s.dec_k;
  if (tf_mode(p)) {
    if (KS == 3 && C == 64 && q.nsplit == 2) return tc_launch_tf<32, EPI_DGRAD, 3, 0>(p, q, smem, st_);
    if (KS == 3 && C == 32) return tc_launch_tf<32, EPI_DGRAD, 3, 0>(p, q, smem, st_);
    if (KS == 3 && C == 16) return tc_launch_tf<16, EPI_DGRAD, 3, 0>(p, q, smem, st_);
    if (KS == 5 && C == 32) return tc_launch_tf<32, EPI_DGRAD, 5, 0>(p, q, smem, st_);
    if (KS == 5 && C == 16) return tc_launch_tf<16, EPI_DGRAD, 5, 0>(p, q, smem, st_);
    set_error("conv_tc (tf32): unsupported C=%d k=%d", C, KS);
    return 1;
  }
  if (KS == 3 && C == 64) return tc_launch_k<64, EPI_DGRAD, 3, 0>(p, q, smem, st_);
  if (KS == 3 && C == 32) return tc_launch_k<32, EPI_DGRAD, 3, 0>(p, q, smem, st_);
  if (KS == 3 && C == 16) return tc_launch_k<16, EPI_DGRAD, 3, 0>(p, q, smem, st_);
  if (KS == 5 && C == 32) return tc_launch_k<32, EPI_DGRAD, 5, 0>(p, q, smem, st_);
  if (KS == 5 && C == 16) return tc_launch_k<16, EPI_DGRAD, 5, 0>(p, q, smem, st_);
  set_error("conv_tc: unsupported C=%d k=%d", C, KS);
  return 1;
}

static const TcMaps* find_map(Plan* p, const void* buf) {
  TcState* st = tc_state(p);
  for (int l = 0; l < p->s.dec_layers; ++l)
    if (buf == p->act[l]) return &st->map_act[l];
  for (int i = 0; i < 2; ++i)
    if (buf == p->gbuf[i]) return &st->map_g[i];
  if (buf == (const void*)p->seed4) return &st->map_seed;
  return nullptr;
}

// 1 when the last data-gradient can reduce its output over the border classes itself (EPI_DSUM)
int tc_dsum_fused(const Plan* p) {
  const TcState* st = reinterpret_cast<const TcState*>(p->tc);
  return st && st->rs && p->s.dec_k == 3 && !getenv("IODINE_TC_NO_DSUM");
}

static void use_worklist(const Plan* p, const TcState* st, int which, TcParams* q) {
  q->rs_segs = p->s.W / 128;
  q->items = st->rs_grid[which];
  q->itab = st->itab[which];
  q->coff = st->coff[which];
}

int tc_launch_conv(Plan* p, int layer, bool dgrad, const void* in, const void* act_prev, void* out, float* G,
                   cudaStream_t st_) {
  TcState* st = tc_state(p);
  const TcMaps* map = find_map(p, in);
  IOD_REQUIRE(map != nullptr, "conv_tc: source buffer has no tensor map");
  IOD_REQUIRE(!G || (dgrad && tc_dsum_fused(p)), "conv_tc: fused class sums need the row-streaming data-gradient");
  TcParams q;
  q.maps = *map;
  fill_common(p, st->rs ? st->g_cc_rs : st->g_cc, p->C / st->pw, p->C / st->nsplit_cc, p->C, &q);
  if (st->rs) use_worklist(p, st, st->nsplit_cc > 1 ? 1 : 0, &q);
  q.G = G;
  q.apf = (dgrad && st->rs && !getenv("IODINE_TC_NO_APF")) ? 1 : 0;
  // traversal direction: read a buffer in the opposite direction to the one it was written in (the collapsed
  // first layer and the 4->C data-gradient write ascending), so that the freshest ~100 MB are still L2 hits
  if (!getenv("IODINE_TC_NO_REV")) q.rev = dgrad ? ((p->s.dec_layers - 1 - layer) % 2 == 0) : (layer % 2 == 1);
  q.wimg = st->rs ? (dgrad ? st->w_bwd_rs[layer] : st->w_fwd_rs[layer]) : (dgrad ? st->w_bwd[layer] : st->w_fwd[layer]);
  q.bias = dgrad ? nullptr : p->dec[layer].b;
  q.actp = reinterpret_cast<const uint4*>(act_prev);
  q.out = out;
  q.in = reinterpret_cast<const uint4*>(in);
  q.zero_row = reinterpret_cast<const uint4*>(st->zero_row);
  if (st->rs) {
    if (dgrad && G) return tc_launch_rs<EPI_DSUM>(p, q, st->g_cc_rs.smem, false, st_);
    if (dgrad) return tc_launch_rs<EPI_DGRAD>(p, q, st->g_cc_rs.smem, false, st_);
    return tc_launch_rs<EPI_FWD>(p, q, st->g_cc_rs.smem, false, st_);
  }
  if (dgrad) return tc_launch_cc<EPI_DGRAD>(p, q, st->g_cc.smem, st_);
  return tc_launch_cc<EPI_FWD>(p, q, st->g_cc.smem, st_);
}

int tc_launch_out4(Plan* p, const void* in, float* out4, cudaStream_t st_) {
  TcState* st = tc_state(p);
  const TcMaps* map = find_map(p, in);
  IOD_REQUIRE(map != nullptr, "conv_tc: source buffer has no tensor map");
  TcParams q;
  q.maps = *map;
  fill_common(p, st->rs ? st->g_out_rs : st->g_out, p->C / st->pw, 16, 16, &q);
  if (st->rs) use_worklist(p, st, 0, &q);
  if (!getenv("IODINE_TC_NO_REV")) q.rev = ((p->s.dec_layers - 1) % 2 == 0);   // opposite to the last forward layer (or to layer 1)
  q.wimg = st->rs ? st->w_out_rs : st->w_out;
  q.bias = p->out_b;
  q.actp = nullptr;
  q.out = out4;
  q.in = reinterpret_cast<const uint4*>(in);
  q.zero_row = reinterpret_cast<const uint4*>(st->zero_row);
  if (st->rs) return tc_launch_rs<EPI_OUT4>(p, q, st->g_out_rs.smem, true, st_);
  return tc_launch_o4(p, q, st->g_out.smem, st_);
}

int tc_launch_dgrad_in4(Plan* p, const float* seed8, const void* act_prev, void* gout, cudaStream_t st_) {
  TcState* st = tc_state(p);
  (void)seed8;
  TcParams q;
  q.maps = st->map_seed;
  fill_common(p, st->g_in4, 1, p->C / st->nsplit_in4, p->C, &q);
  q.wimg = st->w_in4;
  q.bias = nullptr;
  q.actp = reinterpret_cast<const uint4*>(act_prev);
  q.out = gout;
  return tc_launch_i4(p, q, st->g_in4.smem, st_);
}

// the row-streaming work lists ([0]: one CTA per range, [1]: nsplit_cc CTAs per range; cut at image boundaries) for other row-streaming kernels
// (wgrad_tc.cu); returns 0 when this plan has none
int tc_rs_worklist(const Plan* p, int which, const int4** itab, const int32_t** coff, int* grid, const void** zero_row) {
  const TcState* st = reinterpret_cast<const TcState*>(p->tc);
  if (!st || !st->rs || !st->itab[which] || p->s.W != 128) return 0;
  if (itab) *itab = st->itab[which];
  if (coff) *coff = st->coff[which];
  if (grid) *grid = st->rs_grid[which];
  if (zero_row) *zero_row = st->zero_row;
  return 1;
}

// ---- helpers around the chunk-planar layout -------------------------------------------------------
// act0[n][k][y][x][8] = ELU(u[n][class(y,x)][co] + ptab_c[k][y][x][8])   (first decoder layer, collapsed)
// A thread owns one (pixel, 8-channel plane) and walks TC_L1_NB slot-images with the coordinate-table
// values in registers (the table is read once per group of images, not once per image).
constexpr int TC_L1_NB = 32;             // slots per thread: the pixel's table entries (64 B from L2) are read once per TC_L1_NB stores of 16 B
// mode: 0 = bf16 planes of 8, 1 = fp16 planes of 8, 2 = tf32 planes of 4 (two planes per thread, same 8 channels)
// ELU needs exp(u + p) for the non-positive pre-activations.  u depends on (slot, border class, channel), p on (pixel,
// channel): exp(u) (sample_u_kernel) and exp(p) (tc_ptab_exp_kernel, once per weight load) are tables, and the kernel
// multiplies -- 4 instructions per value instead of 7 and no MUFU (235 M exponentials per launch at 16 per clock and SM
// were 52 us of the kernel's 138).  Slots / weights whose |u| / |p| leave +-IOD_L1_EXP_RANGE take the evaluated form.
template <int mode>
__global__ void __launch_bounds__(256, 4)
tc_layer1_kernel(const float* __restrict__ u, const float* __restrict__ eu, const int* __restrict__ ubig,
                 const float* __restrict__ ptab_c, const float* __restrict__ eptab_c, const int* __restrict__ pbig,
                 uint4* __restrict__ act0, int H, int W, int C, int KS, int BK) {
  const int k = blockIdx.y;                            // group of 8 channels
  const int P = KS / 2, HW = H * W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const int y = pix / W, x = pix - y * W;
  const int cls = border_class(y, H, P) * KS + border_class(x, W, P);
  float4 p0, p1, e0, e1;
  if (mode == 2) {                                     // ptab_c is [C/4][HW][4]
    const size_t i0 = ((size_t)(2 * k) * HW + pix) * 4, i1 = ((size_t)(2 * k + 1) * HW + pix) * 4;
    p0 = __ldg(reinterpret_cast<const float4*>(ptab_c + i0));
    p1 = __ldg(reinterpret_cast<const float4*>(ptab_c + i1));
    e0 = __ldg(reinterpret_cast<const float4*>(eptab_c + i0));
    e1 = __ldg(reinterpret_cast<const float4*>(eptab_c + i1));
  } else {                                             // [C/8][HW][8]
    const size_t i0 = ((size_t)k * HW + pix) * 8;
    const float4* pv = reinterpret_cast<const float4*>(ptab_c + i0);
    const float4* ev = reinterpret_cast<const float4*>(eptab_c + i0);
    p0 = __ldg(pv); p1 = __ldg(pv + 1);
    e0 = __ldg(ev); e1 = __ldg(ev + 1);
  }
  const int n0 = blockIdx.z * TC_L1_NB;
  const int n1 = (n0 + TC_L1_NB < BK) ? n0 + TC_L1_NB : BK;
  // one decision for the block's slots, taken before the loop (a per-slot test would put a dependent load and a branch
  // in front of every iteration's loads)
  bool big = __ldg(pbig) != 0;
  for (int n = n0; n < n1; ++n) big |= __ldg(ubig + n) != 0;
  // pointers stepped per slot (the index arithmetic of the straightforward form was half of the kernel's instructions)
  const size_t u_step = (size_t)KS * KS * C / 4;                       // float4 per slot
  const float4* up = reinterpret_cast<const float4*>(u + ((size_t)n0 * KS * KS + cls) * C + k * 8);
  const float4* xp = reinterpret_cast<const float4*>(eu + ((size_t)n0 * KS * KS + cls) * C + k * 8);
  const size_t o_step = (size_t)(mode == 2 ? C / 4 : C / 8) * HW;      // uint4 per slot
  uint4* op = act0 + ((size_t)n0 * (mode == 2 ? C / 4 : C / 8) + (mode == 2 ? 2 * k : k)) * HW + pix;
  auto store = [&](float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
    if (mode == 2) {
      op[0] = make_uint4(__float_as_uint(to_tf32(v0)), __float_as_uint(to_tf32(v1)), __float_as_uint(to_tf32(v2)), __float_as_uint(to_tf32(v3)));
      op[HW] = make_uint4(__float_as_uint(to_tf32(v4)), __float_as_uint(to_tf32(v5)), __float_as_uint(to_tf32(v6)), __float_as_uint(to_tf32(v7)));
    } else {
      uint4 o;
      o.x = pack_h2(v0, v1, mode); o.y = pack_h2(v2, v3, mode); o.z = pack_h2(v4, v5, mode); o.w = pack_h2(v6, v7, mode);
      op[0] = o;
    }
    op += o_step;
  };
  if (big) {
#pragma unroll 4
    for (int n = n0; n < n1; ++n) {
      const float4 u0 = __ldg(up), u1 = __ldg(up + 1);
      up += u_step;
      // hardware exponential: |error| ~1e-7 absolute, far below the operand rounding that follows
      store(elu_fast(u0.x + p0.x), elu_fast(u0.y + p0.y), elu_fast(u0.z + p0.z), elu_fast(u0.w + p0.w),
            elu_fast(u1.x + p1.x), elu_fast(u1.y + p1.y), elu_fast(u1.z + p1.z), elu_fast(u1.w + p1.w));
    }
    return;
  }
  auto elu2 = [](float uu, float pp, float eu_, float ep_) { const float pre = uu + pp; return pre > 0.f ? pre : fmaf(eu_, ep_, -1.f); };
#pragma unroll 4
  for (int n = n0; n < n1; ++n) {
    const float4 u0 = __ldg(up), u1 = __ldg(up + 1), x0 = __ldg(xp), x1 = __ldg(xp + 1);
    up += u_step; xp += u_step;
    store(elu2(u0.x, p0.x, x0.x, e0.x), elu2(u0.y, p0.y, x0.y, e0.y), elu2(u0.z, p0.z, x0.z, e0.z), elu2(u0.w, p0.w, x0.w, e0.w),
          elu2(u1.x, p1.x, x1.x, e1.x), elu2(u1.y, p1.y, x1.y, e1.y), elu2(u1.z, p1.z, x1.z, e1.z), elu2(u1.w, p1.w, x1.w, e1.w));
  }
}

static int plane_mode(const Plan* p) { return tf_mode(p) ? 2 : (p->s.precision == IODINE_FP16 ? 1 : 0); }

int tc_launch_layer1(Plan* p, void* act0, cudaStream_t st_) {
  TcState* st = tc_state(p);
  dim3 grid((p->HW + 255) / 256, p->C / 8, (p->BK + TC_L1_NB - 1) / TC_L1_NB);
#define IOD_L1(m) tc_layer1_kernel<m><<<grid, 256, 0, st_>>>(p->u, p->eu, p->ubig, st->ptab_c, st->eptab_c, st->pbig, \
                                                      reinterpret_cast<uint4*>(act0), p->s.H, p->s.W, p->C, p->s.dec_k, p->BK)
  const int pm = plane_mode(p);
  if (pm == 2) IOD_L1(2); else if (pm == 1) IOD_L1(1); else IOD_L1(0);
#undef IOD_L1
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// one plane value (16 bytes) -> up to 8 floats; returns the number of channels it holds
__device__ __forceinline__ int plane_unpack(const uint4& v, int mode, float* o) {
  if (mode == 2) {
    o[0] = __uint_as_float(v.x); o[1] = __uint_as_float(v.y); o[2] = __uint_as_float(v.z); o[3] = __uint_as_float(v.w);
    o[4] = o[5] = o[6] = o[7] = 0.f;
    return 4;
  }
  const float2 a = unpack_h2(v.x, mode), b = unpack_h2(v.y, mode), c = unpack_h2(v.z, mode), d = unpack_h2(v.w, mode);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
  return 8;
}

// G[n][class][co] = sum over pixels of the class of g[n][plane][y][x][co in plane]  (layer-1 dgrad collapse)
// The block walks its rows in maximal runs of equal row class.  Inside a run a thread owns column
// x = tid % min(W,256) and every (256/W)-th row, four loads in flight; its column class is fixed, so
// interior columns are block-reduced once per run and the 2*(k/2) border columns flush their own sums:
// a handful of atomics per (block, run) instead of one per border pixel.
__global__ void __launch_bounds__(256)
tc_class_sum_kernel(const uint4* __restrict__ g, float* __restrict__ G, int H, int W, int C, int KS, int rows_per_block,
                    int mode) {
  const int PW = mode == 2 ? 4 : 8;
  const int n = blockIdx.z, k = blockIdx.y;            // k: plane
  const int P = KS / 2, HW = H * W;
  const int y_lo = blockIdx.x * rows_per_block;
  const int y_hi = (y_lo + rows_per_block < H) ? y_lo + rows_per_block : H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float red[8][8];
  float* Gn = G + (size_t)n * KS * KS * C + k * PW;
  const uint4* gp = g + ((size_t)n * (C / PW) + k) * HW;
  const int Wc = W < 256 ? W : 256;
  const int nsub = 256 / Wc;                        // rows in flight per pass (1 when W >= 256)
  const int x0 = threadIdx.x % Wc, ysub = threadIdx.x / Wc;
  const bool active = ysub < nsub;
  const int cx0 = border_class(x0, W, P);
  float run[8], brun[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { run[e] = 0.f; brun[e] = 0.f; }
  auto add = [&](const uint4& v, float* acc) {
    float f[8];
    plane_unpack(v, mode, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  };
  int y = y_lo;
  while (y < y_hi) {
    const int cy = border_class(y, H, P);            // block-uniform run [y, y_end) of equal row class
    int y_end = y + 1;
    while (y_end < y_hi && border_class(y_end, H, P) == cy) ++y_end;
    if (active) {
      for (int yy = y + ysub; yy < y_end; yy += 4 * nsub) {
        uint4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int yq = yy + q * nsub;
          v[q] = (yq < y_end) ? __ldg(gp + (size_t)yq * W + x0) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) add(v[q], cx0 == P ? run : brun);
        for (int x = x0 + 256; x < W; x += 256) {     // W > 256 only
          for (int q = 0; q < 4; ++q) {
            const int yq = yy + q * nsub;
            if (yq >= y_end) break;
            const uint4 w = __ldg(gp + (size_t)yq * W + x);
            const int cx = border_class(x, W, P);
            if (cx == P) add(w, run);
            else {
              float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              add(w, t);
              for (int e = 0; e < PW; ++e) atomicAdd(Gn + (size_t)(cy * KS + cx) * C + e, t[e]);
            }
          }
        }
      }
    }
    // flush the run
    if (active && cx0 != P) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (e < PW) atomicAdd(Gn + (size_t)(cy * KS + cx0) * C + e, brun[e]);
        brun[e] = 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) run[e] = warp_sum(run[e]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][e] = run[e];
    __syncthreads();
    if (threadIdx.x < PW) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
      atomicAdd(Gn + (size_t)(cy * KS + P) * C + threadIdx.x, t);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) run[e] = 0.f;
    y = y_end;
  }
}

int tc_launch_class_sum(Plan* p, const void* g, cudaStream_t st_) {
  const int rpb = 32;
  dim3 grid((p->s.H + rpb - 1) / rpb, p->C / tc_state(p)->pw, p->BK);
  tc_class_sum_kernel<<<grid, 256, 0, st_>>>(reinterpret_cast<const uint4*>(g), p->G, p->s.H, p->s.W, p->C,
                                             p->s.dec_k, rpb, plane_mode(p));
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// chunk-planar [n][C/PW][HW][PW] -> NHWC fp32 [n][HW][C]   (debug reads only)
__global__ void tc_export_kernel(const void* __restrict__ src, float* __restrict__ dst, size_t total,
                                 int HW, int C, int mode) {
  const int PW = mode == 2 ? 4 : 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t pn = i / C;
    const int pix = (int)(pn % HW);
    const size_t n = pn / HW;
    const size_t j = ((n * (C / PW) + c / PW) * HW + pix) * PW + c % PW;
    dst[i] = mode == 2 ? reinterpret_cast<const float*>(src)[j] : from_h(reinterpret_cast<const uint16_t*>(src)[j], mode);
  }
}

int tc_export_f32(Plan* p, const void* src_bf16, float* dst, size_t n, cudaStream_t st_) {
  tc_export_kernel<<<p->num_sms * 4, 256, 0, st_>>>(src_bf16, dst, n, p->HW, p->C, plane_mode(p));
  IOD_LAUNCH_CHECK(p);
  return 0;
}

// seed plane (16-bit: [n][HW][8], 4 real channels; tf32: [n][HW][4] fp32) -> fp32 [n][HW][4]   (debug reads only)
__global__ void tc_export_seed_kernel(const void* __restrict__ src, float* __restrict__ dst, size_t npix, int mode) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix * 4; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = mode == 2 ? reinterpret_cast<const float*>(src)[i]
                       : from_h(reinterpret_cast<const uint16_t*>(src)[(i / 4) * 8 + i % 4], mode);
}

int tc_export_seed(Plan* p, const void* seed8, float* dst, cudaStream_t st_) {
  tc_export_seed_kernel<<<p->num_sms * 4, 256, 0, st_>>>(seed8, dst, (size_t)p->BK * p->HW, plane_mode(p));
  IOD_LAUNCH_CHECK(p);
  return 0;
}

}  // namespace iod
