// ari.cu -- device-side Adjusted Rand Index of predicted slot masks against ground-truth masks.
//
// Replaces the evaluator tail of the reference: lib/eval/ari_eval.py:32-39 (argmax over
// the K predicted masks, one-hot) and lib/utils/ari.py:36-54 (contingency table via byte AND + pixel sums)
// followed by lib/utils/ari.py:6-33 (ARI from the table).  Integer work, bit-exact; the ARI itself is f64.
// The reference moves [K,H,W] + [N,H,W] masks to the host per image (ari_eval.py:38); here only B doubles leave
// the device.
#include "common.cuh"

namespace iod {

// table[b][g][k] = #{pixels p : (gt[b][g][p] & 1) != 0 and argmax_k' mask[b][k'][p] == k}
//   mask: [B][K][HW] fp32 (IODINE.reconstruct's mask[B,K,1,H,W]);  gt: [B][G][HW] uint8 (the reference's
//   mask.byte(); bitwise AND with the 0/1 one-hot keeps bit 0);  n_gt[b] <= G valid ground-truth masks of image b
//   (the reference's list of per-image (N,H,W) tensors, padded to G).
// torch.argmax semantics: first maximal index, NaN counts as maximal.
template <int KMAX, int GMAX>
__global__ void __launch_bounds__(256)
ari_table_kernel(const float* __restrict__ mask, const uint8_t* __restrict__ gt, const int32_t* __restrict__ n_gt,
                 unsigned long long* __restrict__ table, int K, int G, int HW) {
  const int b = blockIdx.y;
  __shared__ unsigned int s_tab[GMAX * KMAX];
  for (int i = threadIdx.x; i < GMAX * KMAX; i += blockDim.x) s_tab[i] = 0u;
  __syncthreads();
  const int ng = n_gt[b] < G ? n_gt[b] : G;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    int best = 0;
    float bv = mask[((size_t)b * K) * HW + pix];
    bool best_nan = bv != bv;
    for (int k = 1; k < K; ++k) {
      const float v = mask[((size_t)b * K + k) * HW + pix];
      if (!best_nan && (v != v || v > bv)) { best = k; bv = v; best_nan = v != v; }
    }
    for (int g = 0; g < ng; ++g)
      if (gt[((size_t)b * G + g) * HW + pix] & 1u) atomicAdd(&s_tab[g * KMAX + best], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < GMAX * KMAX; i += blockDim.x) {
    const int g = i / KMAX, k = i % KMAX;
    if (g < G && k < K && s_tab[i]) atomicAdd(&table[((size_t)b * G + g) * K + k], (unsigned long long)s_tab[i]);
  }
}

// compute_ari (lib/utils/ari.py:6-33) per image, in f64; comb(x, 2) = x (x - 1) / 2
__global__ void ari_from_table_kernel(const unsigned long long* __restrict__ table, const int32_t* __restrict__ n_gt,
                                      double* __restrict__ ari, int B, int G, int K) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int ng = n_gt[b] < G ? n_gt[b] : G;
  const unsigned long long* t = table + (size_t)b * G * K;
  double comb_a = 0.0, comb_b = 0.0, comb_t = 0.0, n = 0.0;
  for (int g = 0; g < ng; ++g) {
    double a = 0.0;
    for (int k = 0; k < K; ++k) {
      const double v = (double)t[g * K + k];
      a += v;
      comb_t += v * (v - 1.0) * 0.5;
    }
    comb_a += a * (a - 1.0) * 0.5;
    n += a;
  }
  for (int k = 0; k < K; ++k) {
    double c = 0.0;
    for (int g = 0; g < ng; ++g) c += (double)t[g * K + k];
    comb_b += c * (c - 1.0) * 0.5;
  }
  const double comb_n = n * (n - 1.0) * 0.5;
  if (comb_b == comb_a && comb_a == comb_n && comb_n == comb_t) {
    ari[b] = 1.0;                                     // "the perfect case" (ari.py:23-25)
  } else {
    ari[b] = (comb_t - comb_a * comb_b / comb_n) / (0.5 * (comb_a + comb_b) - (comb_a * comb_b) / comb_n);
  }
}

}  // namespace iod

using namespace iod;

extern "C" {

IODINE_API int iodine_ari(const float* mask, const uint8_t* gt_masks, const int32_t* n_gt, int32_t B, int32_t K,
                          int32_t G, int32_t H, int32_t W, uint64_t* table_out, double* ari_out, void* stream) {
  IOD_REQUIRE(mask && gt_masks && n_gt && table_out, "iodine_ari: null argument");
  IOD_REQUIRE(B > 0 && K > 0 && K <= 16 && G > 0 && G <= 16 && H > 0 && W > 0, "iodine_ari: unsupported B=%d K=%d G=%d", B, K, G);
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  IOD_CHECK_CUDA(cudaMemsetAsync(table_out, 0, (size_t)B * G * K * sizeof(uint64_t), st));
  int gx = (HW + 255) / 256;
  if (gx > 64) gx = 64;
  dim3 grid(gx, B);
  ari_table_kernel<16, 16><<<grid, 256, 0, st>>>(mask, gt_masks, n_gt, reinterpret_cast<unsigned long long*>(table_out), K, G, HW);
  IOD_CHECK_CUDA(cudaGetLastError());
  if (ari_out) {
    ari_from_table_kernel<<<(B + 127) / 128, 128, 0, st>>>(reinterpret_cast<const unsigned long long*>(table_out), n_gt,
                                                           ari_out, B, G, K);
    IOD_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
