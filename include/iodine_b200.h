/*
 * iodine_b200.h -- C ABI of the B200-native IODINE refinement-loop engine.
 *
 * The reference (zhixuan-lin/IODINE) has no FFI of its own: the seam is the Python class
 * surface of lib/modeling/iodine.py.  Each entry point below names
 * the reference function it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; iodine_last_error()
 *     returns a thread-local, NUL-terminated description of the last failure.
 *   - unless a name ends in _host, every data pointer is a caller-owned DEVICE pointer
 *     (fp32, contiguous, the reference's own NCHW / [B,K,L] layouts).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued on it and nothing synchronises unless stated.
 *   - a plan is bound to the device that is current at iodine_plan_create() time, is not
 *     thread-safe and is not re-entrant (mirrors the per-instance state of the reference
 *     module, iodine.py:37-52).
 *   - no hidden allocations after iodine_plan_create(): the activation workspace is
 *     queried with iodine_plan_workspace_bytes() and supplied by the caller.
 */
#ifndef IODINE_B200_H_
#define IODINE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define IODINE_API __attribute__((visibility("default")))
#else
#define IODINE_API
#endif

#define IODINE_ABI_VERSION 2
#define IODINE_MAX_LAYERS 8

/* arithmetic of the decoder convolutions (everything else is always fp32) */
enum IodinePrecision {
  IODINE_FP32 = 0,  /* FFMA, fp32 activations: the exact path                           */
  IODINE_BF16 = 1,  /* tcgen05 kind::f16 (bf16 operands, fp32 accumulate in TMEM)      */
  IODINE_TF32 = 2,  /* tcgen05 kind::tf32: fp32 activations and weights rounded to tf32 (10-bit mantissa, 8-bit
                       exponent), fp32 accumulate in TMEM -- what cuDNN runs the reference's nn.Conv2d with on a GPU */
  IODINE_FP16 = 3   /* tcgen05 kind::f16 with fp16 operands (10-bit mantissa), fp32 accumulate */
};

/* Mirrors the cfg.ARCH fields the reference model reads (iodine.py:10-32,
 * lib/config/defaults.py:35-100) plus the batch geometry. */
typedef struct IodineShape {
  int32_t B;           /* images per call handled by this plan (per rank)               */
  int32_t K;           /* ARCH.SLOTS                                                     */
  int32_t L;           /* ARCH.DIM_LATENT                                                */
  int32_t H, W;        /* ARCH.IMG_SIZE (square in the reference; kept separate)        */
  int32_t T;           /* ARCH.ITERS                                                     */
  int32_t img_c;       /* ARCH.IMG_CHANNELS, must be 3                                   */
  int32_t dec_layers;  /* ARCH.DEC.CONV_LAYERS                                           */
  int32_t dec_chan;    /* ARCH.DEC.CONV_CHAN                                             */
  int32_t dec_k;       /* ARCH.DEC.KERNEL_SIZE                                           */
  int32_t ref_layers;  /* ARCH.REF.CONV_LAYERS                                           */
  int32_t ref_chan;    /* ARCH.REF.CONV_CHAN                                             */
  int32_t ref_k;       /* ARCH.REF.KERNEL_SIZE                                           */
  int32_t ref_stride;  /* ARCH.REF.STRIDE                                                */
  int32_t mlp_units;   /* ARCH.REF.MLP_UNITS                                             */
  int32_t layernorm;   /* ARCH.LAYERNORM                                                 */
  float   sigma;       /* ARCH.SIGMA                                                     */
  int32_t precision;   /* enum IodinePrecision                                           */
  /* K-split (SURVEY.md 8e fallback, for B < number of GPUs -- e.g. one image with K = 16): slot_ranks > 1 makes
   * this plan own the slots [slot_rank * K/slot_ranks, (slot_rank+1) * K/slot_ranks) of every image; K stays
   * ARCH.SLOTS and must divide.  Every per-slot tensor of the entry points (eps, z, posterior, LSTM state) then
   * holds K/slot_ranks slots; pred / mask / mean of decode()/reconstruct() are complete ([B,K,...]) on every rank.
   * Needs iodine_plan_set_comm with exactly slot_ranks ranks before the first step.  0 or 1 = whole images. */
  int32_t slot_ranks;
  int32_t slot_rank;
} IodineShape;

/* One pointer per state_dict key of the reference model, PyTorch layouts
 * (conv OIHW, linear [out,in]), fp32, DEVICE memory.  iodine_plan_set_weights() repacks
 * them once into kernel layouts held inside the plan. */
typedef struct IodineWeights {
  const float* dec_w[IODINE_MAX_LAYERS];  /* decoder.mlc.layers.{i}.weight               */
  const float* dec_b[IODINE_MAX_LAYERS];  /* decoder.mlc.layers.{i}.bias                 */
  const float* dec_out_w;                 /* decoder.conv.weight  [4,C,k,k]              */
  const float* dec_out_b;                 /* decoder.conv.bias    [4]                    */
  const float* ref_w[IODINE_MAX_LAYERS];  /* refine.mlc.layers.{i}.weight                */
  const float* ref_b[IODINE_MAX_LAYERS];  /* refine.mlc.layers.{i}.bias                  */
  const float* mlp_w;                     /* refine.mlp.layers.0.weight [M,Cr]           */
  const float* mlp_b;                     /* refine.mlp.layers.0.bias                    */
  const float* lstm_w_ih;                 /* refine.lstm.weight_ih [4M, M+4L]            */
  const float* lstm_w_hh;                 /* refine.lstm.weight_hh [4M, M]               */
  const float* lstm_b_ih;                 /* refine.lstm.bias_ih                         */
  const float* lstm_b_hh;                 /* refine.lstm.bias_hh                         */
  const float* mean_w;                    /* refine.mean_update.weight [L,M]             */
  const float* mean_b;                    /* refine.mean_update.bias                     */
  const float* logvar_w;                  /* refine.logvar_update.weight                 */
  const float* logvar_b;                  /* refine.logvar_update.bias                   */
  const float* init_mean;                 /* posterior.init_mean [L]                     */
  const float* init_logvar;               /* posterior.init_logvar [L]                   */
} IodineWeights;

/* One pointer per state_dict key, as IodineWeights: where iodine_train_step writes the gradient of the loss with
 * respect to that parameter (same shape and layout as the parameter, fp32, DEVICE memory; NULL = not wanted). */
typedef struct IodineGrads {
  float* dec_w[IODINE_MAX_LAYERS];
  float* dec_b[IODINE_MAX_LAYERS];
  float* dec_out_w;
  float* dec_out_b;
  float* ref_w[IODINE_MAX_LAYERS];        /* layer 0: the full 17-channel layout [Cr,17,k,k] */
  float* ref_b[IODINE_MAX_LAYERS];
  float* mlp_w;
  float* mlp_b;
  float* lstm_w_ih;
  float* lstm_w_hh;
  float* lstm_b_ih;
  float* lstm_b_hh;
  float* mean_w;
  float* mean_b;
  float* logvar_w;
  float* logvar_b;
  float* init_mean;
  float* init_logvar;
} IodineGrads;

typedef struct IodinePlan IodinePlan;

IODINE_API int         iodine_abi_version(void);
IODINE_API const char* iodine_last_error(void);

/* IODINE.__init__ (iodine.py:8-33): validates the architecture, sizes every buffer. */
IODINE_API int iodine_plan_create(const IodineShape* shape, IodinePlan** plan_out);
IODINE_API int iodine_plan_destroy(IodinePlan* plan);

/* bytes of activation workspace the caller must supply (device memory, 1024-B aligned) */
IODINE_API int iodine_plan_workspace_bytes(const IodinePlan* plan, size_t* bytes_out);
IODINE_API int iodine_plan_set_workspace(IodinePlan* plan, void* workspace, size_t bytes);

/* nn.Module.load_state_dict for the keys listed in IodineWeights (checkpoint.py:68). */
IODINE_API int iodine_plan_set_weights(IodinePlan* plan, const IodineWeights* w, void* stream);

/* Gaussian.init_unit (iodine.py:607-618) + `self.lstm_hidden = None` (iodine.py:82):
 * post_mean/post_logvar[B,K,L] <- init_mean/init_logvar tiled; lstm_h/lstm_c[B*K,M] <- 0 */
IODINE_API int iodine_init_state(IodinePlan* plan, float* post_mean, float* post_logvar,
                      float* lstm_h, float* lstm_c, void* stream);

/* One iteration of the loop body of IODINE.encode (iodine.py:83-100): elbo() 161-241,
 * the five gradients of (B*elbo).backward() 90 in closed form, get_input_encoding()
 * 243-343, refine() 466-503 and Gaussian.update() 636-651.
 *   x[B,3,H,W]; eps_t[B,K,L]; post_mean/post_logvar[B,K,L] in-out;
 *   lstm_h/lstm_c[B*K,M] in-out (true h / true c of nn.LSTMCell);
 *   elbo_terms_out[2] = { sum_b log-likelihood, sum_b KL } of THIS step (nullable);
 *   aux_out: nullable; if given receives the reference's normalised refinement input
 *            [B,K,17,H,W] followed by latent [B,K,4L] (tests only). */
IODINE_API int iodine_refine_step(IodinePlan* plan, const float* x, const float* eps_t,
                       float* post_mean, float* post_logvar, float* lstm_h, float* lstm_c,
                       float* elbo_terms_out, float* aux_out, void* stream);

/* IODINE.elbo (iodine.py:161-241) for given posterior and noise, no gradients:
 * elbo_terms_out[2] = { sum_b log-likelihood, sum_b KL }. */
IODINE_API int iodine_elbo(IodinePlan* plan, const float* x, const float* eps_t,
                const float* post_mean, const float* post_logvar,
                float* elbo_terms_out, void* stream);

/* IODINE.encode (iodine.py:73-105): eps[T+1,B,K,L]; z_out[B,K,L];
 * elbo_terms_out[T,2] nullable; post_out[2,B,K,L] nullable (final mean, logvar). */
IODINE_API int iodine_encode(IodinePlan* plan, const float* x, const float* eps, float* z_out,
                  float* elbo_terms_out, float* post_out, void* stream);

/* IODINE.decode (iodine.py:59-71): z[B,K,L] -> pred[B,3,H,W], mask[B,K,1,H,W],
 * mean[B,K,3,H,W] (any output may be NULL). */
IODINE_API int iodine_decode(IodinePlan* plan, const float* z, float* pred_out, float* mask_out,
                  float* mean_out, void* stream);

/* IODINE.reconstruct (iodine.py:107-112) = encode + decode. */
IODINE_API int iodine_reconstruct(IodinePlan* plan, const float* x, const float* eps,
                       float* pred_out, float* mask_out, float* mean_out, float* z_out,
                       float* elbo_terms_out, void* stream);

/* Same, HOST buffers (what lib/eval/ari_eval.py:22 + lib/engine/eval.py:27 do around the
 * call: image.to(device) ... .cpu()).  Copies x and eps host->device, runs reconstruct,
 * copies the outputs device->host and synchronises the stream.  Host pointers should be
 * pinned for full PCIe speed.  Staging buffers live inside the caller's workspace. */
IODINE_API int iodine_reconstruct_host(IodinePlan* plan, const float* x_host, const float* eps_host,
                            float* pred_host, float* mask_host, float* mean_host,
                            float* z_host, float* elbo_terms_host, void* stream);

/* The same sequence WITHOUT the final synchronisation: everything is enqueued on `stream` and the host buffers
 * are valid once the caller has synchronised it.  Two plans on two streams double-buffer an evaluation loop:
 * the copies of one batch overlap the kernels of the other (bench.py's e2e leg). */
IODINE_API int iodine_reconstruct_host_async(IodinePlan* plan, const float* x_host, const float* eps_host,
                                  float* pred_host, float* mask_host, float* mean_host,
                                  float* z_host, float* elbo_terms_host, void* stream);

/* The evaluation flow with HOST buffers: lib/engine/eval.py:21-30 (`model.reconstruct(image)` per batch) followed
 * by what lib/eval/ari_eval.py:32-39 keeps of the result -- the per-pixel ARGMAX over the K predicted masks.  Same
 * sequence as iodine_reconstruct_host[_async] (x and eps host->device, encode + decode, results device->host), but
 * the masks come back as argmax_host[B,H,W] uint8 (first maximum, as torch.argmax) instead of the fp32
 * [B,K,1,H,W] + [B,K,3,H,W] tensors: 1 byte per pixel instead of 4*4*K.  pred_host[B,3,H,W], z_host[B,K,L],
 * elbo_terms_host[T,2] as in iodine_reconstruct_host; any output may be NULL. */
IODINE_API int iodine_evaluate_host(IodinePlan* plan, const float* x_host, const float* eps_host, float* pred_host,
                         uint8_t* argmax_host, float* z_host, float* elbo_terms_host, void* stream);
IODINE_API int iodine_evaluate_host_async(IodinePlan* plan, const float* x_host, const float* eps_host,
                               float* pred_host, uint8_t* argmax_host, float* z_host, float* elbo_terms_host,
                               void* stream);

/* The logger side channel of IODINE.elbo (iodine.py:225-239): what the reference hands to `logger.update` on every
 * elbo() evaluation -- pred[0], mask[0,i,0] and mean[0,i] of image 0.  Returns those of the LAST elbo() evaluated by
 * this plan (the last refinement step of encode()/reconstruct(), or iodine_elbo): pred0[3,H,W], mask0[K,H,W],
 * mean0[K,3,H,W], device memory, any may be NULL. */
IODINE_API int iodine_plan_last_elbo_image0(IodinePlan* plan, float* pred0, float* mask0, float* mean0, void* stream);

/* TRAINING: IODINE.forward (iodine.py:115-158) + what lib/engine/train.py:60-65 does with its
 * result -- loss.mean(), optimizer.zero_grad(), loss.backward() -- in one call: the T+1 ELBO evaluations with the
 * refinement deltas kept attached, loss = -sum_i (i+1)/(T+1) elbo_i, and the gradient of the loss with respect to
 * every parameter, by hand-written kernels (csrc/train.cu: decoder weight gradients accumulated inside each step
 * next to the data-gradient chain, one backward sweep over a tape of the T refiner calls incl. the LSTM state
 * chain).  The tape lives in a second caller-supplied workspace, needed for training only.
 *   x[B,3,H,W]; eps[T+1,B,K,L] (the T+1 torch.randn_like draws, iodine.py:632);
 *   global_batch: images behind the batch means (0 = this plan's B).  With a communicator installed the ranks hold
 *     different images of one global batch: pass the global size; gradients, loss and ELBO table are then summed
 *     over the ranks with ONE ncclAllReduce of the flat gradient buffer (replaces DataParallel's gradient reduction);
 *   grads_out: overwritten (the reference zeroes the gradients before backward); loss_out[1];
 *   elbo_terms_out[T+1,2] = { sum_b log-likelihood, sum_b KL } of every ELBO evaluation (nullable). */
IODINE_API int iodine_plan_train_workspace_bytes(IodinePlan* plan, size_t* bytes_out);
IODINE_API int iodine_plan_set_train_workspace(IodinePlan* plan, void* workspace, size_t bytes);
IODINE_API int iodine_train_step(IodinePlan* plan, const float* x, const float* eps, int32_t global_batch,
                      const IodineGrads* grads_out, float* loss_out, float* elbo_terms_out, void* stream);

/* Multi-GPU (replaces torch.nn.DataParallel, lib/modeling/build.py:11-12, and the replica-mean of
 * lib/engine/train.py:61).  The path shards by whole images, one process and one plan per GPU; every K-way reduction
 * is inside one image, so the ONLY cross-rank quantity is the [T,2] table of batch sums behind the ELBO means
 * (iodine.py:193, 220).  After this call iodine_encode / iodine_reconstruct / iodine_reconstruct_host[_async]
 * sum that table over the communicator with one ncclAllReduce on the caller's stream (stream-ordered, no host
 * synchronisation) before they return it; iodine_elbo does the same with its two sums; nothing else crosses ranks.
 * All ranks must make the same calls, with elbo_terms_out given on all of them or on none.  iodine_refine_step stays
 * rank-local.
 *   nccl_comm: an ncclComm_t of the calling process (NULL uninstalls); NCCL is resolved from the process at run
 *   time (dlopen of libnccl.so.2), the library does not link against it.
 * K-split plans (IodineShape.slot_ranks > 1; replaces DataParallel, lib/modeling/build.py:11-12, for a batch smaller
 * than the number of GPUs): the ranks hold different SLOTS of the same images, and the K-way reductions of
 * IODINE.elbo (mask softmax iodine.py:185, mixture logsumexp 213-216, mask posterior 292, leave-one-out 324) cross
 * ranks.  The exchange is ONE ncclAllGather per elbo() evaluation of the decoder's 4-channel output
 * [B,K,H,W,4] fp32 (256 KB per slot at 128x128) on the stream of the call; every rank then evaluates the pixel
 * mixture for all K slots and keeps the seeds / auxiliary channels of its own.  The ELBO table is summed as above
 * (the image log-likelihood is contributed by slot_rank 0 only, the KL by every rank for its slots).  rank / nranks
 * must equal slot_rank / slot_ranks. */
IODINE_API int iodine_plan_set_comm(IodinePlan* plan, void* nccl_comm, int32_t rank, int32_t nranks);

/* Evaluator tail: lib/eval/ari_eval.py:32-39 (argmax over the K predicted masks) +
 * lib/utils/ari.py:36-54 (contingency table) + lib/utils/ari.py:6-33 (ARI), per image, on the device.
 *   mask[B,K,H,W] fp32 (IODINE.reconstruct's mask); gt_masks[B,G,H,W] uint8 (the reference's mask.byte(), padded
 *   to G masks per image); n_gt[B] valid ground-truth masks per image;
 *   table_out[B,G,K] uint64 contingency tables; ari_out[B] f64 (nullable).  K, G <= 16. */
IODINE_API int iodine_ari(const float* mask, const uint8_t* gt_masks, const int32_t* n_gt, int32_t B, int32_t K,
                          int32_t G, int32_t H, int32_t W, uint64_t* table_out, double* ari_out, void* stream);

/* Test hook: copy a named internal buffer of the last step to dst (device memory).
 * names: "out4" [BK,H,W,4], "seed4" [BK,H,W,4], "dz" [BK,L], "act<i>" [BK,H,W,C] (fp32
 * view of decoder layer i's post-activation), "pool" [BK,Cr], "stats" [BK,4,2] (f64),
 * "z" [BK,L].  *bytes_out receives the byte count; dst may be NULL to query the size. */
IODINE_API int iodine_debug_read(IodinePlan* plan, const char* name, void* dst, size_t dst_bytes,
                      size_t* bytes_out, void* stream);

/* Measurement hook (bench.py's roofline): while enabled, every decoder C->C convolution
 * launch (forward and data-gradient -- the dominant kernel) is bracketed by CUDA events on
 * the launching stream.  iodine_plan_profile_read() synchronises those events, returns
 * the summed device time in milliseconds and the number of bracketed launches, and resets.
 * enable = 2 brackets the pixel-mixture (aux-input fuse) kernel of every refinement step instead. */
IODINE_API int iodine_plan_profile(IodinePlan* plan, int enable);
IODINE_API int iodine_plan_profile_read(IodinePlan* plan, double* ms_total_out, uint64_t* launches_out);

/* number of kernel launches issued by this plan since creation (bench's gpu_launches) */
IODINE_API int iodine_plan_launch_count(const IodinePlan* plan, uint64_t* count_out);

#ifdef __cplusplus
}
#endif
#endif /* IODINE_B200_H_ */
