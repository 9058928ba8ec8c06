// plan.cu -- the C ABI (include/iodine_b200.h): plan lifetime, workspace carve-up and the
// per-step kernel sequence that replaces IODINE.encode/decode/reconstruct/elbo
// (reference lib/modeling/iodine.py:59-241).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>

#include "common.cuh"

namespace iod {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_kl(Plan* p, const float* mu, const float* logvar, cudaStream_t st);

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Walk the workspace layout; with base == nullptr only the size is computed.
static size_t carve(Plan* p, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> void* {
    void* r = base ? base + off : nullptr;
    off = align_up(off + bytes, 1024);
    return r;
  };
  const IodineShape& s = p->s;
  const size_t BK = p->BK, HW = p->HW, C = p->C, L = s.L, M = p->M, Cr = p->Cr;
  const size_t eb = act_elem_bytes(p);
  for (int l = 0; l < s.dec_layers; ++l) p->act[l] = take(BK * HW * C * eb);
  for (int i = 0; i < 2; ++i) {
    const bool need = s.dec_layers > 1 || i == 0;      // (gbuf[0] also holds dJ/d(pre-activation 0) of a training step)
    p->gbuf[i] = take(need ? BK * HW * C * eb : 1024);
  }
  // K-split: one buffer for all ranks' slots, rank-major as ncclAllGather fills it; this rank's decoder writes its
  // own window in place (sendbuff == recvbuff + rank * count)
  p->out4_all = (float*)take((size_t)p->ks_ranks * BK * HW * 4 * sizeof(float));
  p->out4 = base ? p->out4_all + (size_t)p->ks_rank * BK * HW * 4 : nullptr;
  p->seed4 = (float*)take(BK * HW * 4 * sizeof(float));
  p->auxs = (float*)take(BK * HW * 12 * sizeof(float));
  const bool rtc = rtc_enabled(p);
  p->enc20 = (float*)take(rtc ? 1024 : BK * HW * 20 * sizeof(float));
  p->enc16 = take(rtc && !rtc_fused_aux(p) ? BK * HW * 32 : 1024);
  p->lik = (float*)take((size_t)s.B * HW * sizeof(float));
  size_t r0 = (size_t)p->ref_h[1] * p->ref_w[1] * Cr;
  size_t r1 = s.ref_layers > 1 ? (size_t)p->ref_h[2] * p->ref_w[2] * Cr : 256;
  p->rbuf[0] = (float*)take(rtc ? 1024 : BK * r0 * sizeof(float));
  p->rbuf[1] = (float*)take(rtc ? 1024 : BK * r1 * sizeof(float));
  p->r16[0] = take(rtc ? BK * r0 * 2 : 1024);
  p->r16[1] = take(rtc ? BK * r1 * 2 : 1024);
  p->z = (float*)take(BK * L * sizeof(float));
  p->u = (float*)take(BK * p->n_class * C * sizeof(float));
  p->eu = (float*)take(BK * p->n_class * C * sizeof(float));
  p->ubig = (int*)take(BK * sizeof(int));
  p->G = (float*)take(BK * p->n_class * C * sizeof(float));
  p->dz = (float*)take(BK * L * sizeof(float));
  p->stats = (double*)take(BK * 8 * sizeof(double));
  p->lnp = (float*)take(BK * 8 * sizeof(float));
  p->accum = (double*)take(2 * sizeof(double));
  p->pool = (float*)take(BK * Cr * sizeof(float));
  p->xin = (float*)take(BK * (M + 4 * L) * sizeof(float));
  // split-K partial sums of the LSTM gate GEMM [BK,4M] and, later in the step, of the two heads [BK,2L]
  p->gates = (float*)take((size_t)LSTM_KSPLIT * BK * (4 * M > 2 * L ? 4 * M : 2 * L) * sizeof(float));
  p->st_mean = (float*)take(BK * L * sizeof(float));
  p->st_logvar = (float*)take(BK * L * sizeof(float));
  p->st_h = (float*)take(BK * M * sizeof(float));
  p->st_c = (float*)take(BK * M * sizeof(float));
  p->st_z = (float*)take(BK * L * sizeof(float));
  p->st_terms = (float*)take((size_t)(s.T + 1) * 2 * sizeof(float));
  p->hx = (float*)take((size_t)s.B * 3 * HW * sizeof(float));
  p->heps = (float*)take((size_t)(s.T + 1) * BK * L * sizeof(float));
  p->hpred = (float*)take((size_t)s.B * 3 * HW * sizeof(float));
  const size_t BKt = (size_t)s.B * p->K_total;               // decode() returns all K slots of every image
  p->hmask = (float*)take(BKt * HW * sizeof(float));
  p->hmean = (float*)take(BKt * 3 * HW * sizeof(float));
  p->hamax = (uint8_t*)take((size_t)s.B * HW);
  // logger side channel (iodine.py:226-239): image 0 of the last elbo() evaluation
  p->log_pred = (float*)take((size_t)3 * HW * sizeof(float));
  p->log_mask = (float*)take((size_t)p->K_total * HW * sizeof(float));
  p->log_mean = (float*)take((size_t)p->K_total * 3 * HW * sizeof(float));
  return off;
}

static int alloc_f(float** dst, size_t n) {
  IOD_CHECK_CUDA(cudaMalloc((void**)dst, (n ? n : 1) * sizeof(float)));
  return 0;
}

__global__ void terms_kernel(const double* __restrict__ accum, float* __restrict__ out) {
  if (threadIdx.x < 2) out[threadIdx.x] = (float)accum[threadIdx.x];
}
__global__ void sample_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                              const float* __restrict__ eps, float* __restrict__ z, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    z[i] = mu[i] + expf(0.5f * lv[i]) * eps[i];
}

// ---------------------------------------------------------------- profiling brackets
// cls: 1 = decoder C->C convolutions, 2 = the pixel-mixture (aux-input fuse) kernel
static void prof_mark(Plan* p, cudaStream_t st, int cls = 1) {
  if (p->profiling != cls) return;
  if (p->prof_used == p->prof_events.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    p->prof_events.push_back(e);
  }
  cudaEventRecord(p->prof_events[p->prof_used++], st);
}

void plan_prof_mark(Plan* p, cudaStream_t st, int cls) { prof_mark(p, st, cls); }

static int gather_out4(Plan* p, cudaStream_t st);

// ---------------------------------------------------------------- decoder forward / dgrad
static int decoder_forward(Plan* p, const float* mu, const float* lv, const float* eps,
                           const float* z_in, cudaStream_t st) {
  const IodineShape& s = p->s;
  if (launch_sample_l1(p, mu, lv, eps, z_in, (float*)p->act[0], st)) return 1;
  for (int l = 1; l < s.dec_layers; ++l) {
    prof_mark(p, st);
    if (tc_mode(p)) {
      if (tc_launch_conv(p, l, false, p->act[l - 1], nullptr, p->act[l], nullptr, st)) return 1;
    } else {
      if (launch_conv_cc(p, (const float*)p->act[l - 1], p->dec[l].w, p->dec[l].b, nullptr,
                         (float*)p->act[l], nullptr, 0, st))
        return 1;
    }
    prof_mark(p, st);
  }
  if (tc_mode(p)) {
    if (tc_launch_out4(p, p->act[s.dec_layers - 1], p->out4, st)) return 1;
  } else {
    if (launch_conv_out4(p, (const float*)p->act[s.dec_layers - 1], p->out4, st)) return 1;
  }
  return gather_out4(p, st);
}

static int decoder_dgrad(Plan* p, cudaStream_t st) {
  const IodineShape& s = p->s;
  const int n = s.dec_layers;
  IOD_CHECK_CUDA(cudaMemsetAsync(p->G, 0, (size_t)p->BK * p->n_class * p->C * sizeof(float), st));
  if (tc_mode(p)) {
    if (tc_launch_dgrad_in4(p, p->seed4, p->act[n - 1], p->gbuf[0], st)) return 1;
  } else {
    if (launch_dgrad_in4(p, p->seed4, (const float*)p->act[n - 1], (float*)p->gbuf[0], st)) return 1;
  }
  int cur = 0;
  bool summed = false;                 // class sums already produced by the last data-gradient's epilogue
  for (int l = n - 1; l >= 1; --l) {
    const bool last = (l == 1);
    prof_mark(p, st);
    if (tc_mode(p)) {
      const bool fuse = last && tc_dsum_fused(p);
      if (tc_launch_conv(p, l, true, p->gbuf[cur], p->act[l - 1], p->gbuf[cur ^ 1], fuse ? p->G : nullptr, st)) return 1;
      summed = fuse;
    } else {
      if (launch_conv_cc(p, (const float*)p->gbuf[cur], p->dec[l].wt, nullptr,
                         (const float*)p->act[l - 1], last ? nullptr : (float*)p->gbuf[cur ^ 1],
                         last ? p->G : nullptr, last ? 2 : 1, st))
        return 1;
    }
    prof_mark(p, st);
    cur ^= 1;
  }
  // the tensor-core path stores dJ/d(pre-activation 1) and reduces it over the border classes
  // of the collapsed first layer in a separate bandwidth-bound pass
  if (tc_mode(p) && !summed) return tc_launch_class_sum(p, p->gbuf[cur], st);
  return 0;
}

// log_image0: keep image 0 of this elbo() evaluation for the logger side channel (every write replaces the previous
// one, iodine.py:226-239, so inside a loop only the LAST step's needs to be produced)
static int refine_step(Plan* p, const float* x, const float* eps_t, float* mu, float* lv, float* h,
                       float* c, float* terms_out, float* aux_out, cudaStream_t st, bool log_image0 = true) {
  if (decoder_forward(p, mu, lv, eps_t, nullptr, st)) return 1;
  if (launch_mixture(p, x, true, st, rtc_fused_aux(p))) return 1;
  if (log_image0 && launch_recombine(p, p->log_pred, p->log_mask, p->log_mean, 1, st)) return 1;   // what elbo() hands the logger
  if (decoder_dgrad(p, st)) return 1;
  if (launch_post_grads(p, mu, lv, eps_t, nullptr, st)) return 1;
  if (rtc_enabled(p)) {
    if (!rtc_fused_aux(p) && launch_assemble16(p, x, st)) return 1;   // fused: layer 0 assembles its own operand
    if (rtc_launch_refine_convs(p, st)) return 1;
  } else {
    if (launch_assemble(p, x, st)) return 1;
    if (launch_refine_convs(p, p->enc20, st)) return 1;
  }
  if (aux_out && launch_export_aux(p, x, aux_out, st)) return 1;
  if (launch_head(p, mu, lv, h, c, st)) return 1;
  if (terms_out) {
    terms_kernel<<<1, 32, 0, st>>>(p->accum, terms_out);
    IOD_LAUNCH_CHECK(p);
  }
  return 0;
}

static int check_ready(Plan* p) {
  IOD_REQUIRE(p != nullptr, "null plan");
  IOD_REQUIRE(p->ws != nullptr, "workspace not set (iodine_plan_set_workspace)");
  IOD_REQUIRE(p->weights_set, "weights not set (iodine_plan_set_weights)");
  return 0;
}

static int do_encode(Plan* p, const float* x, const float* eps, float* z_out, float* terms,
                     float* post_out, cudaStream_t st) {
  const IodineShape& s = p->s;
  const size_t nl = (size_t)p->BK * s.L;
  if (launch_init_state(p, p->st_mean, p->st_logvar, p->st_h, p->st_c, st)) return 1;
  for (int t = 0; t < s.T; ++t) {
    if (refine_step(p, x, eps + t * nl, p->st_mean, p->st_logvar, p->st_h, p->st_c,
                    terms ? terms + 2 * t : nullptr, nullptr, st, /*log_image0=*/t == s.T - 1))
      return 1;
  }
  sample_kernel<<<64, 256, 0, st>>>(p->st_mean, p->st_logvar, eps + (size_t)s.T * nl, z_out, (int)nl);
  IOD_LAUNCH_CHECK(p);
  if (post_out) {
    IOD_CHECK_CUDA(cudaMemcpyAsync(post_out, p->st_mean, nl * sizeof(float), cudaMemcpyDeviceToDevice, st));
    IOD_CHECK_CUDA(cudaMemcpyAsync(post_out + nl, p->st_logvar, nl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// encode() through a CUDA graph.  The first call with a given pointer tuple runs eagerly, the second captures
// (on a plan-owned stream: the caller's may be the legacy default stream, which cannot be captured) and every
// later one replays.  A plan keeps Plan::N_GRAPHS instantiated graphs keyed by the pointer tuple (least recently
// used one evicted), so that a caller alternating between buffer sets -- double buffering -- still replays.
// Profiling brackets (events) and IODINE_NO_GRAPH=1 keep the eager path.
static int do_encode_g(Plan* p, const float* x, const float* eps, float* z_out, float* terms, float* post_out,
                       cudaStream_t st) {
  if (!p->graphs || p->profiling) return do_encode(p, x, eps, z_out, terms, post_out, st);
  const void* key[5] = {x, eps, z_out, terms, post_out};
  Plan::EncodeGraph* g = nullptr;
  for (int i = 0; i < Plan::N_GRAPHS; ++i)
    if (p->graph[i].exec && !memcmp(p->graph[i].key, key, sizeof(key))) g = &p->graph[i];
  if (!g) {
    // seen once before (eagerly)?  then capture now; otherwise remember the tuple and run eagerly
    int seen = -1;
    for (int i = 0; i < Plan::N_GRAPHS; ++i)
      if (!p->graph[i].exec && p->graph[i].warm && !memcmp(p->graph[i].key, key, sizeof(key))) seen = i;
    if (seen < 0) {
      int victim = 0;                               // a free slot, else the least recently used one
      for (int i = 0; i < Plan::N_GRAPHS; ++i) {
        if (!p->graph[i].exec && !p->graph[i].warm) { victim = i; break; }
        if (p->graph[i].used < p->graph[victim].used) victim = i;
      }
      Plan::EncodeGraph& v = p->graph[victim];
      if (v.exec) { cudaGraphExecDestroy(v.exec); v.exec = nullptr; }
      memcpy(v.key, key, sizeof(key));
      v.warm = true;
      v.used = ++p->graph_clock;
      return do_encode(p, x, eps, z_out, terms, post_out, st);
    }
    g = &p->graph[seen];
    if (!p->gstream) {
      IOD_CHECK_CUDA(cudaStreamCreateWithFlags(&p->gstream, cudaStreamNonBlocking));
      IOD_CHECK_CUDA(cudaEventCreateWithFlags(&p->gev_in, cudaEventDisableTiming));
      IOD_CHECK_CUDA(cudaEventCreateWithFlags(&p->gev_out, cudaEventDisableTiming));
    }
    const uint64_t n0 = p->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(p->gstream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      p->graphs = false;
      return do_encode(p, x, eps, z_out, terms, post_out, st);
    }
    const int rc = do_encode(p, x, eps, z_out, terms, post_out, p->gstream);
    const cudaError_t ce = cudaStreamEndCapture(p->gstream, &graph);
    const uint64_t nk = p->launches - n0;
    p->launches = n0;
    if (rc || ce != cudaSuccess || !graph || cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      g->exec = nullptr;
      p->graphs = false;                        // this driver cannot capture the sequence: stay eager
      return do_encode(p, x, eps, z_out, terms, post_out, st);
    }
    cudaGraphDestroy(graph);
    g->kernels = nk;
  }
  g->used = ++p->graph_clock;
  IOD_CHECK_CUDA(cudaEventRecord(p->gev_in, st));
  IOD_CHECK_CUDA(cudaStreamWaitEvent(p->gstream, p->gev_in, 0));
  IOD_CHECK_CUDA(cudaGraphLaunch(g->exec, p->gstream));
  IOD_CHECK_CUDA(cudaEventRecord(p->gev_out, p->gstream));
  IOD_CHECK_CUDA(cudaStreamWaitEvent(st, p->gev_out, 0));
  p->launches += g->kernels;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Cross-rank exchange: the path shards by whole images, so the only quantity that
// crosses ranks is the [T,2] table of (sum_b log-likelihood, sum_b KL) behind the ELBO means of
// iodine.py:193,220.  It is summed with ONE ncclAllReduce on the caller's stream after the loop.
// NCCL is resolved at run time from the process (the host application -- torch here -- already
// carries libnccl.so.2), so the library has no link-time dependency on it and single-rank users
// never load it.  Prototype and enum values: nccl.h (ncclFloat32 = 7, ncclSum = 0, ncclSuccess = 0).
// ------------------------------------------------------------------------------------------------
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
static nccl_allgather_fn g_nccl_allgather = nullptr;
static nccl_allreduce_fn g_nccl_allreduce = nullptr;
static nccl_errstr_fn g_nccl_errstr = nullptr;

static int resolve_nccl() {
  if (g_nccl_allreduce) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy the host process already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("iodine_plan_set_comm: cannot load libnccl.so.2 (%s)", dlerror());
    return 1;
  }
  g_nccl_allreduce = reinterpret_cast<nccl_allreduce_fn>(dlsym(h, "ncclAllReduce"));
  g_nccl_allgather = reinterpret_cast<nccl_allgather_fn>(dlsym(h, "ncclAllGather"));
  g_nccl_errstr = reinterpret_cast<nccl_errstr_fn>(dlsym(h, "ncclGetErrorString"));
  if (!g_nccl_allreduce || !g_nccl_allgather) {
    g_nccl_allreduce = nullptr;
    set_error("iodine_plan_set_comm: libnccl has no ncclAllReduce / ncclAllGather");
    return 1;
  }
  return 0;
}

// K-split: every rank's [B,K_local,H,W,4] decoder output -> out4_all on all ranks (in place), the one data-path
// collective of that mode (SURVEY.md 8e "K-split fallback")
static int gather_out4(Plan* p, cudaStream_t st) {
  if (p->ks_ranks <= 1) return 0;
  IOD_REQUIRE(p->comm != nullptr, "K-split plan (slot_ranks = %d) has no communicator: call iodine_plan_set_comm first",
              p->ks_ranks);
  const size_t count = (size_t)p->BK * p->HW * 4;
  const int rc = g_nccl_allgather(p->out4, p->out4_all, count, /*ncclFloat32*/ 7, p->comm, st);
  if (rc != 0) {
    set_error("ncclAllGather of the decoder output failed on rank %d/%d: %s", p->comm_rank, p->comm_nranks,
              g_nccl_errstr ? g_nccl_errstr(rc) : "?");
    return 1;
  }
  return 0;
}

// sum the per-step table over the communicator (no-op for a single rank or a null table)
static int reduce_terms(Plan* p, float* terms, size_t count, cudaStream_t st) {
  if (!p->comm || !terms || !count) return 0;
  const int rc = g_nccl_allreduce(terms, terms, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, p->comm, st);
  if (rc != 0) {
    set_error("ncclAllReduce of the ELBO table failed on rank %d/%d: %s", p->comm_rank, p->comm_nranks,
              g_nccl_errstr ? g_nccl_errstr(rc) : "?");
    return 1;
  }
  return 0;
}

int plan_check_ready(Plan* p) { return check_ready(p); }
int plan_decoder_forward(Plan* p, const float* mu, const float* lv, const float* eps, const float* z_in, cudaStream_t st) {
  return decoder_forward(p, mu, lv, eps, z_in, st);
}
int plan_allreduce_sum(Plan* p, float* buf, size_t count, cudaStream_t st) { return reduce_terms(p, buf, count, st); }

static int do_decode(Plan* p, const float* z, float* pred, float* mask, float* mean, cudaStream_t st,
                     uint8_t* amax = nullptr) {
  if (decoder_forward(p, nullptr, nullptr, nullptr, z, st)) return 1;
  return launch_recombine(p, pred, mask, mean, p->s.B, st, amax);
}

}  // namespace iod

using namespace iod;

// =========================================================================== C ABI
extern "C" {

IODINE_API int iodine_abi_version(void) { return IODINE_ABI_VERSION; }
IODINE_API const char* iodine_last_error(void) { return g_err; }

static int plan_build(Plan* p, const IodineShape& s);

IODINE_API int iodine_plan_create(const IodineShape* shape, IodinePlan** plan_out) {
  IOD_REQUIRE(shape && plan_out, "iodine_plan_create: null argument");
  const IodineShape& s = *shape;
  IOD_REQUIRE(s.B > 0 && s.K > 0 && s.K <= 16, "unsupported B=%d K=%d (1 <= K <= 16)", s.B, s.K);
  IOD_REQUIRE(s.img_c == 3, "IMG_CHANNELS must be 3 (got %d)", s.img_c);
  IOD_REQUIRE(s.L >= 2 && s.L <= 256, "unsupported DIM_LATENT=%d", s.L);
  IOD_REQUIRE(s.dec_layers >= 1 && s.dec_layers <= IODINE_MAX_LAYERS, "unsupported DEC.CONV_LAYERS=%d", s.dec_layers);
  IOD_REQUIRE(s.ref_layers >= 1 && s.ref_layers <= IODINE_MAX_LAYERS, "unsupported REF.CONV_LAYERS=%d", s.ref_layers);
  IOD_REQUIRE(s.dec_chan == 16 || s.dec_chan == 32 || s.dec_chan == 64, "unsupported DEC.CONV_CHAN=%d", s.dec_chan);
  IOD_REQUIRE(s.ref_chan == 16 || s.ref_chan == 32 || s.ref_chan == 64, "unsupported REF.CONV_CHAN=%d", s.ref_chan);
  IOD_REQUIRE(s.dec_k == 3 || s.dec_k == 5, "unsupported DEC.KERNEL_SIZE=%d", s.dec_k);
  IOD_REQUIRE(s.ref_k == 3 || s.ref_k == 5, "unsupported REF.KERNEL_SIZE=%d", s.ref_k);
  IOD_REQUIRE(s.ref_stride == 1 || s.ref_stride == 2, "unsupported REF.STRIDE=%d", s.ref_stride);
  IOD_REQUIRE(s.H >= 2 * s.dec_k && s.W >= 2 * s.dec_k, "image %dx%d too small for kernel %d", s.H, s.W, s.dec_k);
  IOD_REQUIRE(s.mlp_units >= 1 && s.T >= 0 && s.sigma > 0.f, "bad MLP_UNITS/ITERS/SIGMA");
  IOD_REQUIRE(s.precision == IODINE_FP32 || s.precision == IODINE_BF16 || s.precision == IODINE_FP16 ||
                  s.precision == IODINE_TF32,
              "unsupported precision %d", s.precision);
  IOD_REQUIRE(s.slot_ranks >= 0 && s.slot_ranks <= 16, "bad slot_ranks=%d", s.slot_ranks);
  if (s.slot_ranks > 1) {
    IOD_REQUIRE(s.K % s.slot_ranks == 0, "K-split: SLOTS=%d is not a multiple of slot_ranks=%d", s.K, s.slot_ranks);
    IOD_REQUIRE(s.slot_rank >= 0 && s.slot_rank < s.slot_ranks, "K-split: slot_rank %d outside [0, %d)", s.slot_rank,
                s.slot_ranks);
  }

  Plan* p = new Plan();
  if (plan_build(p, s)) {                       // nothing half-built survives a failed create
    iodine_plan_destroy(reinterpret_cast<IodinePlan*>(p));
    return 1;
  }
  *plan_out = reinterpret_cast<IodinePlan*>(p);
  return 0;
}

static int plan_build(Plan* p, const IodineShape& s_in) {
  p->s = s_in;
  p->K_total = s_in.K;
  if (s_in.slot_ranks > 1) {                    // K-split: from here on the plan's K is its own share of the slots
    p->ks_ranks = s_in.slot_ranks;
    p->ks_rank = s_in.slot_rank;
    p->s.K = s_in.K / s_in.slot_ranks;
  }
  const IodineShape& s = p->s;
  for (int l = 0; l < IODINE_MAX_LAYERS; ++l) { p->ref_wp[l] = nullptr; p->ref_b[l] = nullptr; }
  // (K-split steps contain an NCCL collective: they run eagerly)
  p->graphs = getenv("IODINE_NO_GRAPH") == nullptr && p->ks_ranks == 1;
  IOD_CHECK_CUDA(cudaGetDevice(&p->device));
  IOD_CHECK_CUDA(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));
  p->BK = s.B * s.K; p->HW = s.H * s.W; p->M = s.mlp_units; p->C = s.dec_chan; p->Cr = s.ref_chan;
  p->n_class = s.dec_k * s.dec_k;
  p->ref_h[0] = s.H; p->ref_w[0] = s.W;
  for (int l = 0; l < s.ref_layers; ++l) {
    const int pad = s.ref_k / 2;
    p->ref_h[l + 1] = (p->ref_h[l] + 2 * pad - s.ref_k) / s.ref_stride + 1;
    p->ref_w[l + 1] = (p->ref_w[l] + 2 * pad - s.ref_k) / s.ref_stride + 1;
    IOD_REQUIRE(p->ref_h[l + 1] >= 1 && p->ref_w[l + 1] >= 1, "refine layer %d output is empty", l);
  }
  if (tc_mode(p) && !tc_supported(p)) return 1;
  const size_t C = p->C, L = s.L, M = p->M, Cr = p->Cr, kk = (size_t)s.dec_k * s.dec_k,
               rkk = (size_t)s.ref_k * s.ref_k;
  if (alloc_f(&p->wsum, kk * C * L) || alloc_f(&p->wsumT, kk * C * L) || alloc_f(&p->ptab, (size_t)p->HW * C)) return 1;
  for (int l = 1; l < s.dec_layers; ++l)
    if (alloc_f(&p->dec[l].w, kk * C * C) || alloc_f(&p->dec[l].wt, kk * C * C) || alloc_f(&p->dec[l].b, C)) return 1;
  if (alloc_f(&p->out_w, kk * C * 4) || alloc_f(&p->out_wt, kk * C * 4) || alloc_f(&p->out_b, 4)) return 1;
  if (alloc_f(&p->ref_w0, rkk * 20 * Cr)) return 1;
  for (int l = 0; l < s.ref_layers; ++l) {
    p->ref_wp[l] = nullptr;
    if (l > 0 && alloc_f(&p->ref_wp[l], rkk * Cr * Cr)) return 1;
    if (alloc_f(&p->ref_b[l], Cr)) return 1;
  }
  if (alloc_f(&p->mlp_w, M * Cr) || alloc_f(&p->mlp_b, M) || alloc_f(&p->w_ih, 4 * M * (M + 4 * L)) ||
      alloc_f(&p->w_hh, 4 * M * M) || alloc_f(&p->b_ih, 4 * M) || alloc_f(&p->b_hh, 4 * M) ||
      alloc_f(&p->head_w, 2 * L * M) || alloc_f(&p->head_b, 2 * L) || alloc_f(&p->init_mean, L) ||
      alloc_f(&p->init_logvar, L))
    return 1;
  if (tc_mode(p) && tc_alloc(p)) return 1;
  if (tc_mode(p) && !getenv("IODINE_REFINE_FFMA") && rtc_alloc(p)) return 1;
  p->ws_need = carve(p, nullptr);
  return 0;
}

IODINE_API int iodine_plan_destroy(IodinePlan* plan) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (!p) return 0;
  int cur_dev = -1;                              // the plan's allocations live on ITS device
  cudaGetDevice(&cur_dev);
  if (cur_dev != p->device) cudaSetDevice(p->device);
  cudaFree(p->wsum); cudaFree(p->wsumT); cudaFree(p->ptab);
  for (int l = 0; l < IODINE_MAX_LAYERS; ++l) {
    cudaFree(p->dec[l].w); cudaFree(p->dec[l].wt); cudaFree(p->dec[l].b);
    cudaFree(p->ref_wp[l]); cudaFree(p->ref_b[l]);
  }
  cudaFree(p->out_w); cudaFree(p->out_wt); cudaFree(p->out_b); cudaFree(p->ref_w0);
  cudaFree(p->mlp_w); cudaFree(p->mlp_b); cudaFree(p->w_ih); cudaFree(p->w_hh); cudaFree(p->b_ih);
  cudaFree(p->b_hh); cudaFree(p->head_w); cudaFree(p->head_b); cudaFree(p->init_mean);
  cudaFree(p->init_logvar);
  tc_free(p);
  rtc_free(p);
  train_free(p);
  for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
  for (int i = 0; i < Plan::N_GRAPHS; ++i)
    if (p->graph[i].exec) cudaGraphExecDestroy(p->graph[i].exec);
  if (p->gstream) { cudaStreamDestroy(p->gstream); cudaEventDestroy(p->gev_in); cudaEventDestroy(p->gev_out); }
  const int plan_dev = p->device;
  delete p;
  if (cur_dev >= 0 && cur_dev != plan_dev) cudaSetDevice(cur_dev);
  cudaGetLastError();
  return 0;
}

IODINE_API int iodine_plan_workspace_bytes(const IodinePlan* plan, size_t* bytes_out) {
  IOD_REQUIRE(plan && bytes_out, "null argument");
  *bytes_out = reinterpret_cast<const Plan*>(plan)->ws_need;
  return 0;
}

IODINE_API int iodine_plan_set_workspace(IodinePlan* plan, void* workspace, size_t bytes) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p && workspace, "null argument");
  IOD_REQUIRE(bytes >= p->ws_need, "workspace too small: %zu < %zu", bytes, p->ws_need);
  IOD_REQUIRE(((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");
  p->ws = workspace; p->ws_bytes = bytes;
  for (int i = 0; i < Plan::N_GRAPHS; ++i) {     // captured graphs point into the old workspace
    if (p->graph[i].exec) cudaGraphExecDestroy(p->graph[i].exec);
    p->graph[i] = Plan::EncodeGraph();
  }
  carve(p, (char*)workspace);
  if (tc_mode(p) && tc_on_workspace(p)) return 1;
  return 0;
}

IODINE_API int iodine_plan_set_weights(IodinePlan* plan, const IodineWeights* w, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p && w, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (launch_setup_weights(p, w, st)) return 1;
  if (tc_mode(p) && tc_setup_weights(p, w, st)) return 1;
  if (rtc_enabled(p) && rtc_setup_weights(p, w, st)) return 1;
  p->weights_set = true;
  return 0;
}

IODINE_API int iodine_init_state(IodinePlan* plan, float* post_mean, float* post_logvar, float* lstm_h,
                      float* lstm_c, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  return launch_init_state(p, post_mean, post_logvar, lstm_h, lstm_c, (cudaStream_t)stream);
}

IODINE_API int iodine_refine_step(IodinePlan* plan, const float* x, const float* eps_t, float* post_mean,
                       float* post_logvar, float* lstm_h, float* lstm_c, float* elbo_terms_out,
                       float* aux_out, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  IOD_REQUIRE(x && eps_t && post_mean && post_logvar && lstm_h && lstm_c, "null tensor argument");
  return refine_step(p, x, eps_t, post_mean, post_logvar, lstm_h, lstm_c, elbo_terms_out, aux_out,
                     (cudaStream_t)stream);
}

IODINE_API int iodine_elbo(IodinePlan* plan, const float* x, const float* eps_t, const float* post_mean,
                const float* post_logvar, float* elbo_terms_out, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  IOD_REQUIRE(x && eps_t && post_mean && post_logvar && elbo_terms_out, "null tensor argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (decoder_forward(p, post_mean, post_logvar, eps_t, nullptr, st)) return 1;
  if (launch_mixture(p, x, false, st)) return 1;
  if (launch_recombine(p, p->log_pred, p->log_mask, p->log_mean, 1, st)) return 1;
  if (launch_kl(p, post_mean, post_logvar, st)) return 1;
  terms_kernel<<<1, 32, 0, st>>>(p->accum, elbo_terms_out);
  IOD_LAUNCH_CHECK(p);
  return reduce_terms(p, elbo_terms_out, 2, st);
}

IODINE_API int iodine_encode(IodinePlan* plan, const float* x, const float* eps, float* z_out,
                  float* elbo_terms_out, float* post_out, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  IOD_REQUIRE(x && eps && z_out, "null tensor argument");
  if (do_encode_g(p, x, eps, z_out, elbo_terms_out, post_out, (cudaStream_t)stream)) return 1;
  return reduce_terms(p, elbo_terms_out, (size_t)p->s.T * 2, (cudaStream_t)stream);
}

IODINE_API int iodine_plan_set_comm(IodinePlan* plan, void* nccl_comm, int32_t rank, int32_t nranks) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p, "null plan");
  if (!nccl_comm) {
    p->comm = nullptr; p->comm_rank = 0; p->comm_nranks = 1;
    return 0;
  }
  IOD_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "iodine_plan_set_comm: rank %d outside [0, %d)", rank, nranks);
  IOD_REQUIRE(p->ks_ranks == 1 || (nranks == p->ks_ranks && rank == p->ks_rank),
              "iodine_plan_set_comm: a K-split plan (slot_rank %d of %d) needs a communicator of exactly those ranks "
              "(got rank %d of %d)", p->ks_rank, p->ks_ranks, rank, nranks);
  if (resolve_nccl()) return 1;
  p->comm = nccl_comm; p->comm_rank = rank; p->comm_nranks = nranks;
  return 0;
}

IODINE_API int iodine_decode(IodinePlan* plan, const float* z, float* pred_out, float* mask_out, float* mean_out,
                  void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  IOD_REQUIRE(z, "null tensor argument");
  return do_decode(p, z, pred_out, mask_out, mean_out, (cudaStream_t)stream);
}

IODINE_API int iodine_reconstruct(IodinePlan* plan, const float* x, const float* eps, float* pred_out,
                       float* mask_out, float* mean_out, float* z_out, float* elbo_terms_out,
                       void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  IOD_REQUIRE(x && eps, "null tensor argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (do_encode_g(p, x, eps, p->st_z, elbo_terms_out, nullptr, st)) return 1;
  if (reduce_terms(p, elbo_terms_out, (size_t)p->s.T * 2, st)) return 1;
  if (z_out)
    IOD_CHECK_CUDA(cudaMemcpyAsync(z_out, p->st_z, (size_t)p->BK * p->s.L * sizeof(float),
                                   cudaMemcpyDeviceToDevice, st));
  return do_decode(p, p->st_z, pred_out, mask_out, mean_out, st);
}

IODINE_API int iodine_plan_last_elbo_image0(IodinePlan* plan, float* pred0, float* mask0, float* mean0, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t HW = p->HW, K = p->K_total;
  if (pred0) IOD_CHECK_CUDA(cudaMemcpyAsync(pred0, p->log_pred, 3 * HW * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (mask0) IOD_CHECK_CUDA(cudaMemcpyAsync(mask0, p->log_mask, K * HW * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (mean0) IOD_CHECK_CUDA(cudaMemcpyAsync(mean0, p->log_mean, K * 3 * HW * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

IODINE_API int iodine_reconstruct_host_async(IodinePlan* plan, const float* x_host, const float* eps_host,
                                  float* pred_host, float* mask_host, float* mean_host, float* z_host,
                                  float* elbo_terms_host, void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  IOD_REQUIRE(x_host && eps_host, "null tensor argument");
  cudaStream_t st = (cudaStream_t)stream;
  const IodineShape& s = p->s;
  const size_t HW = p->HW, BK = p->BK;
  IOD_CHECK_CUDA(cudaMemcpyAsync(p->hx, x_host, (size_t)s.B * 3 * HW * sizeof(float), cudaMemcpyHostToDevice, st));
  IOD_CHECK_CUDA(cudaMemcpyAsync(p->heps, eps_host, (size_t)(s.T + 1) * BK * s.L * sizeof(float), cudaMemcpyHostToDevice, st));
  if (do_encode_g(p, p->hx, p->heps, p->st_z, p->st_terms, nullptr, st)) return 1;
  if (reduce_terms(p, p->st_terms, (size_t)p->s.T * 2, st)) return 1;
  if (do_decode(p, p->st_z, pred_host ? p->hpred : nullptr, mask_host ? p->hmask : nullptr,
                mean_host ? p->hmean : nullptr, st))
    return 1;
  if (pred_host) IOD_CHECK_CUDA(cudaMemcpyAsync(pred_host, p->hpred, (size_t)s.B * 3 * HW * sizeof(float), cudaMemcpyDeviceToHost, st));
  const size_t BKt = (size_t)s.B * p->K_total;
  if (mask_host) IOD_CHECK_CUDA(cudaMemcpyAsync(mask_host, p->hmask, BKt * HW * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (mean_host) IOD_CHECK_CUDA(cudaMemcpyAsync(mean_host, p->hmean, BKt * 3 * HW * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (z_host) IOD_CHECK_CUDA(cudaMemcpyAsync(z_host, p->st_z, BK * s.L * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (elbo_terms_host) IOD_CHECK_CUDA(cudaMemcpyAsync(elbo_terms_host, p->st_terms, (size_t)s.T * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  return 0;
}

IODINE_API int iodine_evaluate_host_async(IodinePlan* plan, const float* x_host, const float* eps_host,
                               float* pred_host, uint8_t* argmax_host, float* z_host, float* elbo_terms_host,
                               void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  if (check_ready(p)) return 1;
  IOD_REQUIRE(x_host && eps_host, "null tensor argument");
  cudaStream_t st = (cudaStream_t)stream;
  const IodineShape& s = p->s;
  const size_t HW = p->HW, BK = p->BK;
  IOD_CHECK_CUDA(cudaMemcpyAsync(p->hx, x_host, (size_t)s.B * 3 * HW * sizeof(float), cudaMemcpyHostToDevice, st));
  IOD_CHECK_CUDA(cudaMemcpyAsync(p->heps, eps_host, (size_t)(s.T + 1) * BK * s.L * sizeof(float), cudaMemcpyHostToDevice, st));
  if (do_encode_g(p, p->hx, p->heps, p->st_z, p->st_terms, nullptr, st)) return 1;
  if (reduce_terms(p, p->st_terms, (size_t)p->s.T * 2, st)) return 1;
  if (do_decode(p, p->st_z, pred_host ? p->hpred : nullptr, nullptr, nullptr, st, argmax_host ? p->hamax : nullptr)) return 1;
  if (pred_host) IOD_CHECK_CUDA(cudaMemcpyAsync(pred_host, p->hpred, (size_t)s.B * 3 * HW * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (argmax_host) IOD_CHECK_CUDA(cudaMemcpyAsync(argmax_host, p->hamax, (size_t)s.B * HW, cudaMemcpyDeviceToHost, st));
  if (z_host) IOD_CHECK_CUDA(cudaMemcpyAsync(z_host, p->st_z, BK * s.L * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (elbo_terms_host) IOD_CHECK_CUDA(cudaMemcpyAsync(elbo_terms_host, p->st_terms, (size_t)s.T * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  return 0;
}

IODINE_API int iodine_evaluate_host(IodinePlan* plan, const float* x_host, const float* eps_host, float* pred_host,
                         uint8_t* argmax_host, float* z_host, float* elbo_terms_host, void* stream) {
  if (iodine_evaluate_host_async(plan, x_host, eps_host, pred_host, argmax_host, z_host, elbo_terms_host, stream)) return 1;
  IOD_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

IODINE_API int iodine_reconstruct_host(IodinePlan* plan, const float* x_host, const float* eps_host,
                            float* pred_host, float* mask_host, float* mean_host, float* z_host,
                            float* elbo_terms_host, void* stream) {
  if (iodine_reconstruct_host_async(plan, x_host, eps_host, pred_host, mask_host, mean_host, z_host, elbo_terms_host, stream))
    return 1;
  IOD_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

IODINE_API int iodine_debug_read(IodinePlan* plan, const char* name, void* dst, size_t dst_bytes, size_t* bytes_out,
                      void* stream) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p && name && p->ws, "null argument / workspace not set");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t BK = p->BK, HW = p->HW;
  const void* src = nullptr;
  size_t bytes = 0;
  bool act_view = false;
  int act_idx = 0;
  if (!strcmp(name, "out4")) { src = p->out4; bytes = BK * HW * 4 * sizeof(float); }
  else if (!strcmp(name, "seed4")) {
    src = p->seed4; bytes = BK * HW * 4 * sizeof(float);
    if (tc_mode(p)) {
      if (bytes_out) *bytes_out = bytes;
      if (!dst) return 0;
      IOD_REQUIRE(dst_bytes >= bytes, "destination too small for %s: %zu < %zu", name, dst_bytes, bytes);
      return tc_export_seed(p, p->seed4, (float*)dst, st);
    }
  }
  else if (!strcmp(name, "dz")) { src = p->dz; bytes = BK * p->s.L * sizeof(float); }
  else if (!strcmp(name, "z")) { src = p->z; bytes = BK * p->s.L * sizeof(float); }
  else if (!strcmp(name, "G")) { src = p->G; bytes = BK * p->n_class * p->C * sizeof(float); }
  else if (!strcmp(name, "pool")) { src = p->pool; bytes = BK * p->Cr * sizeof(float); }
  else if (!strcmp(name, "stats")) { src = p->stats; bytes = BK * 8 * sizeof(double); }
  else if (!strcmp(name, "auxs")) { src = p->auxs; bytes = BK * HW * 12 * sizeof(float); }
  else if (!strncmp(name, "act", 3) || !strncmp(name, "gbuf", 4)) {
    const bool is_g = name[0] == 'g';
    act_idx = atoi(name + (is_g ? 4 : 3));
    IOD_REQUIRE(act_idx >= 0 && act_idx < (is_g ? 2 : p->s.dec_layers), "bad buffer index in %s", name);
    src = is_g ? p->gbuf[act_idx] : p->act[act_idx];
    bytes = BK * HW * p->C * sizeof(float);
    act_view = true;
  } else {
    set_error("iodine_debug_read: unknown buffer '%s'", name);
    return 1;
  }
  if (bytes_out) *bytes_out = bytes;
  if (!dst) return 0;
  IOD_REQUIRE(dst_bytes >= bytes, "destination too small for %s: %zu < %zu", name, dst_bytes, bytes);
  if (act_view && tc_mode(p)) return tc_export_f32(p, src, (float*)dst, BK * HW * p->C, st);
  IOD_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

IODINE_API int iodine_plan_profile(IodinePlan* plan, int enable) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p, "null plan");
  p->profiling = enable;
  p->prof_used = 0;
  return 0;
}

IODINE_API int iodine_plan_profile_read(IodinePlan* plan, double* ms_total_out, uint64_t* launches_out) {
  Plan* p = reinterpret_cast<Plan*>(plan);
  IOD_REQUIRE(p && ms_total_out && launches_out, "null argument");
  double tot = 0.0;
  for (size_t i = 0; i + 1 < p->prof_used; i += 2) {
    IOD_CHECK_CUDA(cudaEventSynchronize(p->prof_events[i + 1]));
    float ms = 0.f;
    IOD_CHECK_CUDA(cudaEventElapsedTime(&ms, p->prof_events[i], p->prof_events[i + 1]));
    tot += ms;
  }
  *ms_total_out = tot;
  *launches_out = p->prof_used / 2;
  p->prof_used = 0;
  return 0;
}

IODINE_API int iodine_plan_launch_count(const IodinePlan* plan, uint64_t* count_out) {
  IOD_REQUIRE(plan && count_out, "null argument");
  *count_out = reinterpret_cast<const Plan*>(plan)->launches;
  return 0;
}

}  // extern "C"
