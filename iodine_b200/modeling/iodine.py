"""IODINE -- host-side mirror of the reference model class (lib/modeling/iodine.py:7-408).

Same constructor (``IODINE(ARCH)``), same methods (``encode / decode / reconstruct / elbo``,
attribute ``sigma``) and the same state_dict keys and parameter shapes, so reference
checkpoints load unchanged (lib/utils/checkpoint.py:68).  The sub-modules below are
PARAMETER CONTAINERS only: all arithmetic of the refinement loop runs in the hand-written
CUDA library behind ``iodine_b200.engine.RefinementEngine`` (C ABI, include/iodine_b200.h).
There is no PyTorch/CPU fallback -- on a non-CUDA device the methods raise.
"""
import torch
from torch import nn

from .. import _cabi
from ..engine import RefinementEngine
from ..utils.vis_logger import logger


class _MultiLayerConv(nn.Module):
    """Parameter container for reference MultiLayerConv (iodine.py:570-594)."""

    def __init__(self, dim_in, dim_out, n_layers, kernel_size, stride=1):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(n_layers):
            self.layers.append(nn.Conv2d(dim_in, dim_out, kernel_size=kernel_size,
                                         padding=kernel_size // 2, stride=stride))
            dim_in = dim_out


class _MLP(nn.Module):
    """Parameter container for reference MLP (iodine.py:543-567)."""

    def __init__(self, dim_in, dim_out, n_layers):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(n_layers):
            self.layers.append(nn.Linear(dim_in, dim_out))
            dim_in = dim_out


class _RefinementNetwork(nn.Module):
    """Parameter container for reference RefinementNetwork (iodine.py:446-464); creation
    order matches so that a seeded default init reproduces the reference's weights."""

    def __init__(self, dim_in, dim_conv, dim_hidden, dim_out, n_layers, kernel_size, stride):
        super().__init__()
        self.mlc = _MultiLayerConv(dim_in, dim_conv, n_layers, kernel_size, stride=stride)
        self.mlp = _MLP(dim_conv, dim_hidden, n_layers=1)
        self.lstm = nn.LSTMCell(dim_hidden + 4 * dim_out, dim_hidden)
        self.mean_update = nn.Linear(dim_hidden, dim_out)
        self.logvar_update = nn.Linear(dim_hidden, dim_out)


class _Decoder(nn.Module):
    """Parameter container for reference Decoder (iodine.py:412-423)."""

    def __init__(self, dim_in, dim_hidden, n_layers, kernel_size, img_size):
        super().__init__()
        self.mlc = _MultiLayerConv(dim_in + 2, dim_hidden, n_layers, kernel_size)
        self.conv = nn.Conv2d(dim_hidden, 4, kernel_size=kernel_size, stride=1,
                              padding=kernel_size // 2)
        self.img_size = img_size


class _Gaussian(nn.Module):
    """Posterior state holder (reference Gaussian, iodine.py:596-604)."""

    def __init__(self, dim_latent):
        super().__init__()
        self.mean = None
        self.logvar = None
        self.init_mean = nn.Parameter(data=torch.zeros(dim_latent))
        self.init_logvar = nn.Parameter(data=torch.zeros(dim_latent))


# image-shaped encodings -> their channels in the full stack (reference code order, iodine.py:277-340)
_SPATIAL_ENCODINGS = (('image', (0, 1, 2)), ('means', (3, 4, 5)), ('mask', (6,)), ('mask_logits', (7,)),
                      ('mask_posterior', (8,)), ('grad_means', (9, 10, 11)), ('grad_mask', (12,)),
                      ('likelihood', (13,)), ('leave_one_out_likelihood', (14,)), ('coordinate', (15, 16)))

ALL_ENCODINGS = ('posterior', 'grad_post', 'image', 'means', 'mask', 'mask_logits',
                 'mask_posterior', 'grad_means', 'grad_mask', 'likelihood',
                 'leave_one_out_likelihood', 'coordinate')


class _TrainStep(torch.autograd.Function):
    """Graph node of ``IODINE.forward``: the library has already produced dLoss/dParameter for every parameter when
    the loss is returned; backward scales them by the incoming gradient (``loss.mean()`` of a 0-dim loss: 1)."""

    @staticmethod
    def forward(ctx, model, x, eps, names, *params):
        loss, grads = model._train_step(x, eps)
        ctx.grads = [grads[k] for k in names]
        return loss

    @staticmethod
    def backward(ctx, gout):
        return (None, None, None, None) + tuple(gout * g for g in ctx.grads)


class IODINE(nn.Module):
    def __init__(self, ARCH, precision='fp32'):
        nn.Module.__init__(self)
        self.arch = ARCH
        self.dim_latent = ARCH.DIM_LATENT
        self.n_iters = ARCH.ITERS
        self.K = ARCH.SLOTS
        self.encodings = list(ARCH.ENCODING)
        self.img_channels = ARCH.IMG_CHANNELS
        self.img_size = ARCH.IMG_SIZE
        self.sigma = ARCH.SIGMA
        self.use_layernorm = ARCH.LAYERNORM
        self.use_stop_gradient = ARCH.STOP_GRADIENT
        if precision not in _cabi.PRECISIONS:
            raise ValueError('precision must be one of %s' % sorted(_cabi.PRECISIONS))
        self.precision = precision
        unknown = [e for e in self.encodings if e not in ALL_ENCODINGS]
        if unknown:
            raise ValueError('unknown entries in ARCH.ENCODING: %s' % unknown)
        if 'posterior' not in self.encodings or 'grad_post' not in self.encodings:
            # the reference's LSTMCell is built for M + 4L inputs whatever the list says (iodine.py:462), so its
            # forward fails on the concatenation without both latent encodings; refuse at construction instead
            raise ValueError("ARCH.ENCODING must contain 'posterior' and 'grad_post' (the reference's LSTMCell "
                             'input is MLP_UNITS + 4*DIM_LATENT wide, lib/modeling/iodine.py:462)')
        # channels of the FULL 17-channel stack (reference code order, iodine.py:277-340) that this list selects,
        # in the order get_input_encoding() concatenates them
        self._enc_channels = [c for name, chans in _SPATIAL_ENCODINGS if name in self.encodings for c in chans]
        if not self._enc_channels:
            raise ValueError('ARCH.ENCODING selects no image-shaped encoding')
        if self.img_channels != 3:
            raise NotImplementedError('ARCH.IMG_CHANNELS must be 3')

        input_size, lambda_size = self.get_input_size()
        self.refine = _RefinementNetwork(
            input_size, ARCH.REF.CONV_CHAN, ARCH.REF.MLP_UNITS, ARCH.DIM_LATENT,
            ARCH.REF.CONV_LAYERS, kernel_size=ARCH.REF.KERNEL_SIZE, stride=ARCH.REF.STRIDE)
        self.decoder = _Decoder(dim_in=ARCH.DIM_LATENT, dim_hidden=ARCH.DEC.CONV_CHAN,
                                n_layers=ARCH.DEC.CONV_LAYERS, kernel_size=ARCH.DEC.KERNEL_SIZE,
                                img_size=self.img_size)
        self.posterior = _Gaussian(self.dim_latent)

        # per-call state, as on the reference instance (iodine.py:37-52)
        self.lstm_hidden = None
        self.z = None
        self.mean = None
        self.mask = None
        self.elbo_terms = None      # [T,2]: (sum_b log-lik, sum_b KL) per refinement step
        self._engines = {}
        self._comm = None           # (NcclComm, rank, nranks) installed by set_comm()
        self._slot_split = None     # (rank, nranks) installed by set_slot_split(): this replica owns K/nranks slots
        self._weights_sig = {}
        self._weights_epoch = 0
        self.max_images_per_call = None   # None = automatic (fit the workspace in free HBM)
        # with a communicator installed the engine returns ELBO sums over ALL ranks' images; means are then taken
        # over this many images (set by iodine_b200.parallel.SlotShard; None = the local batch)
        self.global_batch = None

    # ------------------------------------------------------------------ bookkeeping
    def get_input_size(self):
        """reference iodine.py:345-374: (channels of the refinement input, width of the latent vector)."""
        return len(self._enc_channels), 4 * self.dim_latent

    def _engine_state_dict(self):
        """state_dict in the shapes the engine takes (the full 17-channel stack).  A partial ARCH.ENCODING list
        means a first refinement conv with fewer input channels: its weight is scattered into the full layout with
        zeros for the channels the list leaves out, which is arithmetically the reference's smaller convolution
        (the engine still evaluates all 17 channels; a NaN in an unselected channel -- the reference's own
        un-stabilised mask_posterior can produce one -- would leak through the zero weight)."""
        sd = self.state_dict()
        if len(self._enc_channels) != 17:
            w = sd['refine.mlc.layers.0.weight']
            full = w.new_zeros(w.shape[0], 17, w.shape[2], w.shape[3])
            full[:, self._enc_channels] = w
            sd = dict(sd)
            sd['refine.mlc.layers.0.weight'] = full
        return sd

    def _device(self):
        return self.posterior.init_mean.device

    def _sig(self):
        return (self._weights_epoch,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_weights(self):
        """Force the engines to repack the weights on their next call.  Parameter changes that go through autograd-
        visible in-place ops (optimizers, ``load_state_dict``, ``p.mul_``) are detected automatically from the
        tensors' version counters; edits through ``p.data`` do not bump those counters and need this call."""
        self._weights_epoch += 1

    def load_state_dict(self, *args, **kw):
        out = super().load_state_dict(*args, **kw)
        self.invalidate_weights()
        return out

    def _apply(self, fn, *args, **kw):
        out = super()._apply(fn, *args, **kw)         # .to() / .cuda() / .half(): new storages
        if hasattr(self, '_weights_epoch'):
            self.invalidate_weights()
        return out

    def _mean_batch(self, B):
        """number of images the engine's batch sums cover: all ranks' when a communicator is installed"""
        if self._comm is not None and self.global_batch:
            return int(self.global_batch)
        return B

    def _engine(self, B):
        dev = self._device()
        key = (int(B), str(dev), self.precision, self._slot_split)
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= 3:          # a few plans (full chunk, tail chunk, a second batch size)
                old = next(iter(self._engines))
                self._engines.pop(old).close()
                self._weights_sig.pop(old, None)
            eng = RefinementEngine(self.arch, B, dev, self.precision, slot_split=self._slot_split)
            if self._comm is not None:
                eng.set_comm(*self._comm)
            self._engines[key] = eng
        sig = self._sig()
        if self._weights_sig.get(key) != sig:
            eng.set_weights(self._engine_state_dict())
            self._weights_sig[key] = sig
        return eng

    def set_comm(self, comm, rank=0, nranks=1):
        """Multi-GPU (replaces DataParallel, lib/modeling/build.py:11-12): with an ``iodine_b200.parallel.NcclComm``
        installed, ``elbo_terms`` of encode()/reconstruct() are the sums over ALL ranks' images -- one in-stream
        ncclAllReduce per engine call (``iodine_plan_set_comm``).  Every rank must then make the same number of
        engine calls (``SlotShard`` checks this).  ``comm=None`` removes it."""
        self._comm = None if comm is None else (comm, int(rank), int(nranks))
        for eng in self._engines.values():
            eng.set_comm(comm, rank, nranks)

    def set_slot_split(self, comm, rank, nranks):
        """K-split (SURVEY.md 8e fallback, for a batch smaller than the number of GPUs): this replica owns the slots
        ``[rank*K/nranks, (rank+1)*K/nranks)`` of every image.  ``eps`` passed to encode()/reconstruct() and the
        ``z`` / posterior they leave then carry K/nranks slots; ``pred, mask, mean`` are complete on every rank;
        the library all-gathers the decoder's 4-channel output once per elbo() evaluation (``iodine_plan_set_comm``
        on a K-split plan).  ``iodine_b200.parallel.KSplit`` drives this from a global batch."""
        if self.K % int(nranks):
            raise ValueError('K-split: SLOTS=%d is not a multiple of %d ranks' % (self.K, nranks))
        for eng in self._engines.values():
            eng.close()
        self._engines, self._weights_sig = {}, {}
        self._slot_split = (int(rank), int(nranks)) if int(nranks) > 1 else None
        self._comm = None if comm is None else (comm, int(rank), int(nranks))

    def _local_slots(self):
        return self.K // (self._slot_split[1] if self._slot_split else 1)

    def _chunk(self, B):
        if self.max_images_per_call:
            return min(B, int(self.max_images_per_call))
        dev = self._device()
        free, _ = torch.cuda.mem_get_info(dev)
        a = self.arch
        eb = 2 if self.precision in ('bf16', 'fp16') else 4
        per_img = self.K * a.IMG_SIZE * a.IMG_SIZE * (
            (a.DEC.CONV_LAYERS + 2) * a.DEC.CONV_CHAN * eb + (8 + 12 + 20 + 4) * 4
            + a.REF.CONV_CHAN * 4 // 2)
        cached = sum(e.workspace_bytes for e in self._engines.values())
        return max(1, min(B, int(0.8 * (free + cached) // max(per_img, 1))))

    def _noise(self, B, eps):
        T, K, L = self.n_iters, self._local_slots(), self.dim_latent
        if eps is None:
            # reference: torch.randn_like on the model's device, T+1 draws (iodine.py:632)
            return torch.randn(T + 1, B, K, L, device=self._device(), dtype=torch.float32)
        assert tuple(eps.shape) == (T + 1, B, K, L), 'eps must be [T+1,B,K,L]'
        return eps

    def _require_cuda(self, x):
        if self._device().type != 'cuda':
            raise _cabi.IodineError(
                'iodine_b200.IODINE runs on CUDA only (model is on %s). The reference CPU path '
                'is not re-implemented here; there is no fallback.' % self._device())
        return x.to(self._device())

    # ------------------------------------------------------------------ public API
    def decode(self, z):
        """reference iodine.py:59-71: z[B,K,L] -> pred[B,3,H,W], mask[B,K,1,H,W], mean[B,K,3,H,W]"""
        z = self._require_cuda(z)
        B = z.shape[0]
        outs = [self._engine(b1 - b0).decode(z[b0:b1]) for b0, b1 in self._spans(B)]
        return tuple(torch.cat(t, dim=0) if len(t) > 1 else t[0] for t in zip(*outs))

    def _spans(self, B):
        c = self._chunk(B)
        return [(b0, min(B, b0 + c)) for b0 in range(0, B, c)]

    @torch.no_grad()
    def encode(self, x, eps=None):
        """reference iodine.py:73-105: x[B,3,H,W] -> z[B,K,L].  ``eps`` ([T+1,B,K,L]) injects
        the noise the reference draws with torch.randn_like; default = fresh device noise."""
        x = self._require_cuda(x)
        B = x.shape[0]
        eps = self._noise(B, eps)
        zs, terms, posts = [], 0, []
        for b0, b1 in self._spans(B):
            eng = self._engine(b1 - b0)
            z, t, post = eng.encode(x[b0:b1], eps[:, b0:b1])
            if b0 == 0 and self.n_iters > 0:
                self._log_last_elbo(x, eng)
            zs.append(z)
            posts.append(post)
            terms = terms + t
        self.elbo_terms = terms
        post = torch.cat(posts, dim=1)
        self.posterior.mean, self.posterior.logvar = post[0], post[1]
        self.z = torch.cat(zs, dim=0)
        self._log_scalars(B)
        return self.z

    @torch.no_grad()
    def reconstruct(self, x, eps=None):
        """reference iodine.py:107-112: encode + decode."""
        x = self._require_cuda(x)
        B = x.shape[0]
        eps = self._noise(B, eps)
        outs, terms = [], 0
        for b0, b1 in self._spans(B):
            eng = self._engine(b1 - b0)
            pred, mask, mean, z, t = eng.reconstruct(x[b0:b1], eps[:, b0:b1])
            if b0 == 0 and self.n_iters > 0:
                self._log_last_elbo(x, eng)
            outs.append((pred, mask, mean, z))
            terms = terms + t
        pred, mask, mean, z = (torch.cat(t, dim=0) if len(t) > 1 else t[0] for t in zip(*outs))
        self.elbo_terms, self.z, self.mean, self.mask = terms, z, mean, mask
        self._log_scalars(B)
        return pred, mask, mean

    @torch.no_grad()
    def elbo(self, x, eps=None):
        """reference iodine.py:161-241: single-pass ELBO (mean over batch) for the current
        posterior (``self.posterior.mean/logvar``; the learnt initial posterior if unset)."""
        x = self._require_cuda(x)
        B, K, L = x.shape[0], self._local_slots(), self.dim_latent
        mu, lv = self.posterior.mean, self.posterior.logvar
        if mu is None or mu.shape[0] != B:
            mu = self.posterior.init_mean.detach()[None, None].repeat(B, K, 1)
            lv = self.posterior.init_logvar.detach()[None, None].repeat(B, K, 1)
        if eps is None:
            eps = torch.randn(B, K, L, device=self._device())
        terms = 0
        for b0, b1 in self._spans(B):
            eng = self._engine(b1 - b0)
            terms = terms + eng.elbo_terms(x[b0:b1], eps[b0:b1], mu[b0:b1], lv[b0:b1])
            if b0 == 0:
                self._log_last_elbo(x, eng)
        nb = self._mean_batch(B)
        logger.update(kl=terms[1] / nb, likelihood=terms[0] / nb)
        return (terms[0] - terms[1]) / nb

    def elbo_per_step(self, B=None):
        """ELBO of every refinement step of the last encode()/reconstruct() call, as the
        reference would have returned from elbo() inside the loop (mean over batch)."""
        t = self.elbo_terms
        B = self._mean_batch(B or self.z.shape[0])
        return (t[:, 0] - t[:, 1]) / B

    def forward(self, x, eps=None):
        """Training objective (reference iodine.py:115-158): ``-sum_i (i+1)/(T+1) * elbo_i`` over the T in-loop ELBOs
        and the final one (151-153, 158), as a 0-dim tensor.

        With autograd enabled the tensor carries a graph node whose backward hands every parameter its gradient, so
        the reference's training loop (``lib/engine/train.py:60-65``: ``loss = model(data); loss = loss.mean();
        optimizer.zero_grad(); loss.backward(); optimizer.step()``) runs unchanged.  Loss AND gradients are computed
        inside this call by the library's training step (``iodine_train_step``, csrc/train.cu: hand-written
        weight-gradient / LSTM-backward kernels, no autograd through the loop); ``backward()`` only scales and
        delivers them.  Under ``torch.no_grad()`` the value comes from the inference kernels alone."""
        x = self._require_cuda(x)
        B, T = x.shape[0], self.n_iters
        eps = self._noise(B, eps)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if self._slot_split is not None:
                raise _cabi.IodineError('training is not available on a K-split replica')
            names = [k for k, _ in self.named_parameters()]
            return _TrainStep.apply(self, x, eps, names, *[p for _, p in self.named_parameters()])
        with torch.no_grad():
            self.encode(x, eps=eps)                                   # T in-loop ELBOs + final posterior
            elbos = list(self.elbo_per_step(B))
            final = 0
            for b0, b1 in self._spans(B):
                eng = self._engine(b1 - b0)
                final = final + eng.elbo_terms(x[b0:b1], eps[T, b0:b1], self.posterior.mean[b0:b1],
                                               self.posterior.logvar[b0:b1])
                if b0 == 0:
                    self._log_last_elbo(x, eng)
            nb = self._mean_batch(B)
            elbos.append((final[0] - final[1]) / nb)
            logger.update(kl=final[1] / nb, likelihood=final[0] / nb)          # the last elbo() call is the final one
            loss = 0
            for i, e in enumerate(elbos):
                loss = loss + (i + 1) / len(elbos) * e
            logger.update(init_mean=self.posterior.init_mean.detach().mean(),
                          init_logvar=self.posterior.init_logvar.detach().mean())
            return -loss

    @torch.no_grad()
    def _train_step(self, x, eps):
        """loss (0-dim) and {state_dict key: dLoss/dParameter} through ``iodine_train_step``; the batch is chunked like
        the inference calls (gradients and loss are additive over images once every mean divides by the full batch)."""
        B = x.shape[0]
        nb = self._mean_batch(B)
        sd = self._engine_state_dict()
        total, loss, terms = None, 0, 0
        for b0, b1 in self._spans(B):
            eng = self._engine(b1 - b0)
            grads = {k: torch.empty_like(v, dtype=torch.float32, device=self._device()) for k, v in sd.items()}
            l, t = eng.train_step(x[b0:b1], eps[:, b0:b1], grads, global_batch=nb)
            if b0 == 0:
                self._log_last_elbo(x, eng)
            loss, terms = loss + l, terms + t
            if total is None:
                total = grads
            else:
                for k in total:
                    total[k] += grads[k]
        if len(self._enc_channels) != 17:                             # partial ARCH.ENCODING: the selected channels
            total['refine.mlc.layers.0.weight'] = total['refine.mlc.layers.0.weight'][:, self._enc_channels].contiguous()
        self.elbo_terms = terms
        logger.update(kl=terms[-1, 1] / nb, likelihood=terms[-1, 0] / nb)       # the last elbo() call is the final one
        logger.update(init_mean=self.posterior.init_mean.detach().mean(),
                      init_logvar=self.posterior.init_logvar.detach().mean())
        return loss, total

    # ------------------------------------------------------------------ side channel (A9)
    # The reference writes to the logger inside EVERY elbo() call (iodine.py:225-239), each write replacing the
    # previous one, so what a consumer finds after encode()/reconstruct() are the quantities of the LAST in-loop
    # elbo() -- refinement step T-1, not the final decode -- and after forward()/elbo() those of that last call.
    # The engine keeps image 0 of its last elbo() evaluation for exactly this (iodine_plan_last_elbo_image0).
    def _log_scalars(self, B):
        if self.elbo_terms is not None and self.elbo_terms.numel():
            nb = self._mean_batch(B)
            logger.update(kl=self.elbo_terms[-1, 1] / nb, likelihood=self.elbo_terms[-1, 0] / nb)

    def _log_last_elbo(self, x, eng):
        """``eng`` has just processed the chunk that holds image 0"""
        pred0, mask0, mean0 = eng.last_elbo_image0()
        logger.update(image=x[0], pred=pred0)
        logger.update(**{'mask_{}'.format(i): mask0[i] for i in range(self.K)})
        logger.update(**{'pred_{}'.format(i): mean0[i] for i in range(self.K)})

    def state_for_debug(self, B):
        return self._engine(B)
