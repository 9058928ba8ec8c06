"""RefinementEngine: owns one C-ABI plan (``include/iodine_b200.h``) plus its workspace and
feeds it torch CUDA tensors.  torch is plumbing here (device memory + current stream); all
arithmetic of the refinement loop happens inside ``libiodine_b200.so``.
"""
import ctypes as C

import torch

from . import _cabi


def _ptr(t):
    if t is None:
        return None
    assert t.is_contiguous(), 'tensor must be contiguous'
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def shape_from_arch(arch, B, precision='fp32', slot_split=None):
    """IodineShape from a ``cfg.ARCH``-like object (fields read at reference
    lib/modeling/iodine.py:10-32).  ``slot_split = (rank, nranks)``: K-split plan (this rank owns K/nranks slots)."""
    s = _cabi.IodineShape()
    s.B, s.K, s.L = int(B), int(arch.SLOTS), int(arch.DIM_LATENT)
    s.H = s.W = int(arch.IMG_SIZE)
    s.T = int(arch.ITERS)
    s.img_c = int(arch.IMG_CHANNELS)
    s.dec_layers, s.dec_chan, s.dec_k = (int(arch.DEC.CONV_LAYERS), int(arch.DEC.CONV_CHAN),
                                         int(arch.DEC.KERNEL_SIZE))
    s.ref_layers, s.ref_chan, s.ref_k, s.ref_stride = (
        int(arch.REF.CONV_LAYERS), int(arch.REF.CONV_CHAN), int(arch.REF.KERNEL_SIZE),
        int(arch.REF.STRIDE))
    s.mlp_units = int(arch.REF.MLP_UNITS)
    s.layernorm = 1 if arch.LAYERNORM else 0
    s.sigma = float(arch.SIGMA)
    s.precision = _cabi.PRECISIONS[precision]
    if slot_split is not None and int(slot_split[1]) > 1:
        s.slot_rank, s.slot_ranks = int(slot_split[0]), int(slot_split[1])
    return s


class RefinementEngine:
    """One plan = one (architecture, per-call batch B, precision, device)."""

    def __init__(self, arch, B, device, precision='fp32', slot_split=None):
        """``slot_split = (rank, nranks)``: K-split (include/iodine_b200.h, IodineShape.slot_ranks) -- this engine
        owns K/nranks slots of every image: eps / z / posterior / LSTM state carry ``self.K`` = K/nranks slots,
        decode()/reconstruct() return all ``self.K_total`` slots.  Needs ``set_comm`` before the first step."""
        self.lib = _cabi.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _cabi.IodineError('the refinement engine runs on CUDA devices only (got %s); '
                                    'there is no CPU fallback' % self.device)
        self.arch, self.B, self.precision = arch, int(B), precision
        self.shape = shape_from_arch(arch, B, precision, slot_split)
        self.K_total = self.shape.K
        self.K = self.shape.K // max(1, self.shape.slot_ranks)
        self.L, self.T = self.shape.L, self.shape.T
        self.H, self.W, self.M = self.shape.H, self.shape.W, self.shape.mlp_units
        self._plan = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_plan_create(C.byref(self.shape), C.byref(self._plan)))
            need = C.c_size_t()
            _cabi.check(self.lib.iodine_plan_workspace_bytes(self._plan, C.byref(need)))
            self.workspace_bytes = need.value
            self._ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
            base = (self._ws.data_ptr() + 1023) // 1024 * 1024
            _cabi.check(self.lib.iodine_plan_set_workspace(self._plan, C.c_void_p(base), need.value))
        self._weights_key = None
        self._keep = None
        # persistent I/O buffers of encode()/reconstruct(): stable device pointers let the library replay the
        # whole call as one CUDA graph (plan.cu: do_encode_g); user tensors are copied in / cloned out
        B, K, L, T = self.B, self.K, self.L, self.T
        self._x_in = self._new(B, 3, self.H, self.W)
        self._eps_in = self._new(T + 1, B, K, L)
        self._z = self._new(B, K, L)
        self._terms = self._new(max(T, 1), 2)
        self._post = self._new(2, B, K, L)

    def close(self):
        if self._plan:
            self.lib.iodine_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def set_weights(self, sd):
        """sd: reference state_dict keys -> fp32 CUDA tensors."""
        keep = {}

        def g(key):
            t = sd[key].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep[key] = t
            return t.data_ptr()

        w = _cabi.IodineWeights()
        for i in range(self.shape.dec_layers):
            w.dec_w[i] = g('decoder.mlc.layers.%d.weight' % i)
            w.dec_b[i] = g('decoder.mlc.layers.%d.bias' % i)
        w.dec_out_w, w.dec_out_b = g('decoder.conv.weight'), g('decoder.conv.bias')
        for i in range(self.shape.ref_layers):
            w.ref_w[i] = g('refine.mlc.layers.%d.weight' % i)
            w.ref_b[i] = g('refine.mlc.layers.%d.bias' % i)
        w.mlp_w, w.mlp_b = g('refine.mlp.layers.0.weight'), g('refine.mlp.layers.0.bias')
        w.lstm_w_ih, w.lstm_w_hh = g('refine.lstm.weight_ih'), g('refine.lstm.weight_hh')
        w.lstm_b_ih, w.lstm_b_hh = g('refine.lstm.bias_ih'), g('refine.lstm.bias_hh')
        w.mean_w, w.mean_b = g('refine.mean_update.weight'), g('refine.mean_update.bias')
        w.logvar_w, w.logvar_b = g('refine.logvar_update.weight'), g('refine.logvar_update.bias')
        w.init_mean, w.init_logvar = g('posterior.init_mean'), g('posterior.init_logvar')
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_plan_set_weights(self._plan, C.byref(w), _stream()))
            torch.cuda.current_stream().synchronize()   # sources may be freed afterwards
        self._keep = None

    # ------------------------------------------------------------------ training
    GRAD_KEYS = None

    def _grad_struct(self, grads):
        """IodineGrads over ``grads``: state_dict key -> fp32 CUDA tensor of the parameter's (engine) shape"""
        g = _cabi.IodineGrads()
        p = lambda k: grads[k].data_ptr()
        for i in range(self.shape.dec_layers):
            g.dec_w[i], g.dec_b[i] = p('decoder.mlc.layers.%d.weight' % i), p('decoder.mlc.layers.%d.bias' % i)
        g.dec_out_w, g.dec_out_b = p('decoder.conv.weight'), p('decoder.conv.bias')
        for i in range(self.shape.ref_layers):
            g.ref_w[i], g.ref_b[i] = p('refine.mlc.layers.%d.weight' % i), p('refine.mlc.layers.%d.bias' % i)
        g.mlp_w, g.mlp_b = p('refine.mlp.layers.0.weight'), p('refine.mlp.layers.0.bias')
        g.lstm_w_ih, g.lstm_w_hh = p('refine.lstm.weight_ih'), p('refine.lstm.weight_hh')
        g.lstm_b_ih, g.lstm_b_hh = p('refine.lstm.bias_ih'), p('refine.lstm.bias_hh')
        g.mean_w, g.mean_b = p('refine.mean_update.weight'), p('refine.mean_update.bias')
        g.logvar_w, g.logvar_b = p('refine.logvar_update.weight'), p('refine.logvar_update.bias')
        g.init_mean, g.init_logvar = p('posterior.init_mean'), p('posterior.init_logvar')
        return g

    def train_step(self, x, eps, grads, global_batch=0):
        """IODINE.forward + loss.backward() (reference iodine.py:115-158, lib/engine/train.py:60-65) in one library
        call.  ``grads``: state_dict key -> preallocated fp32 CUDA tensor (engine shapes: refine layer 0 with all 17
        input channels), overwritten with dLoss/dParameter.  Returns (loss 0-dim, elbo_terms [T+1,2])."""
        B, K, L, T = self.B, self.K, self.L, self.T
        x = self._f32(x, (B, 3, self.H, self.W))
        eps = self._f32(eps, (T + 1, B, K, L))
        with torch.cuda.device(self.device):
            if getattr(self, '_train_ws', None) is None:
                need = C.c_size_t()
                _cabi.check(self.lib.iodine_plan_train_workspace_bytes(self._plan, C.byref(need)))
                self._train_ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
                base = (self._train_ws.data_ptr() + 1023) // 1024 * 1024
                _cabi.check(self.lib.iodine_plan_set_train_workspace(self._plan, C.c_void_p(base), need.value))
                self.train_workspace_bytes = need.value
            loss, terms = self._new(1), self._new(T + 1, 2)
            gs = self._grad_struct(grads)
            _cabi.check(self.lib.iodine_train_step(self._plan, _ptr(x), _ptr(eps), int(global_batch), C.byref(gs),
                                                   _ptr(loss), _ptr(terms), _stream()))
        return loss[0], terms

    # ------------------------------------------------------------------ helpers
    def _f32(self, t, shape=None):
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), 'expected %s, got %s' % (shape, tuple(t.shape))
        return t

    def _new(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ entry points
    def init_state(self):
        B, K, L, M = self.B, self.K, self.L, self.M
        mu, lv = self._new(B, K, L), self._new(B, K, L)
        h, c = self._new(B * K, M), self._new(B * K, M)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_init_state(self._plan, _ptr(mu), _ptr(lv), _ptr(h), _ptr(c), _stream()))
        return mu, lv, h, c

    def refine_step(self, x, eps_t, mu, lv, h, c, want_aux=False):
        """In-place on mu, lv, h, c.  Returns (terms[2], aux or None, latent or None)."""
        B, K, L = self.B, self.K, self.L
        x = self._f32(x, (B, 3, self.H, self.W))
        eps_t = self._f32(eps_t, (B, K, L))
        terms = self._new(2)
        aux = None
        if want_aux:
            aux = self._new(B * K * 17 * self.H * self.W + B * K * 4 * L)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_refine_step(
                self._plan, _ptr(x), _ptr(eps_t), _ptr(mu), _ptr(lv), _ptr(h), _ptr(c),
                _ptr(terms), _ptr(aux), _stream()))
        if want_aux:
            n = B * K * 17 * self.H * self.W
            return terms, aux[:n].view(B, K, 17, self.H, self.W), aux[n:].view(B, K, 4 * L)
        return terms, None, None

    def elbo_terms(self, x, eps_t, mu, lv):
        x = self._f32(x, (self.B, 3, self.H, self.W))
        terms = self._new(2)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_elbo(self._plan, _ptr(x), _ptr(self._f32(eps_t)),
                                             _ptr(self._f32(mu)), _ptr(self._f32(lv)), _ptr(terms), _stream()))
        return terms

    def _stage_inputs(self, x, eps):
        B, K, L, T = self.B, self.K, self.L, self.T
        assert tuple(x.shape) == (B, 3, self.H, self.W), 'expected %s, got %s' % ((B, 3, self.H, self.W), tuple(x.shape))
        assert tuple(eps.shape) == (T + 1, B, K, L), 'eps must be [T+1,B,K,L]'
        self._x_in.copy_(x)
        self._eps_in.copy_(eps)

    def encode(self, x, eps):
        T = self.T
        with torch.cuda.device(self.device):
            self._stage_inputs(x, eps)
            _cabi.check(self.lib.iodine_encode(self._plan, _ptr(self._x_in), _ptr(self._eps_in), _ptr(self._z),
                                               _ptr(self._terms), _ptr(self._post), _stream()))
            return self._z.clone(), self._terms[:T].clone(), self._post.clone()

    def decode(self, z):
        B, K, Kt = self.B, self.K, self.K_total
        z = self._f32(z, (B, K, self.L))
        pred, mask, mean = (self._new(B, 3, self.H, self.W), self._new(B, Kt, 1, self.H, self.W),
                            self._new(B, Kt, 3, self.H, self.W))
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_decode(self._plan, _ptr(z), _ptr(pred), _ptr(mask), _ptr(mean), _stream()))
        return pred, mask, mean

    def reconstruct(self, x, eps):
        B, K, T = self.B, self.K_total, self.T
        pred, mask, mean = (self._new(B, 3, self.H, self.W), self._new(B, K, 1, self.H, self.W),
                            self._new(B, K, 3, self.H, self.W))
        with torch.cuda.device(self.device):
            self._stage_inputs(x, eps)
            _cabi.check(self.lib.iodine_reconstruct(self._plan, _ptr(self._x_in), _ptr(self._eps_in), _ptr(pred),
                                                    _ptr(mask), _ptr(mean), _ptr(self._z), _ptr(self._terms), _stream()))
            return pred, mask, mean, self._z.clone(), self._terms[:T].clone()

    def reconstruct_host(self, x_host, eps_host, out=None, sync=True):
        """HOST (ideally pinned) buffers in and out; synchronises unless ``sync=False`` (then the outputs are valid
        after the caller synchronises the current stream -- used to double-buffer two engines on two streams).  ``out`` may carry
        pre-allocated pinned tensors 'pred','mask','mean','z','terms'."""
        B, K, L, T = self.B, self.K, self.L, self.T
        assert x_host.device.type == 'cpu' and eps_host.device.type == 'cpu'
        x_host = x_host.to(torch.float32).contiguous()
        eps_host = eps_host.to(torch.float32).contiguous()
        if out is None:
            pin = torch.cuda.is_available()
            mk = lambda *s: torch.empty(*s, dtype=torch.float32, pin_memory=pin)
            Kt = self.K_total
            out = dict(pred=mk(B, 3, self.H, self.W), mask=mk(B, Kt, 1, self.H, self.W),
                       mean=mk(B, Kt, 3, self.H, self.W), z=mk(B, K, L), terms=mk(max(T, 1), 2))
        with torch.cuda.device(self.device):
            fn = self.lib.iodine_reconstruct_host if sync else self.lib.iodine_reconstruct_host_async
            _cabi.check(fn(
                self._plan, _ptr(x_host), _ptr(eps_host), _ptr(out['pred']), _ptr(out['mask']),
                _ptr(out['mean']), _ptr(out['z']), _ptr(out['terms']), _stream()))
        return out

    def evaluate_host(self, x_host, eps_host, out=None, sync=True):
        """The evaluation flow with HOST (ideally pinned) buffers: like ``reconstruct_host`` but the masks come back
        as their per-pixel argmax (uint8 [B,H,W]) -- what lib/eval/ari_eval.py:32-39 keeps of them.  ``out`` may carry
        pre-allocated pinned tensors 'pred','argmax','z','terms'."""
        B, K, L, T = self.B, self.K, self.L, self.T
        assert x_host.device.type == 'cpu' and eps_host.device.type == 'cpu'
        x_host = x_host.to(torch.float32).contiguous()
        eps_host = eps_host.to(torch.float32).contiguous()
        if out is None:
            pin = torch.cuda.is_available()
            out = dict(pred=torch.empty(B, 3, self.H, self.W, pin_memory=pin),
                       argmax=torch.empty(B, self.H, self.W, dtype=torch.uint8, pin_memory=pin),
                       z=torch.empty(B, K, L, pin_memory=pin), terms=torch.empty(max(T, 1), 2, pin_memory=pin))
        with torch.cuda.device(self.device):
            fn = self.lib.iodine_evaluate_host if sync else self.lib.iodine_evaluate_host_async
            _cabi.check(fn(self._plan, _ptr(x_host), _ptr(eps_host), _ptr(out.get('pred')), _ptr(out.get('argmax')),
                           _ptr(out.get('z')), _ptr(out.get('terms')), _stream()))
        return out

    def last_elbo_image0(self):
        """(pred[3,H,W], mask[K,H,W], mean[K,3,H,W]) of image 0 as the LAST elbo() evaluation of this plan produced
        them -- what the reference hands its logger (iodine.py:225-239)."""
        K, H, W = self.K_total, self.H, self.W
        pred, mask, mean = self._new(3, H, W), self._new(K, H, W), self._new(K, 3, H, W)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_plan_last_elbo_image0(self._plan, _ptr(pred), _ptr(mask), _ptr(mean), _stream()))
        return pred, mask, mean

    def debug_read(self, name, dtype=torch.float32):
        need = C.c_size_t()
        _cabi.check(self.lib.iodine_debug_read(self._plan, name.encode(), None, 0, C.byref(need), _stream()))
        buf = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.iodine_debug_read(self._plan, name.encode(), _ptr(buf), need.value,
                                                   C.byref(need), _stream()))
        return buf.view(dtype)

    def set_comm(self, comm, rank=0, nranks=1):
        """Install (or, with ``comm=None``, remove) an NCCL communicator: afterwards encode()/reconstruct()
        return the per-step ``[T,2]`` sums all-reduced over its ranks (``iodine_plan_set_comm``).
        ``comm`` is an ``iodine_b200.parallel.NcclComm`` or a raw ``ncclComm_t`` address."""
        handle = getattr(comm, 'handle', comm)
        self._comm_keep = comm                      # the communicator must outlive the plan's use of it
        _cabi.check(self.lib.iodine_plan_set_comm(self._plan, C.c_void_p(handle or None), int(rank), int(nranks)))

    def launch_count(self):
        n = C.c_uint64()
        _cabi.check(self.lib.iodine_plan_launch_count(self._plan, C.byref(n)))
        return n.value

    def profile(self, enable):
        _cabi.check(self.lib.iodine_plan_profile(self._plan, int(enable)))

    def profile_read(self):
        """(summed device ms of the bracketed decoder-conv launches, number of launches)"""
        ms, n = C.c_double(), C.c_uint64()
        _cabi.check(self.lib.iodine_plan_profile_read(self._plan, C.byref(ms), C.byref(n)))
        return ms.value, n.value
