"""Run the UNMODIFIED reference (``/root/reference/lib/modeling/iodine.py``) on CPU with
injected noise and record every tensor the hot path produces.  TEST INFRASTRUCTURE.

Only usable where ``/root/reference`` is mounted (the build container).  The GPU box has
no reference tree: tests there use the committed fixtures in ``tests/golden/`` that
``oracle/make_golden.py`` produced with this module.
"""
import contextlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get('IODINE_REFERENCE_ROOT', '/root/reference')
# the GPU box has no /root/reference: oracle/build_ref.py leaves a byte-for-byte copy of the path's modules in
# oracle/_ref (git-ignored build output that travels with the snapshot)
_SHIPPED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
if not os.path.isfile(os.path.join(REFERENCE_ROOT, 'lib', 'modeling', 'iodine.py')) and \
        os.path.isfile(os.path.join(_SHIPPED, 'lib', 'modeling', 'iodine.py')):
    REFERENCE_ROOT = _SHIPPED


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'lib', 'modeling', 'iodine.py'))


def import_reference():
    """Import ``lib.modeling.iodine`` from the reference tree, untouched."""
    if not reference_available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import lib.modeling.iodine as ref_iodine  # noqa: E402
    return ref_iodine


@contextlib.contextmanager
def injected_noise(eps):
    """Replace ``torch.randn_like`` (the reference's only RNG draw, iodine.py:632) by a
    reader of the pre-drawn ``eps[T+1, B, K, L]`` so CPU and CUDA runs see the same noise."""
    state = {'i': 0}
    orig = torch.randn_like

    def fake(t, **kw):
        e = eps[state['i']].to(dtype=t.dtype).reshape(t.shape)
        state['i'] += 1
        return e

    torch.randn_like = fake
    try:
        yield state
    finally:
        torch.randn_like = orig


def build_reference_model(arch, seed=0, sharpen=1.0, dtype=torch.float32):
    """Default PyTorch init under ``torch.manual_seed(seed)`` (the reference never
    re-initialises, iodine.py:54-57).  ``sharpen`` scales the decoder's last conv so masks
    are far from uniform."""
    ref = import_reference()
    torch.manual_seed(seed)
    model = ref.IODINE(arch).to(dtype)
    if sharpen != 1.0:
        with torch.no_grad():
            model.decoder.conv.weight.mul_(sharpen)
            model.posterior.init_logvar.add_(-1.0)
            model.posterior.init_mean.add_(0.25)
    return model


def make_inputs(arch, B, seed_x=1, seed_eps=123, dtype=torch.float32):
    """x ~ U[0,1) and eps ~ N(0,1) exactly as SURVEY.md 8(c)/(d) prescribe."""
    g = torch.Generator().manual_seed(seed_x)
    x = torch.rand(B, 3, arch.IMG_SIZE, arch.IMG_SIZE, generator=g, dtype=torch.float32)
    ge = torch.Generator().manual_seed(seed_eps)
    eps = torch.randn(arch.ITERS + 1, B, arch.SLOTS, arch.DIM_LATENT, generator=ge,
                      dtype=torch.float32)
    return x.to(dtype), eps.to(dtype)


def run_reference_trace(model, x, eps, keep_aux=True):
    """Replays ``IODINE.encode`` (iodine.py:73-105) statement by statement -- calling the
    reference's own ``elbo`` / ``get_input_encoding`` / ``refine`` / ``update`` -- and
    records the intermediates; then ``decode`` (59-71).  Returns a dict of CPU tensors."""
    B = x.shape[0]
    tr = {'steps': []}
    with injected_noise(eps):
        model.posterior.init_unit(B, model.K)
        model.lstm_hidden = None
        for i in range(model.n_iters):
            elbo = model.elbo(x)
            (B * elbo).backward(retain_graph=False)
            inp, latent = model.get_input_encoding(x)
            st = {
                'elbo': elbo.detach().clone(),
                'kl': model.posterior.kl_divergence().detach().mean(0).sum(),
                'll': model.log_likelihood.detach().mean(0).sum(),
                'z': model.z.detach().clone(),
                'mean': model.mean.detach().clone(),
                'mask_logits': model.mask_logits.detach().clone(),
                'mask': model.mask.detach().clone(),
                'mean_grad': model.mean.grad.detach().clone(),
                'mask_grad': model.mask.grad.detach().clone(),
                'post_mean': model.posterior.mean.detach().clone(),
                'post_logvar': model.posterior.logvar.detach().clone(),
                'post_mean_grad': model.posterior.mean.grad.detach().clone(),
                'post_logvar_grad': model.posterior.logvar.grad.detach().clone(),
                'latent': latent.clone(),
            }
            if keep_aux:
                st['aux'] = inp.clone()
            md, ld, model.lstm_hidden = model.refine(inp, latent, model.lstm_hidden)
            md, ld = md.detach(), ld.detach()
            st['mean_delta'] = md.clone()
            st['logvar_delta'] = ld.clone()
            # NOTE iodine.py:488 unpacks LSTMCell's (h, c) as (c, h); we store by true role
            st['lstm_h'] = model.lstm_hidden[0].detach().clone()
            st['lstm_c'] = model.lstm_hidden[1].detach().clone()
            model.posterior.update(md, ld)
            tr['steps'].append(st)
        tr['post_mean'] = model.posterior.mean.detach().clone()
        tr['post_logvar'] = model.posterior.logvar.detach().clone()
        z = model.posterior.sample()
        tr['z'] = z.detach().clone()
        pred, mask, mean = model.decode(z)
        tr['pred'] = pred.detach().clone()
        tr['mask'] = mask.detach().clone()
        tr['mean'] = mean.detach().clone()
    model.zero_grad(set_to_none=True)
    return tr


def run_reference_reconstruct(model, x, eps):
    """``model.reconstruct(x)`` (iodine.py:107-112) through the reference's public API."""
    with injected_noise(eps):
        pred, mask, mean = model.reconstruct(x)
    model.zero_grad(set_to_none=True)
    return pred.detach(), mask.detach(), mean.detach()


def run_reference_training_step(model, x, eps):
    """``loss = model(x)`` (IODINE.forward, iodine.py:115-158) followed by what ``lib/engine/train.py:60-64`` does
    with it: ``loss.mean()``, ``optimizer.zero_grad()`` -- which discards the parameter gradients the in-loop
    ``(B * elbo).backward(retain_graph=True)`` calls have accumulated -- and ``loss.backward()``.
    Returns (loss value, {state_dict key: gradient})."""
    with injected_noise(eps):
        loss = model(x)
    loss = loss.mean()
    model.zero_grad(set_to_none=True)
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    return loss.detach().clone(), grads
