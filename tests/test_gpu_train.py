"""GPU: the training step (reference lib/modeling/iodine.py:115-158 + lib/engine/train.py:60-65).

``loss = model(x); loss.backward()`` on the native module must fill ``.grad`` of EVERY state_dict key with the
reference's autograd gradient: checked against the committed fixtures of the unmodified reference
(tests/golden/train_*.npz, oracle/make_train_golden.py) and against the explicit restatement
(oracle/train_restatement.py, itself pinned to the reference's autograd in tests/test_train_oracle.py) on further
shapes.  Bars: fp32 kernels 2e-3 (the fixtures are the reference's own fp32 autograd), tensor-core modes 2e-2 relative
to the largest entry of each gradient tensor (10-bit operand mantissas through T+1 decoder passes)."""
import os

import numpy as np
import pytest
import torch

from oracle import arch as A
from oracle import make_train_golden as MTG
from oracle import restatement as S
from oracle import train_restatement as TR

from helpers import GOLDEN, seeded_model, t

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BAR = {'fp32': 2e-3, 'tf32': 2e-2, 'fp16': 2e-2}


def _rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


def _step(model, x, eps):
    model.to(DEV).zero_grad(set_to_none=True)
    loss = model(x.to(DEV), eps=eps.to(DEV))
    loss = loss.mean()                                # lib/engine/train.py:61
    loss.backward()
    torch.cuda.synchronize()
    return loss, {k: p.grad for k, p in model.named_parameters()}


@pytest.mark.parametrize('prec', ['fp32', 'tf32', 'fp16'])
@pytest.mark.parametrize('name', ['train_tiny_b2_sharp', 'train_test5x5_b2_sharp'])
def test_backward_fills_every_gradient_like_the_reference(name, prec):
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    arch_name, over, B, sharpen = MTG.CASES[name]
    arch = A.arch_by_name(arch_name, **over)
    model = seeded_model(arch, sharpen, precision=prec)
    loss, grads = _step(model, t(g['x']), t(g['eps']))
    bar = BAR[prec]
    assert abs(loss.item() - float(g['loss'])) <= (1e-4 if prec == 'fp32' else 2e-3) * abs(float(g['loss'])), (loss.item(), float(g['loss']))
    errs = {}
    for k in model.state_dict():
        assert grads[k] is not None, k
        assert tuple(grads[k].shape) == tuple(g['grad/' + k].shape), k
        errs[k] = _rel(grads[k], g['grad/' + k])
    print(prec, name, {k: '%.1e' % v for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < bar}
    assert not bad, bad


@pytest.mark.parametrize('over,B,sharpen', [
    (dict(iters=1), 2, 2.0),                          # a single refiner call: no LSTM chain
    (dict(img_size=20, ref_layers=3), 2, 2.0),        # odd stride-2 sizes (conv_transpose output_padding)
    (dict(dec_layers=1), 3, 2.0),                     # decoder = collapsed first layer + decoder.conv only
    (dict(layernorm=False), 2, 3.0),
])
def test_training_step_against_the_restatement_on_other_shapes(over, B, sharpen):
    arch = A.arch_by_name('tiny', **over)
    model = seeded_model(arch, sharpen, precision='fp32')
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(B, 3, arch.IMG_SIZE, arch.IMG_SIZE, generator=gen)
    eps = torch.randn(arch.ITERS + 1, B, arch.SLOTS, arch.DIM_LATENT, generator=gen)
    sd64 = S.state_dict_to(model.state_dict(), torch.float64)
    want_loss, want, _ = TR.loss_and_grads(sd64, arch, x.double(), eps.double())
    loss, grads = _step(model, x, eps)
    assert abs(loss.item() - want_loss.item()) <= 1e-4 * abs(want_loss.item())
    errs = {k: _rel(grads[k], want[k]) for k in want}
    bad = {k: v for k, v in errs.items() if not v < 2e-3}
    assert not bad, bad


def test_optimizer_step_changes_the_engine_weights_and_lowers_the_loss():
    """lib/engine/train.py:60-65 verbatim: a few Adam steps on one batch must reduce the loss, i.e. the gradients
    point the right way and the engine picks the updated parameters up"""
    arch = A.arch_by_name('tiny')
    model = seeded_model(arch, 2.0, precision='fp32').to(DEV)
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(4, 3, arch.IMG_SIZE, arch.IMG_SIZE, generator=gen).to(DEV)
    eps = torch.randn(arch.ITERS + 1, 4, arch.SLOTS, arch.DIM_LATENT, generator=gen).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    losses = []
    for _ in range(8):
        loss = model(x, eps=eps)
        loss = loss.mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


def test_no_grad_forward_still_returns_the_value():
    arch = A.arch_by_name('tiny')
    model = seeded_model(arch, 2.0).to(DEV)
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, arch.IMG_SIZE, arch.IMG_SIZE, generator=gen).to(DEV)
    eps = torch.randn(arch.ITERS + 1, 2, arch.SLOTS, arch.DIM_LATENT, generator=gen).to(DEV)
    with torch.no_grad():
        v = model(x, eps=eps)
    loss = model(x, eps=eps)
    assert not v.requires_grad and loss.requires_grad
    assert abs(v.item() - loss.item()) <= 1e-5 * abs(v.item())


# --------------------------------------------------------------------------- CLEVR6 layer sizes (128x128, C = 64)
@pytest.mark.parametrize('prec', ['fp32', 'tf32', 'fp16', 'bf16'])
def test_training_step_at_clevr6_size_against_reference_golden(prec):
    """the shapes bench.py --mode train times; 16-bit modes run the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cu)
    for the C -> C layers.  Fixture: sampled entries + sums of the unmodified reference's gradients
    (oracle/make_train_golden_size.py)."""
    from oracle import make_golden as MG
    from oracle import make_golden_size as MS
    from oracle import make_train_golden_size as TS
    g = dict(np.load(os.path.join(GOLDEN, TS.NAME + '.npz')))
    arch, x, eps = TS.case_inputs()
    assert abs(x.double().sum().item() - float(g['x_checksum'])) <= 1e-9 * abs(float(g['x_checksum']))
    model = seeded_model(arch, TS.SHARPEN, precision=prec)
    cs = MG.weights_checksum(model.state_dict())
    assert abs(cs - float(g['weights_checksum'])) <= 1e-9 * abs(cs)
    loss, grads = _step(model, x, eps)
    bar = {'fp32': 2e-3, 'tf32': 2e-2, 'fp16': 2e-2, 'bf16': 6e-2}[prec]
    assert abs(loss.item() - float(g['loss'])) <= (1e-4 if prec == 'fp32' else 3e-3) * abs(float(g['loss']))
    errs = {}
    for k in model.state_dict():
        flat = grads[k].detach().reshape(-1).double().cpu()
        want = torch.from_numpy(g['grad_s/' + k]).double()
        gmax = float(g['grad_max/' + k])
        e_s = ((flat[MS.sample_index(flat.numel(), TS.NS)] - want).abs().max() / gmax).item()
        e_sum = abs(flat.sum().item() - float(g['grad_sum/' + k])) / float(g['grad_abs/' + k])
        errs[k] = max(e_s, e_sum)
    print(prec, {k: '%.1e' % v for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < bar}
    assert not bad, bad
