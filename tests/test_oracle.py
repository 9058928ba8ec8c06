"""CPU: the restatement (oracle/restatement.py) against (a) the committed golden vectors the
unmodified reference produced and (b) the live reference when /root/reference is mounted."""
import pytest
import torch

from oracle import arch as A
from oracle import make_golden as MG
from oracle import ref_loader as R
from oracle import restatement as S

from helpers import golden_state_dict, rel_err, t

FULL = [n for n, c in MG.CASES.items() if c[4] == 'full']
SMALL_OUT = ['dsprites_b2', 'dsprites_b2_sharp']
TOL = 2e-5   # fp32 restatement vs fp32 autograd reference (different summation orders)


@pytest.mark.parametrize('name', FULL + SMALL_OUT)
def test_restatement_matches_golden(name):
    g, arch, B, sd, _ = golden_state_dict(name)
    x, eps = t(g['x']), t(g['eps'])
    tr = S.encode_trace(sd, arch, x, eps, want_aux=True)
    keys = MG.STEP_KEYS_FULL if name in FULL else MG.STEP_KEYS_OUT
    for i, st in enumerate(tr['steps']):
        for k in keys:
            e = rel_err(st[k], g['s%d_%s' % (i, k)])
            assert e < TOL, (name, i, k, e)
    for k in ('post_mean', 'post_logvar', 'z', 'pred', 'mask', 'mean'):
        e = rel_err(tr[k], g['final_' + k])
        assert e < TOL, (name, k, e)


def test_restatement_matches_golden_clevr6():
    g, arch, B, sd, _ = golden_state_dict('clevr6_b1')
    tr = S.encode_trace(sd, arch, t(g['x']), t(g['eps']))
    for i, st in enumerate(tr['steps']):
        for k in ('elbo', 'post_mean', 'post_logvar'):
            assert rel_err(st[k], g['s%d_%s' % (i, k)]) < TOL, (i, k)
    for k in ('pred', 'mask', 'z'):
        assert rel_err(tr[k], g['final_' + k]) < TOL, k


@pytest.mark.skipif(not R.reference_available(), reason='reference tree not mounted')
@pytest.mark.parametrize('arch_name,over,B,sharpen', [
    ('tiny', {}, 3, 1.0), ('tiny', {}, 1, 7.0), ('test5x5', {}, 2, 3.0),
    ('tiny', dict(layernorm=False), 2, 3.0),                       # ARCH.LAYERNORM off (iodine.py:264, 300-330)
    ('tiny', dict(slots=11, iters=7), 1, 4.0),                     # BASELINE config #4's K / T
    ('tiny', dict(slots=1), 2, 2.0),                               # a single slot: softmax / leave-one-out degenerate
    ('tiny', dict(img_size=24, ref_layers=3, dec_layers=3), 2, 2.0),
    ('dsprites', dict(img_size=32), 1, 3.0),                       # config #1's layer counts and widths, smaller image
])
def test_restatement_matches_live_reference(arch_name, over, B, sharpen):
    arch = A.arch_by_name(arch_name, **over)
    model = R.build_reference_model(arch, sharpen=sharpen)
    x, eps = R.make_inputs(arch, B, seed_x=7, seed_eps=11)
    ref = R.run_reference_trace(model, x, eps)
    sd = S.state_dict_to(model.state_dict(), torch.float32)
    tr = S.encode_trace(sd, arch, x, eps, want_aux=True)
    for i, (a, b) in enumerate(zip(tr['steps'], ref['steps'])):
        for k in b:
            assert rel_err(a[k], b[k]) < TOL, (i, k)
    pred, mask, mean = R.run_reference_reconstruct(model, x, eps)
    assert rel_err(tr['pred'], pred) < TOL and rel_err(tr['mask'], mask) < TOL


def test_closed_form_gradients_match_autograd():
    """Row A4 of SURVEY.md: the five gradients in closed form vs torch.autograd on the same
    restated ELBO (fp64), independent of the reference tree."""
    arch = A.arch_by_name('tiny')
    g, _, B, sd, _ = golden_state_dict('tiny_b2_sharp')
    sd = S.state_dict_to(sd, torch.float64)
    x, eps = t(g['x']).double(), t(g['eps']).double()
    mu = t(g['s1_post_mean']).double().requires_grad_(True)
    lv = t(g['s1_post_logvar']).double().requires_grad_(True)
    z = mu + torch.exp(0.5 * lv) * eps[1]
    mean, logits, acts, _ = S.decoder_forward(sd, z, arch.IMG_SIZE)
    mean.retain_grad()
    mx = S.mixture(x, mean, logits, arch.SIGMA)
    J = mx['ll_sum'] - S.kl_elementwise(mu, lv).sum()
    J.backward()
    dz = S.decoder_dgrad(sd, [a.detach() for a in acts], mx['seed4'].detach(), arch.DIM_LATENT)
    dz = dz.reshape(B, arch.SLOTS, -1)
    mu_g = dz - mu.detach()
    lv_g = dz * 0.5 * torch.exp(0.5 * lv.detach()) * eps[1] - 0.5 * (torch.exp(lv.detach()) - 1)
    assert rel_err(mu_g, mu.grad) < 1e-10
    assert rel_err(lv_g, lv.grad) < 1e-10
    assert rel_err(mx['mean_grad'].detach(), mean.grad) < 1e-10
