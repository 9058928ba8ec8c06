"""Shared test helpers: seeded weights (no reference tree needed), golden loading, metrics."""
import os

import numpy as np
import torch

from oracle import arch as A
from oracle import make_golden as MG
from oracle import restatement as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def seeded_model(arch, sharpen=1.0, precision='fp32'):
    """iodine_b200.IODINE with the default init of torch.manual_seed(0) -- bit-identical to the
    reference's (tests/test_host.py checks that when /root/reference is present) -- and the
    same 'sharpen' edit as oracle.ref_loader.build_reference_model."""
    from iodine_b200.modeling.iodine import IODINE
    torch.manual_seed(0)
    m = IODINE(arch, precision=precision)
    if sharpen != 1.0:
        with torch.no_grad():
            m.decoder.conv.weight.mul_(sharpen)
            m.posterior.init_logvar.add_(-1.0)
            m.posterior.init_mean.add_(0.25)
    return m


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    arch_name, over, B, sharpen, detail = MG.CASES[name]
    arch = A.arch_by_name(arch_name, **over)
    return g, arch, B, sharpen, detail


def golden_state_dict(name):
    g, arch, B, sharpen, detail = load_golden(name)
    m = seeded_model(arch, sharpen)
    sd = S.state_dict_to(m.state_dict(), torch.float32)
    cs = MG.weights_checksum(sd)
    assert abs(cs - float(g['weights_checksum'])) <= 1e-9 * abs(cs), \
        'seeded weights differ from the ones the golden vectors were made with'
    return g, arch, B, sd, m


def rel_err(a, b):
    """max|a-b| / max|b| (b = the trusted side)."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def t(x):
    return torch.from_numpy(np.asarray(x))


def load_golden_size(name, precision='fp32'):
    """Compact at-size golden (oracle/make_golden_size.py): inputs and weights are regenerated from their seeds and
    pinned by the stored checksums.  Returns (golden dict, arch, B, x, eps, model)."""
    from oracle import make_golden_size as MS
    c = MS.CASES[name]
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    arch, x, eps = MS.case_inputs(name)
    assert abs(x.double().sum().item() - float(g['x_checksum'])) <= 1e-9 * abs(float(g['x_checksum']))
    assert abs(eps.double().abs().sum().item() - float(g['eps_checksum'])) <= 1e-9 * abs(float(g['eps_checksum']))
    m = seeded_model(arch, c['sharpen'], precision=precision)
    cs = MG.weights_checksum(m.state_dict())
    assert abs(cs - float(g['weights_checksum'])) <= 1e-9 * abs(cs), 'seeded weights differ from the golden ones'
    return g, arch, c['B'], x, eps, m


def compact_errors(g, arch, B, pred, mask, mean, z, elbo_steps):
    """Errors of a reconstruct() result against a compact golden: sampled values (max |diff| / max |ref|), exact
    per-image sums (relative), per-step ELBO (relative), z (relative L2) and the argmax of the masks on every pixel
    whose reference top-2 margin exceeds 4e-3 (mismatch count)."""
    from oracle import make_golden_size as MS
    pred, mask, mean, z = (torch.as_tensor(v).detach().cpu() for v in (pred, mask, mean, z))
    e = {}
    for nm, v in (('pred', pred), ('mask', mask), ('mean', mean)):
        flat = v.reshape(-1)
        e[nm] = rel_err(flat[MS.sample_index(flat.numel())], g['final_%s_s' % nm])
    e['pred_sum'] = rel_err(pred.double().sum(dim=(2, 3)), g['final_pred_sum'])
    e['mask_sum'] = rel_err(mask.double().sum(dim=(2, 3, 4)), g['final_mask_sum'])
    e['mean_sum'] = rel_err(mean.double().sum(dim=(3, 4)), g['final_mean_sum'])
    want = torch.tensor([float(g['s%d_elbo' % i]) for i in range(arch.ITERS)], dtype=torch.float64)
    got = torch.as_tensor(elbo_steps, dtype=torch.float64).cpu()
    e['elbo'] = ((got - want).abs() / want.abs()).max().item()
    zr = t(g['final_z']).double()
    e['z_l2'] = ((z.double() - zr).norm() / zr.norm()).item()
    am = mask[:, :, 0].argmax(dim=1).to(torch.uint8)
    clear = t(g['final_mask_margin']).float() > 4e-3
    e['argmax_mismatch'] = int(((am != t(g['final_mask_argmax'])) & clear).sum().item())
    e['argmax_checked'] = int(clear.sum().item())
    return e
