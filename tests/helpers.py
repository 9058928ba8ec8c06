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
