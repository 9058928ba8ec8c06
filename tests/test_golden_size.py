"""CPU: the oracle restatement against the compact at-size goldens of the unmodified reference
(oracle/make_golden_size.py): K=11/T=7, 256x256/K=16 and the trained-like stress cases.  The B=32 case costs a
minute of CPU and is checked only when IODINE_SLOW_TESTS=1 (the GPU parity tests use it at every run)."""
import os

import pytest
import torch

from oracle import make_golden_size as MS
from oracle import restatement as S

from helpers import compact_errors, load_golden_size

CASES = [n for n in MS.CASES if MS.CASES[n]['B'] == 1 or os.environ.get('IODINE_SLOW_TESTS')]


@pytest.mark.parametrize('name', CASES)
def test_restatement_against_compact_golden(name):
    g, arch, B, x, eps, model = load_golden_size(name)
    assert bool(g['reference_finite'])
    sd = S.state_dict_to(model.state_dict(), torch.float32)
    with torch.no_grad():
        tr = S.encode_trace(sd, arch, x, eps)
    e = compact_errors(g, arch, B, tr['pred'], tr['mask'], tr['mean'], tr['z'], [s['elbo'] for s in tr['steps']])
    print(name, {k: ('%.2e' % v if isinstance(v, float) else v) for k, v in e.items()})
    assert max(e['pred'], e['mask'], e['mean'], e['elbo'], e['pred_sum'], e['mask_sum'], e['mean_sum']) < 1e-4, e
    assert e['z_l2'] < 1e-3 and e['argmax_mismatch'] == 0 and e['argmax_checked'] > 0.9 * B * arch.IMG_SIZE ** 2 * 0.5, e


def test_block_images_are_structured():
    x = MS.block_image(2, 64)
    assert x.shape == (2, 3, 64, 64) and x.min() >= 0.35 and x.max() <= 0.65
    # piecewise constant: most horizontal neighbours are equal
    assert (x[..., 1:] == x[..., :-1]).float().mean() > 0.8
