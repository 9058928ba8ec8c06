"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Recipe: weights = default PyTorch init under torch.manual_seed(0) (optionally
"sharpened": decoder.conv.weight *= s, init_logvar -= 1, init_mean += 0.25, see
oracle/ref_loader.build_reference_model); x = U[0,1) from generator seed 1; eps = N(0,1)
[T+1,B,K,L] from generator seed 123 injected into torch.randn_like.  Weights are not stored:
tests rebuild them from the seed and compare the stored checksum.
"""
import os
import sys

import numpy as np

from . import arch as A
from . import ref_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# name -> (arch name, arch overrides, B, sharpen, detail level)
CASES = {
    'tiny_b2': ('tiny', {}, 2, 1.0, 'full'),
    'tiny_b2_sharp': ('tiny', {}, 2, 6.0, 'full'),
    'test5x5_b2_sharp': ('test5x5', {}, 2, 5.0, 'full'),
    'dsprites_b2': ('dsprites', {}, 2, 1.0, 'outputs'),          # BASELINE config #1 arch, T=3
    'dsprites_b2_sharp': ('dsprites', {}, 2, 4.0, 'outputs'),
    'clevr6_b1': ('clevr6', {}, 1, 1.0, 'outputs'),              # BASELINE config #2 arch, T=5
    'clevr6_b1_sharp': ('clevr6', {}, 1, 4.0, 'outputs'),
}
STEP_KEYS_FULL = ['elbo', 'kl', 'll', 'z', 'mean', 'mask_logits', 'mask', 'mean_grad', 'mask_grad',
                  'post_mean', 'post_logvar', 'post_mean_grad', 'post_logvar_grad', 'latent', 'aux',
                  'mean_delta', 'logvar_delta', 'lstm_h', 'lstm_c']
STEP_KEYS_OUT = ['elbo', 'kl', 'll', 'post_mean', 'post_logvar', 'post_mean_grad',
                 'post_logvar_grad', 'mean_delta', 'logvar_delta']


def weights_checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def make_case(name):
    arch_name, over, B, sharpen, detail = CASES[name]
    arch = A.arch_by_name(arch_name, **over)
    model = R.build_reference_model(arch, seed=0, sharpen=sharpen)
    x, eps = R.make_inputs(arch, B)
    tr = R.run_reference_trace(model, x, eps, keep_aux=(detail == 'full'))
    out = {'x': x.numpy(), 'eps': eps.numpy(),
           'weights_checksum': np.float64(weights_checksum(model.state_dict())),
           'sharpen': np.float64(sharpen), 'B': np.int64(B)}
    keys = STEP_KEYS_FULL if detail == 'full' else STEP_KEYS_OUT
    for t, st in enumerate(tr['steps']):
        for k in keys:
            out['s%d_%s' % (t, k)] = st[k].numpy()
    for k in ('post_mean', 'post_logvar', 'z', 'pred', 'mask', 'mean'):
        out['final_' + k] = tr[k].numpy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    return path


if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    for n in names:
        p = make_case(n)
        print('%s  %.1f KB' % (p, os.path.getsize(p) / 1024))
