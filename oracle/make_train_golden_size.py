"""Generate ``tests/golden/train_clevr6_b2_t2.npz``: one training step of the UNMODIFIED reference
(``loss = model(x); loss.mean(); zero_grad(); loss.backward()``, lib/engine/train.py:60-64) at the CLEVR6 layer sizes
(128x128, C = 64 -- the shapes the tensor-core weight-gradient kernel of csrc/wgrad_tc.cu runs), B = 2, K = 7,
T = 2, fp32 CPU.  The full gradient is 4.4 MB, so the fixture keeps, per parameter: 4096 sampled entries (fixed index
set), the exact sum and the sum of absolute values (float64).  Inputs and weights are regenerated from their seeds
and pinned by checksums.  Build container only:  python -m oracle.make_train_golden_size
TEST INFRASTRUCTURE -- nothing under iodine_b200/ imports this."""
import os

import numpy as np
import torch

from . import arch as A
from . import make_golden as MG
from . import make_golden_size as MS
from . import ref_loader as R

NAME = 'train_clevr6_b2_t2'
B, SHARPEN = 2, 3.0
NS = 4096


def case_arch():
    return A.arch_by_name('clevr6', iters=2)


def case_inputs():
    arch = case_arch()
    x, eps = R.make_inputs(arch, B)
    # block-structured image content on top of the noise, so that masks and gradients are not uniform
    yy, xx = torch.meshgrid(torch.arange(arch.IMG_SIZE), torch.arange(arch.IMG_SIZE), indexing='ij')
    for b in range(B):
        for c in range(3):
            x[b, c] = 0.5 * x[b, c] + 0.5 * (((yy // (16 + 8 * b) + xx // (24 - 4 * c)) % 3) / 2.0)
    return arch, x, eps


def sample(t):
    flat = t.reshape(-1)
    return flat[MS.sample_index(flat.numel(), NS)]


def make():
    arch, x, eps = case_inputs()
    model = R.build_reference_model(arch, seed=0, sharpen=SHARPEN)
    loss, grads = R.run_reference_training_step(model, x, eps)
    out = {'loss': np.float64(loss.item()), 'x_checksum': np.float64(x.double().sum().item()),
           'eps_checksum': np.float64(eps.double().abs().sum().item()),
           'weights_checksum': np.float64(MG.weights_checksum(model.state_dict())), 'sharpen': np.float64(SHARPEN)}
    for k, g in grads.items():
        out['grad_s/' + k] = sample(g).numpy()
        out['grad_sum/' + k] = np.float64(g.double().sum().item())
        out['grad_abs/' + k] = np.float64(g.double().abs().sum().item())
        out['grad_max/' + k] = np.float64(g.double().abs().max().item())
    path = os.path.join(MG.OUT, NAME + '.npz')
    np.savez_compressed(path, **out)
    return path


if __name__ == '__main__':
    print(make())
