"""Generate ``tests/golden/train_*.npz``: loss and every parameter gradient of one training step
(``loss = model(x); loss.backward()``, ``lib/engine/train.py:60-64``) of the UNMODIFIED reference, fp32 autograd on
CPU, with the noise recipe of ``oracle/make_golden.py``.  Build container only:  python -m oracle.make_train_golden
TEST INFRASTRUCTURE -- these are the vectors the weight-gradient kernels will be checked
against on the GPU box, where the reference tree does not exist.
"""
import os

import numpy as np

from . import arch as A
from . import make_golden as MG
from . import ref_loader as R

# name -> (arch name, overrides, B, sharpen)
CASES = {
    'train_tiny_b2_sharp': ('tiny', {}, 2, 4.0),
    'train_test5x5_b2_sharp': ('test5x5', {}, 2, 3.0),
}


def make_case(name):
    arch_name, over, B, sharpen = CASES[name]
    arch = A.arch_by_name(arch_name, **over)
    model = R.build_reference_model(arch, seed=0, sharpen=sharpen)
    x, eps = R.make_inputs(arch, B)
    loss, grads = R.run_reference_training_step(model, x, eps)
    out = {'x': x.numpy(), 'eps': eps.numpy(), 'loss': np.float64(loss.item()),
           'weights_checksum': np.float64(MG.weights_checksum(model.state_dict())),
           'sharpen': np.float64(sharpen), 'B': np.int64(B)}
    for k, g in grads.items():
        out['grad/' + k] = g.numpy()
    path = os.path.join(MG.OUT, name + '.npz')
    np.savez_compressed(path, **out)
    return path


if __name__ == '__main__':
    for n in CASES:
        print(make_case(n))
