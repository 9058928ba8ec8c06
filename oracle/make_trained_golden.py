"""A TRAINED-weights fixture from the unmodified reference (build container only).  TEST INFRASTRUCTURE.

    python -m oracle.make_trained_golden [steps]

Every other golden uses default-initialised weights (optionally "sharpened"), whose masks stay close to uniform
and whose gradients are small.  This script TRAINS the reference model itself -- `loss = model(x)`,
`loss.backward()`, Adam, exactly what lib/engine/train.py:60-65 does -- for a few hundred steps on procedurally
generated sprite images (dSprites architecture, configs/dsprites_noclip.yaml, K=4, T=3), then records
  * the trained weights (fp32, every state_dict key)                  -> tests/golden/trained_dsprites_weights.npz
  * a compact golden of `reconstruct()` on held-out images with them -> tests/golden/trained_dsprites_b4.npz
so that the 16-bit kernels are checked in the regime they will serve: peaked likelihoods, means close to the
image, specialised masks, large gradient seeds.
"""
import os
import sys
import time

import numpy as np
import torch

from . import arch as A
from . import make_golden as MG
from . import make_golden_size as MS
from . import ref_loader as R

ARCH = dict(name='dsprites', over=dict(slots=4, iters=3))
B_TRAIN, B_TEST = 8, 4
NAME = 'trained_dsprites_b4'
WEIGHTS = 'trained_dsprites_weights'


def sprite_images(B, S, seed):
    """Uniform dark background with 2-3 axis-aligned rectangles / discs of saturated random colours."""
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(B, 3, S, S)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing='ij')
    for b in range(B):
        x[b] = (0.05 + 0.1 * torch.rand(3, generator=g))[:, None, None]
        for _ in range(int(torch.randint(2, 4, (1,), generator=g))):
            cy, cx = [int(v) for v in torch.randint(S // 6, S - S // 6, (2,), generator=g)]
            r = int(torch.randint(S // 10, S // 5, (1,), generator=g))
            col = 0.3 + 0.7 * torch.rand(3, generator=g)
            if torch.rand(1, generator=g).item() < 0.5:
                m = ((yy - cy).abs() <= r) & ((xx - cx).abs() <= r)
            else:
                m = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
            x[b][:, m] = col[:, None]
    return x


def test_inputs():
    arch = A.arch_by_name(ARCH['name'], **ARCH['over'])
    x = sprite_images(B_TEST, arch.IMG_SIZE, seed=99991)
    eps = torch.randn(arch.ITERS + 1, B_TEST, arch.SLOTS, arch.DIM_LATENT, generator=torch.Generator().manual_seed(123))
    return arch, x, eps


def main(steps):
    arch, x_test, eps_test = test_inputs()
    model = R.build_reference_model(arch, seed=0)
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    torch.manual_seed(5)
    t0 = time.time()
    # The reference's own arithmetic is not NaN-safe once likelihoods are peaked (un-stabilised exp(sum ll) -> 0/0 in
    # mask_posterior, iodine.py:289-292), so a run can diverge: keep a snapshot every 25 steps and use the LATEST one
    # whose reconstruct() of the held-out images is finite.
    snaps = []
    for it in range(steps):
        x = sprite_images(B_TRAIN, arch.IMG_SIZE, seed=1000 + it)
        loss = model(x).mean()
        if not torch.isfinite(loss):
            print('step %d: loss is not finite, stopping' % it, flush=True)
            break
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it % 20 == 0 or it == steps - 1:
            print('step %d loss %.1f (%.0f s)' % (it, loss.item(), time.time() - t0), flush=True)
        if (it + 1) % 25 == 0 or it == steps - 1:
            snaps.append((it + 1, loss.item(), {k: v.detach().clone() for k, v in model.state_dict().items()}))
    model.zero_grad(set_to_none=True)
    tr = None
    for n_steps, last_loss, sd in reversed(snaps):
        model.load_state_dict(sd)
        if not all(torch.isfinite(v).all() for v in sd.values()):
            continue
        cand = R.run_reference_trace(model, x_test, eps_test, keep_aux=False)
        if all(torch.isfinite(cand[k]).all() for k in ('pred', 'mask', 'mean', 'z')) and \
                all(torch.isfinite(st['elbo']) for st in cand['steps']):
            tr, steps, loss = cand, n_steps, torch.tensor(last_loss)
            break
    assert tr is not None, 'no finite snapshot'
    print('using the snapshot after %d steps (loss %.1f)' % (steps, loss.item()), flush=True)
    np.savez_compressed(os.path.join(MG.OUT, WEIGHTS + '.npz'), **{k: v.numpy() for k, v in sd.items()})
    out = MS.compact(tr, arch, B_TEST)
    out['weights_checksum'] = np.float64(MG.weights_checksum(sd))
    out['x_checksum'] = np.float64(x_test.double().sum().item())
    out['eps_checksum'] = np.float64(eps_test.double().abs().sum().item())
    out['train_steps'] = np.int64(steps)
    out['final_loss'] = np.float64(loss.item())
    out['reference_finite'] = np.bool_(all(np.isfinite(v).all() for v in out.values() if v.dtype.kind == 'f'))
    np.savez_compressed(os.path.join(MG.OUT, NAME + '.npz'), **out)
    m = tr['mask']
    print('mask max percentiles', np.percentile(m.max(dim=1).values.numpy(), [10, 50, 90]), 'finite', out['reference_finite'])


if __name__ == '__main__':
    torch.set_num_threads(int(os.environ.get('OMP_NUM_THREADS', '6')))
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 400)
